// qmc_sweep_lean.cu - the persistent Metropolis kernel with the lean evaluator (14 warps / SM).
// Same per-proposal protocol as k_sweep (qmc_sweep.cu; sampler.py:104-155): the only difference
// is warp_eval_flip_lean, which keeps ~half the shared memory per warp so that twice as many
// chains are resident per SM.  Single-flip proposals of deep models only; everything else runs
// the classic kernel.
#include "qmc_host.h"

namespace qmc {

constexpr int kLeanMaxWarps = 14;     // 448 threads -> 146 registers per thread
constexpr int kLeanAcc = 32;          // accumulators per lane that fit that budget

__global__ void __maxnreg__(136)     // 14 warps x 32 lanes x 144 registers = 64512 <= 65536
k_sweep_lean(DevModel m, const float* __restrict__ params, SweepArgs a, LeanPlan lp, int newf_floats,
             int per_warp_bytes, int staging_floats, int allow_tiled) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    load_params_to_smem(m, params, smem_f);
    const float* sp = smem_f;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    char* wmem = reinterpret_cast<char*>(smem_f + m.smem_param_floats) + (size_t)warp * per_warp_bytes;
    float* arena = reinterpret_cast<float*>(wmem);
    float* newf = arena + lp.arena_floats;
    int8_t* spins_s = reinterpret_cast<int8_t*>(newf + newf_floats);

    const int n = m.n, p = m.p, Ly = m.Ly, Lx = m.Lx;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    float* staging = a.staging + (size_t)slot * staging_floats;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const int lside = 1 + 2 * m.D * p;                     // window of the last layer
    unsigned long long accepted = 0;

    for (int chain = slot; chain < a.S; chain += nslots) {
        int8_t* gspins = a.spins + (size_t)chain * n;
        float* cache = a.cache + (size_t)chain * m.cache_floats;
        for (int i = lane; i < n; i += kWarp) spins_s[i] = gspins[i];
        __syncwarp();
        const unsigned long long gchain = (unsigned long long)(a.chain_id0 + chain);
        for (long long it = 0; it < a.n_steps; ++it) {
            const long long step = a.step0 + it;
            int f0;
            float u;
            if (a.flip_pos) {
                f0 = a.flip_pos[(size_t)it * a.S + chain];
                u = a.uniforms[(size_t)it * a.S + chain];
            } else {
                const uint4 r = philox4x32_10(
                    make_uint4((uint32_t)step, (uint32_t)((unsigned long long)step >> 32),
                               (uint32_t)gchain, (uint32_t)(gchain >> 32)), key);
                f0 = (int)__umulhi(r.x, (uint32_t)n);
                u = (float)(r.w >> 8) * 5.9604644775390625e-8f;
            }
            f0 = __shfl_sync(0xffffffffu, f0, 0);
            u = __shfl_sync(0xffffffffu, u, 0);
            float dre;
            warp_eval_flip_lean<kLeanAcc>(m, lp, sp, arena, spins_s, cache, staging, newf, f0, lane,
                                          allow_tiled, dre);
            const float amp = expf(dre);
            const bool accept = __shfl_sync(0xffffffffu, (int)(amp * amp > u), 0) != 0;   // strict, sampler.py:125
            if (accept) {
                const int y0 = f0 / Lx, x0 = f0 - y0 * Lx;
                int stg = 0;
                for (int l = 0; l < m.D - 1; ++l) {
                    const LayerInfo& L = m.layer[l];
                    const int side = 1 + 2 * (l + 1) * p, rarea = side * side, ncg = L.coutp >> 2;
                    const int ry = y0 - (l + 1) * p, rx = x0 - (l + 1) * p;
                    float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
                    const FastDiv dside(side), darea(rarea);
                    for (int base = 0; base < ncg * rarea; base += 4 * kWarp) {
                        float4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int idx = base + j * kWarp + lane;
                            if (idx < ncg * rarea) v[j] = ldcg4(staging + stg + (size_t)idx * 4);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int idx = base + j * kWarp + lane;
                            if (idx >= ncg * rarea) continue;
                            const int cg = darea.div(idx), pos = idx - cg * rarea;
                            const int y = dside.div(pos), x = pos - y * side;
                            plane4[cg * n + wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx)] = v[j];
                        }
                    }
                    stg += L.coutp * rarea;
                }
                const int ry = y0 - m.D * p, rx = x0 - m.D * p;
                const FastDiv dls(lside);
                for (int pos = lane; pos < lside * lside; pos += kWarp) {
                    const int y = dls.div(pos), x = pos - y * lside;
                    cache[m.fre_off + wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx)] = newf[pos];
                }
                if (lane == 0) spins_s[f0] = -spins_s[f0];
                __syncwarp();
                ++accepted;
            }
            if (lane == 0) {
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = dre;
            }
            if (a.samples && step >= a.therm_its && (step - a.therm_its) % a.its_per_sample == 0) {
                const long long j = (step - a.therm_its) / a.its_per_sample;
                if (j < a.n_sample_slots) {
                    int8_t* dst = a.samples + ((size_t)j * a.S + chain) * n;
                    for (int i = lane; i < n; i += kWarp) dst[i] = spins_s[i];
                }
            }
        }
        for (int i = lane; i < n; i += kWarp) gspins[i] = spins_s[i];
        __syncwarp();
    }
    if (a.n_accept && lane == 0 && accepted) atomicAdd(a.n_accept, accepted);
}

LeanLaunch lean_launch_plan(const qmc_handle* h, int S) {
    LeanLaunch best{};
    best.ok = false;
    const DevModel& m = h->m;
    if (!h->allow_lean || !h->allow_tiled || m.D < 2 || m.r > m.Ly || m.r > m.Lx) return best;
    const int lside = 1 + 2 * m.D * m.p;
    const int newf = round4(lside * lside), spins_bytes = (m.n + 15) & ~15;
    const size_t param_bytes = (size_t)m.smem_param_floats * 4;
    int staging = 0;
    for (int l = 0; l < m.D - 1; ++l) {
        const int side = 1 + 2 * (l + 1) * m.p;
        staging += m.layer[l].coutp * side * side;
    }
    double best_score = -1;
    int wcap = kLeanMaxWarps;
    if (h->max_warps_override > 0 && wcap > h->max_warps_override) wcap = h->max_warps_override;
    for (int w = wcap; w >= 9; --w) {          // below 9 warps the classic kernel's bigger tiles win
        if (param_bytes + (size_t)w * (spins_bytes + 4 * (newf + 64)) > h->max_smem) continue;
        const long long per_warp_budget = (long long)((h->max_smem - param_bytes) / w) - spins_bytes - 4LL * newf;
        if (per_warp_budget < 256) continue;
        LeanPlan lp = lean_plan(m, (int)(per_warp_budget / 4));
        if (!lp.ok) continue;
        const long long slots = (long long)h->num_sms * w, waves = (S + slots - 1) / slots;
        const double eff = (double)S / (double)(waves * slots);
        int extra_bands = 0;
        for (int l = lp.first_gather; l < m.D; ++l) extra_bands += lp.bands[l] - 1;
        const double score = eff * (1.0 + 0.02 * w) * (1.0 - 0.03 * extra_bands);
        if (score > best_score) {
            best_score = score;
            best.lp = lp; best.warps = w; best.newf_floats = newf; best.spins_bytes = spins_bytes;
            best.staging_floats = round4(staging);
            const size_t per_warp = (size_t)(lp.arena_floats + newf) * 4 + spins_bytes;
            best.smem = param_bytes + per_warp * w;
            long long ctas = ((long long)S + w - 1) / w;
            best.grid = (int)(ctas < h->num_sms ? ctas : h->num_sms);
            best.ok = true;
        }
    }
    return best;
}

cudaError_t launch_sweep_lean(const qmc_handle* h, const SweepArgs& a, const LeanLaunch& ll, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ll.smem);
    if (e != cudaSuccess) return e;
    const int per_warp = (ll.lp.arena_floats + ll.newf_floats) * 4 + ll.spins_bytes;
    ++g_launches;
    k_sweep_lean<<<ll.grid, ll.warps * 32, ll.smem, st>>>(h->m, h->d_params, a, ll.lp, ll.newf_floats, per_warp,
                                                         ll.staging_floats, h->allow_tiled ? 1 : 0);
    return cudaGetLastError();
}

} // namespace qmc

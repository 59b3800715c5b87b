// qmc_sweep.cu - K2: persistent batched Metropolis sweep.
// Replaces Sampler.mcmc_step / the while_loop of Sampler.mcmc_op
// (sampler.py:104-155, 168-174): n_steps strictly sequential proposals for S
// independent chains, one launch.
//
// One warp owns a chain for the whole launch: its lattice lives in shared
// memory, its activation cache in HBM/L2.  Per proposal the warp draws the
// flip sites and the acceptance uniform (Philox-4x32-10, or the fed-in arrays
// in parity mode), recomputes only the affected receptive-field windows
// (warp_eval_flip), reduces the log-ratio with warp shuffles, and on accept
// commits the new window activations to the cache.  No block-level barrier is
// executed inside the step loop.
#include "qmc_host.h"

// Compiled twice (Makefile): QMC_MAXW=8 (<= 8 warps per CTA, 255 registers, big register
// tiles - the large-model variant) and QMC_MAXW=16 (128 registers, small models).
#ifndef QMC_MAXW
#define QMC_MAXW 8
#endif

namespace qmc {

#define QMC_CAT2(a, b) a##b
#define QMC_CAT(a, b) QMC_CAT2(a, b)
#define K_SWEEP QMC_CAT(k_sweep_w, QMC_MAXW)
constexpr bool kBig = QMC_MAXW <= 8;

__global__ void __launch_bounds__(QMC_MAXW * 32, 1)
K_SWEEP(DevModel m, const float* __restrict__ params, SweepArgs a, EvalPlan pl, int allow_tiled) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    load_params_to_smem(m, params, smem_f);
    const float* sp = smem_f;
    // broadcast from lane 0 so the compiler knows the warp index (and everything derived from it:
    // chain, task, loop bounds) is warp-uniform and may use the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    char* wmem = reinterpret_cast<char*>(smem_f + m.smem_param_floats) + (size_t)warp * pl.per_warp_bytes;
    float* buf0 = reinterpret_cast<float*>(wmem);
    float* buf1 = buf0 + pl.buf_floats[0];
    float* newf = buf1 + pl.buf_floats[1];
    int8_t* spins_s = reinterpret_cast<int8_t*>(newf + pl.newf_floats);

    const int n = m.n, p = m.p, Ly = m.Ly, Lx = m.Lx;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    float* staging = a.staging + (size_t)slot * pl.staging_floats;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    unsigned long long accepted = 0;

    for (int chain = slot; chain < a.S; chain += nslots) {
        int8_t* gspins = a.spins + (size_t)chain * n;
        float* cache = a.cache + (size_t)chain * m.cache_floats;
        for (int i = lane; i < n; i += kWarp) spins_s[i] = gspins[i];
        __syncwarp();
        const unsigned long long gchain = (unsigned long long)(a.chain_id0 + chain);
        for (long long it = 0; it < a.n_steps; ++it) {
            const long long step = a.step0 + it;
            int f0, f1 = -1;
            float u;
            if (a.flip_pos) {
                const int32_t* fp = a.flip_pos + ((size_t)it * a.S + chain) * a.num_flips;
                f0 = fp[0];
                if (a.num_flips > 1) f1 = fp[1];
                u = a.uniforms[(size_t)it * a.S + chain];
            } else {
                const uint4 r = philox4x32_10(
                    make_uint4((uint32_t)step, (uint32_t)((unsigned long long)step >> 32),
                               (uint32_t)gchain, (uint32_t)(gchain >> 32)), key);
                f0 = (int)__umulhi(r.x, (uint32_t)n);
                if (a.num_flips > 1) f1 = (int)__umulhi(r.y, (uint32_t)n);
                u = (float)(r.w >> 8) * 5.9604644775390625e-8f;   // 2^-24
            }
            // loaded values are warp-uniform by construction; tell the compiler (uniform datapath)
            f0 = __shfl_sync(0xffffffffu, f0, 0);
            f1 = __shfl_sync(0xffffffffu, f1, 0);
            u = __shfl_sync(0xffffffffu, u, 0);
            bool accept;
            float dre = 0.f;
            if (a.num_flips > 1 && f0 == f1) {
                // two flips of the same site cancel (sampler.py:114-115): ratio 1 > u always
                accept = 1.0f > u;
            } else {
                const FlipBox box = make_box(m, a.num_flips, f0, f1);
                Region reg;
                float dim;
                warp_eval_flip<false, kBig>(m, sp, buf0, buf1, spins_s, cache, staging, newf, pl.nfstride,
                                      box, lane, allow_tiled, reg, dre, dim);
                const float amp = expf(dre);            // |exp(z)| = exp(Re z)
                accept = __shfl_sync(0xffffffffu, (int)(amp * amp > u), 0) != 0;   // strict, sampler.py:125
                if (accept) {
                    // commit: new hidden activations (from staging), new factors, spins
                    int rh = box.h0 + 2 * p, rw = box.w0 + 2 * p;
                    if (rh > Ly) rh = Ly;
                    if (rw > Lx) rw = Lx;
                    int ry = box.y0 - p, rx = box.x0 - p, stg = 0;
                    for (int l = 0; l < m.D - 1; ++l) {
                        const LayerInfo& L = m.layer[l];
                        const int rarea = rh * rw, ncg = L.coutp >> 2;
                        float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
                        const FastDiv drw(rw), darea(rarea);
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
                                const int y = drw.div(pos), x = pos - y * rw;
                                const int site = wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx);
                                plane4[cg * n + site] = v[j];
                            }
                        }
                        stg += L.coutp * rarea;
                        ry -= p; rx -= p; rh += 2 * p; rw += 2 * p;
                    }
                    const FastDiv dregw(reg.rw);
                    for (int pos = lane; pos < reg.rh * reg.rw; pos += kWarp) {
                        const int y = dregw.div(pos), x = pos - y * reg.rw;
                        const int site = wrap1(reg.ry + y, Ly) * Lx + wrap1(reg.rx + x, Lx);
                        cache[m.fre_off + site] = newf[pos];
                    }
                    if (lane == 0) {
                        spins_s[f0] = -spins_s[f0];
                        if (a.num_flips > 1) spins_s[f1] = -spins_s[f1];
                    }
                    __syncwarp();
                }
            }
            if (accept) ++accepted;
            if (lane == 0) {
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = dre;
            }
            // sample write-out AFTER the update (sampler.py:135-152)
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

cudaError_t QMC_CAT(launch_sweep_w, QMC_MAXW)(const qmc_handle* h, const SweepArgs& a, const EvalPlan& pl,
                                              const WarpGrid& g, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(K_SWEEP, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    ++g_launches;
    K_SWEEP<<<g.grid, g.warps * 32, g.smem, st>>>(h->m, h->d_params, a, pl, h->allow_tiled ? 1 : 0);
    return cudaGetLastError();
}

#if QMC_MAXW == 8
cudaError_t launch_sweep_w16(const qmc_handle* h, const SweepArgs& a, const EvalPlan& pl,
                             const WarpGrid& g, cudaStream_t st);

int sweep_slots(const qmc_handle* h, int S, int num_flips, EvalPlan* plan, WarpGrid* grid) {
    const DevModel& m = h->m;
    int h0 = 1, w0 = 1;
    if (num_flips > 1) { h0 = m.Ly / 2 + 1; w0 = m.Lx / 2 + 1; }
    if (!box_supported(m, h0, w0)) return -1;
    EvalPlan pl = eval_plan(m, h0, w0, false);
    WarpGrid g = pick_warp_grid(h, pl.per_warp_bytes, 0, S);
    if (!g.ok) return -2;
    if (plan) *plan = pl;
    if (grid) *grid = g;
    return g.grid * g.warps;
}

cudaError_t launch_sweep(const qmc_handle* h, const SweepArgs& a, cudaStream_t st, std::string& err) {
    EvalPlan pl; WarpGrid g;
    const int slots = sweep_slots(h, a.S, a.num_flips, &pl, &g);
    if (slots == -1) { err = "sweep: flip box does not fit the lattice (need h0 + r - 1 <= L for deep models)"; return cudaErrorInvalidValue; }
    if (slots < 0) { err = "sweep: model does not fit in shared memory"; return cudaErrorInvalidValue; }
    return g.warps <= 8 ? launch_sweep_w8(h, a, pl, g, st) : launch_sweep_w16(h, a, pl, g, st);
}
#endif

} // namespace qmc

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

// Compiled three times (Makefile): QMC_MAXW=8 (<= 8 warps per CTA, 255 registers, big register
// tiles - the large-model variant), QMC_MAXW=16 (128 registers, small models) and QMC_MAXW=28
// (72 registers, 7 warps per scheduler: models with <= 8 channels per layer, whose proposals
// are a few thousand instructions of mostly latency - C2: 4096 chains are ONE wave of 148 x 28
// warps, 135 -> 168 M proposals/s, profiles/r02_summary.md, "Classic kernels").
#ifndef QMC_MAXW
#define QMC_MAXW 8
#endif

namespace qmc {

#define QMC_CAT2(a, b) a##b
#define QMC_CAT(a, b) QMC_CAT2(a, b)
#define K_SWEEP QMC_CAT(k_sweep_w, QMC_MAXW)
constexpr int kAcc = QMC_MAXW <= 8 ? 64 : 32;   // accumulators per lane the register budget allows

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
                warp_eval_flip<false, kAcc>(m, sp, buf0, buf1, spins_s, cache, staging, newf, pl.nfstride,
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
    K_SWEEP<<<g.grid, g.warps * 32, g.smem, st>>>(h->m, h->d_params_padded, a, pl, h->allow_tiled ? 1 : 0);
    return cudaGetLastError();
}

#if QMC_MAXW == 16
// ------------------------------------------------------------------------------------------
// Symmetric sweep: Metropolis sampling of |psi_sym|^2, psi_sym = (1/nsym) sum_g psi(.; W o g)
// (BASELINE config 4; SURVEY.md section 8 defines the amplitude, symmetry.ipynb the group).
// The warp evaluates the proposal against every image with the same warp_eval_flip (its own
// parameter block, cache, staging and new-factor buffer per image) and accepts with
//   |sum_g exp(D_g + delta_g) / sum_g exp(D_g)|^2 > u,   D_g = log psi_g - log psi_0 (double).
// Generic conv only (the models this is used with are small), 16-warp variant only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
k_sweep_sym(DevModel m, const float* __restrict__ sym_padded, SweepArgs a, EvalPlan pl, int nsym,
            double* __restrict__ drel_all) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    const int pb = m.smem_param_floats;
    for (int i = threadIdx.x; i < nsym * pb; i += blockDim.x) smem_f[i] = sym_padded[i];
    __syncthreads();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    char* wmem = reinterpret_cast<char*>(smem_f + nsym * pb) + (size_t)warp * pl.per_warp_bytes;
    float* buf0 = reinterpret_cast<float*>(wmem);
    float* buf1 = buf0 + pl.buf_floats[0];
    float* newf = buf1 + pl.buf_floats[1];                      // nsym x (Re, Im) x nfstride
    int8_t* spins_s = reinterpret_cast<int8_t*>(newf + nsym * pl.newf_floats);
    const int n = m.n, p = m.p, Ly = m.Ly, Lx = m.Lx;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    float* staging0 = a.staging + (size_t)slot * nsym * pl.staging_floats;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const size_t cache_img = (size_t)a.S * m.cache_floats;     // stride between image caches
    unsigned long long accepted = 0;
    for (int chain = slot; chain < a.S; chain += nslots) {
        int8_t* gspins = a.spins + (size_t)chain * n;
        double* drel = drel_all + (size_t)chain * nsym * 2;
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
                u = (float)(r.w >> 8) * 5.9604644775390625e-8f;
            }
            bool accept;
            float lr = 0.f;
            if (a.num_flips > 1 && f0 == f1) {
                accept = 1.0f > u;
            } else {
                const FlipBox box = make_box(m, a.num_flips, f0, f1);
                Region reg;
                double nre = 0, nim = 0, dre_ = 0, dim_ = 0, dl_re[8], dl_im[8];
                for (int g = 0; g < nsym; ++g) {
                    float dre, dim;
                    warp_eval_flip<true, 32, false>(m, smem_f + g * pb, buf0, buf1, spins_s,
                                                       a.cache + g * cache_img + (size_t)chain * m.cache_floats,
                                                       staging0 + (size_t)g * pl.staging_floats,
                                                       newf + g * pl.newf_floats, pl.nfstride, box, lane, 0, reg,
                                                       dre, dim);
                    dl_re[g] = dre; dl_im[g] = dim;
                    // weights exp(D_g) are normalised by exp(max Re D) implicitly: D_0 = 0 and the
                    // images differ by O(1), so plain exp in double is safe
                    const double wr = exp(drel[2 * g]), wi = drel[2 * g + 1];
                    const double cr = wr * cos(wi), ci = wr * sin(wi);
                    dre_ += cr; dim_ += ci;
                    const double er = exp((double)dre), ei = (double)dim;
                    const double rr = er * cos(ei), ri = er * sin(ei);
                    nre += cr * rr - ci * ri;
                    nim += cr * ri + ci * rr;
                }
                const double den = dre_ * dre_ + dim_ * dim_;
                const double prob = (nre * nre + nim * nim) / den;      // |psi_sym(s') / psi_sym(s)|^2
                lr = (float)(0.5 * log(prob));
                accept = __shfl_sync(0xffffffffu, (int)((float)prob > u), 0) != 0;
                if (accept) {
                    for (int g = 0; g < nsym; ++g) {
                        float* cache = a.cache + g * cache_img + (size_t)chain * m.cache_floats;
                        const float* stg_g = staging0 + (size_t)g * pl.staging_floats;
                        const float* nf = newf + g * pl.newf_floats;
                        int rh = box.h0 + 2 * p, rw = box.w0 + 2 * p;
                        if (rh > Ly) rh = Ly;
                        if (rw > Lx) rw = Lx;
                        int ry = box.y0 - p, rx = box.x0 - p, stg = 0;
                        for (int l = 0; l < m.D - 1; ++l) {
                            const LayerInfo& L = m.layer[l];
                            const int rarea = rh * rw, ncg = L.coutp >> 2;
                            float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
                            const FastDiv drw(rw), darea(rarea);
                            for (int idx = lane; idx < ncg * rarea; idx += kWarp) {
                                const int cg = darea.div(idx), pos = idx - cg * rarea;
                                const int y = drw.div(pos), x = pos - y * rw;
                                plane4[cg * n + wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx)] =
                                    ldcg4(stg_g + stg + (size_t)idx * 4);
                            }
                            stg += L.coutp * rarea;
                            ry -= p; rx -= p; rh += 2 * p; rw += 2 * p;
                        }
                        const FastDiv dregw(reg.rw);
                        for (int pos = lane; pos < reg.rh * reg.rw; pos += kWarp) {
                            const int y = dregw.div(pos), x = pos - y * reg.rw;
                            const int site = wrap1(reg.ry + y, Ly) * Lx + wrap1(reg.rx + x, Lx);
                            cache[m.fre_off + site] = nf[pos];
                            cache[m.fim_off + site] = nf[pl.nfstride + pos];
                        }
                    }
                    if (lane == 0) {
                        for (int g = 1; g < nsym; ++g) {       // D_g += delta_g - delta_0
                            drel[2 * g] += dl_re[g] - dl_re[0];
                            drel[2 * g + 1] += dl_im[g] - dl_im[0];
                        }
                        spins_s[f0] = -spins_s[f0];
                        if (a.num_flips > 1) spins_s[f1] = -spins_s[f1];
                    }
                    __syncwarp();
                }
            }
            if (accept) ++accepted;
            if (lane == 0) {
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = lr;
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

int sweep_sym_slots(const qmc_handle* h, int S, int num_flips, int nsym, EvalPlan* plan, WarpGrid* grid) {
    const DevModel& m = h->m;
    if (nsym < 1 || nsym > 8) return -3;
    int h0 = 1, w0 = 1;
    if (num_flips > 1) { h0 = m.Ly / 2 + 1; w0 = m.Lx / 2 + 1; }
    if (!box_supported(m, h0, w0)) return -1;
    EvalPlan pl = eval_plan(m, h0, w0, true);
    const size_t per_warp = pl.per_warp_bytes + (size_t)(nsym - 1) * pl.newf_floats * 4;
    const size_t extra_params = (size_t)(nsym - 1) * m.smem_param_floats * 4;
    WarpGrid g = pick_warp_grid(h, per_warp, extra_params, S);
    if (!g.ok) return -2;
    pl.per_warp_bytes = per_warp;
    if (plan) *plan = pl;
    if (grid) *grid = g;
    return g.grid * g.warps;
}

cudaError_t launch_sweep_sym(const qmc_handle* h, const SweepArgs& a, int nsym, double* drel, cudaStream_t st,
                             std::string& err) {
    EvalPlan pl; WarpGrid g;
    const int slots = sweep_sym_slots(h, a.S, a.num_flips, nsym, &pl, &g);
    if (slots == -1) { err = "sweep_sym: flip box does not fit the lattice"; return cudaErrorInvalidValue; }
    if (slots < 0) { err = "sweep_sym: nsym images of the model do not fit in shared memory"; return cudaErrorInvalidValue; }
    cudaError_t e = cudaFuncSetAttribute(k_sweep_sym, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    ++g_launches;
    k_sweep_sym<<<g.grid, g.warps * 32, g.smem, st>>>(h->m, h->d_sym_padded, a, pl, nsym, drel);
    return cudaGetLastError();
}
#endif

#if QMC_MAXW == 8
cudaError_t launch_sweep_w16(const qmc_handle* h, const SweepArgs& a, const EvalPlan& pl,
                             const WarpGrid& g, cudaStream_t st);
cudaError_t launch_sweep_w28(const qmc_handle* h, const SweepArgs& a, const EvalPlan& pl,
                             const WarpGrid& g, cudaStream_t st);

// every layer has <= 8 output channels: the register tiles need <= 32 accumulators per lane and the
// kernel fits 72 registers (the 28-warp object) without spilling its loops
static bool narrow_model(const DevModel& m) {
    for (int l = 0; l < m.D; ++l)
        if (m.layer[l].cout > 8) return false;
    return true;
}

int sweep_slots(const qmc_handle* h, int S, int num_flips, EvalPlan* plan, WarpGrid* grid) {
    const DevModel& m = h->m;
    int h0 = 1, w0 = 1;
    if (num_flips > 1) { h0 = m.Ly / 2 + 1; w0 = m.Lx / 2 + 1; }
    if (!box_supported(m, h0, w0)) return -1;
    EvalPlan pl = eval_plan(m, h0, w0, false);
    WarpGrid g = pick_warp_grid(h, pl.per_warp_bytes, 0, S, narrow_model(m) ? 28 : 16);
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
    return g.warps <= 8 ? launch_sweep_w8(h, a, pl, g, st)
         : g.warps <= 16 ? launch_sweep_w16(h, a, pl, g, st) : launch_sweep_w28(h, a, pl, g, st);
}
#endif

} // namespace qmc

// qmc_energy.cu - K3: local energies, every connected configuration of every
// sample as a receptive-field delta.
// Replaces ising_energy (mcmc_tf.py:59-90) and heisenberg_energy
// (mcmc_tf.py:93-141) including all_windows / interactions
// (helpers.py:149-195): instead of gathering N*L^2 windows of (2K-1)^2 spins
// and running the full network on each, a warp evaluates each flipped
// configuration incrementally against the sample's activation cache (filled by
// K1) and accumulates exp(log_pop) in registers.
#include "qmc_host.h"

// Compiled four times (Makefile): QMC_MAXW=8 (255 registers) / 16 (128 registers) like qmc_sweep.cu, times
// QMC_HAM=0 (TFIM) / 1 (Heisenberg) - one Hamiltonian per translation unit, because ptxas needs minutes for a
// kernel that inlines the evaluator for both (the clean build went from 6.5 to ~3 minutes).
#ifndef QMC_MAXW
#define QMC_MAXW 8
#endif
#ifndef QMC_HAM
#define QMC_HAM 0
#endif
#define QMC_CAT2(a, b) a##b
#define QMC_CAT(a, b) QMC_CAT2(a, b)
#define QMC_CAT4(a, b, c, d) QMC_CAT(QMC_CAT(a, b), QMC_CAT(c, d))
#define K_ENERGY QMC_CAT4(k_energy_w, QMC_MAXW, _h, QMC_HAM)

namespace qmc {

constexpr int kAcc = QMC_MAXW <= 8 ? 64 : 32;

constexpr int kEnergyChunks = 16;  // site chunks per sample (warp tasks = N * chunks)

__global__ void __launch_bounds__(QMC_MAXW * 32, 1)
K_ENERGY(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins, int N,
         const float* __restrict__ cache_all, float2* __restrict__ partial,
         int nchunks, EvalPlan pl, int allow_tiled, ImageStrides is) {
    constexpr int hamiltonian = QMC_HAM;
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    // symmetry images (blockIdx.y): own parameter block and cache, partial sums [image][sample][chunk]
    params += (size_t)blockIdx.y * is.params;
    cache_all += (size_t)blockIdx.y * is.cache;
    partial += (size_t)blockIdx.y * N * nchunks;
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

    const int n = m.n, Ly = m.Ly, Lx = m.Lx;
    const int cs = (n + nchunks - 1) / nchunks;
    const long long ntasks = (long long)N * nchunks;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    int loaded = -1;
    for (long long task = slot; task < ntasks; task += nslots) {
        const int s = (int)(task / nchunks), chunk = (int)(task - (long long)s * nchunks);
        if (s != loaded) {
            __syncwarp();
            for (int i = lane; i < n; i += kWarp) spins_s[i] = spins[(size_t)s * n + i];
            __syncwarp();
            loaded = s;
        }
        const float* cache = cache_all + (size_t)s * m.cache_floats;
        float are = 0.f, aim = 0.f;
        const int i1 = min(n, (chunk + 1) * cs);
        for (int i = chunk * cs; i < i1; ++i) {
            Region reg;
            float dre, dim, sn, cn;
            if (hamiltonian == QMC_HAMILTONIAN_TFIM) {
                const FlipBox box = make_box(m, 1, i, -1);
                warp_eval_flip<true, kAcc>(m, sp, buf0, buf1, spins_s, cache, nullptr, newf, pl.nfstride, box,
                                     lane, allow_tiled, reg, dre, dim);
                const float amp = expf(dre);
                sincosf(dim, &sn, &cn);
                are += amp * cn;               // exp(log_pop), mcmc_tf.py:88
                aim += amp * sn;
            } else {
                const int y = i / Lx, x = i - y * Lx;
                for (int d = 0; d < 2; ++d) {
                    const int j = d == 0 ? (y + 1 == Ly ? 0 : y + 1) * Lx + x
                                         : y * Lx + (x + 1 == Lx ? 0 : x + 1);
                    if (j == i) { are += 1.f; continue; }          // L == 1 along d: s_i s_i = 1
                    const bool aligned = __shfl_sync(0xffffffffu, (int)(spins_s[i] == spins_s[j]), 0) != 0;
                    if (aligned) { are += 1.f; continue; }                   // -(1-1) exp + 1
                    const FlipBox box = make_box(m, 2, i, j);
                    warp_eval_flip<true, kAcc>(m, sp, buf0, buf1, spins_s, cache, nullptr, newf, pl.nfstride,
                                         box, lane, allow_tiled, reg, dre, dim);
                    const float amp = expf(dre);
                    sincosf(dim, &sn, &cn);
                    are += -2.f * amp * cn - 1.f;      // -(1-(-1)) exp(log_pop) + (-1), mcmc_tf.py:138
                    aim += -2.f * amp * sn;
                }
            }
        }
        if (lane == 0) partial[(size_t)s * nchunks + chunk] = make_float2(are, aim);
    }
}

cudaError_t QMC_CAT4(launch_energy_main_w, QMC_MAXW, _h, QMC_HAM)(const qmc_handle* h, const int8_t* spins, int N,
                                                                  const float* cache, float2* partial, int nchunks,
                                                                  const EvalPlan& pl, const WarpGrid& g,
                                                                  cudaStream_t st, int nimg, const float* blocks) {
    cudaError_t e = cudaFuncSetAttribute(K_ENERGY, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    ++g_launches;
    const ImageStrides is{(size_t)h->m.smem_param_floats, (size_t)N * h->m.cache_floats};
    const int gx = nimg > 1 ? (g.grid / nimg > 0 ? g.grid / nimg : 1) : g.grid;     // all images share the SMs
    K_ENERGY<<<dim3(gx, nimg), g.warps * 32, g.smem, st>>>(h->m, blocks ? blocks : h->d_params_padded, spins, N, cache,
                                                           partial, nchunks, pl, h->allow_tiled ? 1 : 0, is);
    return cudaGetLastError();
}

#if QMC_MAXW == 8 && QMC_HAM == 0
#define QMC_DECL_ENERGY_MAIN(W, H)                                                                              \
    cudaError_t launch_energy_main_w##W##_h##H(const qmc_handle* h, const int8_t* spins, int N, const float* cache, \
                                               float2* partial, int nchunks, const EvalPlan& pl, const WarpGrid& g, \
                                               cudaStream_t st, int nimg, const float* blocks);
QMC_DECL_ENERGY_MAIN(8, 1)
QMC_DECL_ENERGY_MAIN(16, 0)
QMC_DECL_ENERGY_MAIN(16, 1)
#undef QMC_DECL_ENERGY_MAIN

// one thread per sample: ordered chunk sum, diagonal term, per-spin normalisation
__global__ void k_energy_finish(DevModel m, const int8_t* __restrict__ spins, int N, int hamiltonian,
                                float field_h, const float2* __restrict__ partial, int nchunks,
                                float2* __restrict__ e_loc, double* __restrict__ moments) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double mre = 0, mim = 0, msq = 0, cnt = 0;
    if (s < N) {
        float re = 0.f, im = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const float2 v = partial[(size_t)s * nchunks + c];
            re += v.x; im += v.y;
        }
        if (hamiltonian == QMC_HAMILTONIAN_TFIM) {
            int aligned = 0;                        // helpers.py:171-195 interactions, summed
            const int8_t* sp = spins + (size_t)s * m.n;
            for (int y = 0; y < m.Ly; ++y)
                for (int x = 0; x < m.Lx; ++x) {
                    const int c = sp[y * m.Lx + x];
                    aligned += c * sp[(y + 1 == m.Ly ? 0 : y + 1) * m.Lx + x];
                    aligned += c * sp[y * m.Lx + (x + 1 == m.Lx ? 0 : x + 1)];
                }
            re = -field_h * re - (float)aligned;    // mcmc_tf.py:88-89
            im = -field_h * im;
        }
        re /= (float)m.n; im /= (float)m.n;         // mcmc_tf.py:90 / 141
        e_loc[s] = make_float2(re, im);
        mre = re; mim = im; msq = (double)re * re + (double)im * im; cnt = 1;
    }
    if (moments) {
        for (int o = 16; o > 0; o >>= 1) {
            mre += __shfl_xor_sync(0xffffffffu, mre, o);
            mim += __shfl_xor_sync(0xffffffffu, mim, o);
            msq += __shfl_xor_sync(0xffffffffu, msq, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0 && cnt > 0) {
            atomicAdd(moments + 0, cnt); atomicAdd(moments + 1, mre);
            atomicAdd(moments + 2, mim); atomicAdd(moments + 3, msq);
        }
    }
}

cudaError_t launch_energy_finish(const qmc_handle* h, const int8_t* spins, int N, int hamiltonian, float field_h,
                                 const float2* partial, int nchunks, float* e_loc, double* moments,
                                 cudaStream_t st) {
    ++g_launches;
    k_energy_finish<<<(N + 127) / 128, 128, 0, st>>>(h->m, spins, N, hamiltonian, field_h, partial, nchunks,
                                                    reinterpret_cast<float2*>(e_loc), moments);
    return cudaGetLastError();
}

int energy_chunks(const qmc_handle* h) { return h->m.n < kEnergyChunks ? h->m.n : kEnergyChunks; }

cudaError_t launch_energy(const qmc_handle* h, int hamiltonian, float field_h, const int8_t* spins,
                          int N, float* workspace, float* e_loc, double* moments, cudaStream_t st,
                          std::string& err) {
    const DevModel& m = h->m;
    const bool heis = hamiltonian == QMC_HAMILTONIAN_HEISENBERG;
    const int h0 = heis ? 2 : 1;
    if (!box_supported(m, h0, h0)) {
        err = "local_energy: receptive field (+1 for Heisenberg bonds) exceeds the lattice";
        return cudaErrorInvalidValue;
    }
    float* cache = workspace;
    const int nchunks = energy_chunks(h);
    float2* partial = reinterpret_cast<float2*>(workspace + (size_t)N * m.cache_floats);
    cudaError_t e = launch_forward(h, spins, N, cache, nullptr, nullptr, st, err);
    if (e != cudaSuccess) return e;
    // TFIM: the in-place persistent kernel (k_energy_ip) when the model is inside its coverage and big enough for
    // it to be the sweep's choice too, else the classic persistent kernel; QMC_FLAG_ENERGY_* force one of the two
    // (the two add the same per-chunk terms in the same order: equal bits)
    const bool want_ip = h->energy_path == 2 || (h->energy_path == 0 && ip_launch_plan(h, 1 << 20).ok);
    if (!heis && want_ip && energy_ip_supported(h)) {
        e = launch_energy_ip(h, spins, N, cache, partial, nchunks, st);
        if (e != cudaSuccess) return e;
        return launch_energy_finish(h, spins, N, hamiltonian, field_h, partial, nchunks, e_loc, moments, st);
    }
    EvalPlan pl = eval_plan(m, h0, h0, true);
    WarpGrid g = pick_warp_grid(h, pl.per_warp_bytes, 0, (long long)N * nchunks);
    if (!g.ok) { err = "local_energy: model does not fit in shared memory"; return cudaErrorInvalidValue; }
    if (heis)
        e = g.warps <= 8 ? launch_energy_main_w8_h1(h, spins, N, cache, partial, nchunks, pl, g, st, 1, nullptr)
                         : launch_energy_main_w16_h1(h, spins, N, cache, partial, nchunks, pl, g, st, 1, nullptr);
    else
        e = g.warps <= 8 ? launch_energy_main_w8_h0(h, spins, N, cache, partial, nchunks, pl, g, st, 1, nullptr)
                         : launch_energy_main_w16_h0(h, spins, N, cache, partial, nchunks, pl, g, st, 1, nullptr);
    if (e != cudaSuccess) return e;
    return launch_energy_finish(h, spins, N, hamiltonian, field_h, partial, nchunks, e_loc, moments, st);
}

// ------------------------------------------------------------------------------------------------
// Symmetry-averaged amplitude (symmetry.ipynb): psi_sym = (1/nsym) sum_g psi_g with
// psi_g = the same lattice through parameter image g.  E_loc[psi_sym](s) = sum_g p_g(s) E_loc[psi_g](s),
// p_g = psi_g(s) / sum_h psi_h(s) (exact identity, pinned in tests/test_oracle_pins.py).  Three launches:
// the forward of all images, the connected-configuration sums of all images (image = blockIdx.y of the
// kernels above), and k_energy_finish_sym, which forms log psi_g - log psi_0 in double from the per-site
// factor planes of the caches (differences first, like the sweep's log_rel), the weights p_g and the
// combination.
// ------------------------------------------------------------------------------------------------
__global__ void k_energy_finish_sym(DevModel m, int nsym, const int8_t* __restrict__ spins, int N, int hamiltonian,
                                    float field_h, const float* __restrict__ caches, const float2* __restrict__ partial,
                                    int nchunks, float2* __restrict__ e_loc, double* __restrict__ moments) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double mre = 0, mim = 0, msq = 0, cnt = 0;
    if (s < N) {
        const size_t cimg = (size_t)N * m.cache_floats;
        const float* c0 = caches + (size_t)s * m.cache_floats;
        int aligned = 0;
        if (hamiltonian == QMC_HAMILTONIAN_TFIM) {
            const int8_t* sp = spins + (size_t)s * m.n;
            for (int y = 0; y < m.Ly; ++y)
                for (int x = 0; x < m.Lx; ++x) {
                    const int c = sp[y * m.Lx + x];
                    aligned += c * sp[(y + 1 == m.Ly ? 0 : y + 1) * m.Lx + x];
                    aligned += c * sp[y * m.Lx + (x + 1 == m.Lx ? 0 : x + 1)];
                }
        }
        double wsum_re = 0, wsum_im = 0, ere = 0, eim = 0;
        for (int g = 0; g < nsym; ++g) {
            const float* cg = c0 + g * cimg;
            double dre = 0, dim = 0;                      // log psi_g - log psi_0
            if (g > 0)
                for (int i = 0; i < m.n; ++i) {
                    dre += (double)cg[m.fre_off + i] - (double)c0[m.fre_off + i];
                    dim += (double)cg[m.fim_off + i] - (double)c0[m.fim_off + i];
                }
            const double a = exp(dre), wr = a * cos(dim), wi = a * sin(dim);     // psi_g / psi_0 (images differ by O(1))
            float re = 0.f, im = 0.f;                     // E_loc[psi_g]: k_energy_finish's arithmetic
            for (int c = 0; c < nchunks; ++c) {
                const float2 v = partial[((size_t)g * N + s) * nchunks + c];
                re += v.x; im += v.y;
            }
            if (hamiltonian == QMC_HAMILTONIAN_TFIM) {
                re = -field_h * re - (float)aligned;
                im = -field_h * im;
            }
            re /= (float)m.n; im /= (float)m.n;
            wsum_re += wr; wsum_im += wi;
            ere += wr * re - wi * im;
            eim += wr * im + wi * re;
        }
        const double den = wsum_re * wsum_re + wsum_im * wsum_im;
        const float re = (float)((ere * wsum_re + eim * wsum_im) / den);        // (sum_g w_g E_g) / (sum_g w_g)
        const float im = (float)((eim * wsum_re - ere * wsum_im) / den);
        e_loc[s] = make_float2(re, im);
        mre = re; mim = im; msq = (double)re * re + (double)im * im; cnt = 1;
    }
    if (moments) {
        for (int o = 16; o > 0; o >>= 1) {
            mre += __shfl_xor_sync(0xffffffffu, mre, o);
            mim += __shfl_xor_sync(0xffffffffu, mim, o);
            msq += __shfl_xor_sync(0xffffffffu, msq, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0 && cnt > 0) {
            atomicAdd(moments + 0, cnt); atomicAdd(moments + 1, mre);
            atomicAdd(moments + 2, mim); atomicAdd(moments + 3, msq);
        }
    }
}

size_t energy_sym_workspace_floats(const qmc_handle* h, int nsym, int N) {
    return (size_t)nsym * N * h->m.cache_floats + (size_t)nsym * N * energy_chunks(h) * 2;
}

cudaError_t launch_energy_sym(const qmc_handle* h, int nsym, int hamiltonian, float field_h, const int8_t* spins, int N,
                              float* workspace, float* e_loc, double* moments, cudaStream_t st, std::string& err) {
    const DevModel& m = h->m;
    const bool heis = hamiltonian == QMC_HAMILTONIAN_HEISENBERG;
    const int h0 = heis ? 2 : 1;
    if (!box_supported(m, h0, h0)) {
        err = "local_energy_sym: receptive field (+1 for Heisenberg bonds) exceeds the lattice";
        return cudaErrorInvalidValue;
    }
    float* caches = workspace;
    const int nchunks = energy_chunks(h);
    float2* partial = reinterpret_cast<float2*>(workspace + (size_t)nsym * N * m.cache_floats);
    cudaError_t e = launch_forward_images(h, nsym, h->d_sym_padded, spins, N, caches, nullptr, nullptr, st, err);
    if (e != cudaSuccess) return e;
    EvalPlan pl = eval_plan(m, h0, h0, true);
    WarpGrid g = pick_warp_grid(h, pl.per_warp_bytes, 0, (long long)N * nchunks);
    if (!g.ok) { err = "local_energy_sym: model does not fit in shared memory"; return cudaErrorInvalidValue; }
    if (heis)
        e = g.warps <= 8 ? launch_energy_main_w8_h1(h, spins, N, caches, partial, nchunks, pl, g, st, nsym, h->d_sym_padded)
                         : launch_energy_main_w16_h1(h, spins, N, caches, partial, nchunks, pl, g, st, nsym, h->d_sym_padded);
    else
        e = g.warps <= 8 ? launch_energy_main_w8_h0(h, spins, N, caches, partial, nchunks, pl, g, st, nsym, h->d_sym_padded)
                         : launch_energy_main_w16_h0(h, spins, N, caches, partial, nchunks, pl, g, st, nsym, h->d_sym_padded);
    if (e != cudaSuccess) return e;
    ++g_launches;
    k_energy_finish_sym<<<(N + 127) / 128, 128, 0, st>>>(m, nsym, spins, N, hamiltonian, field_h, caches, partial, nchunks,
                                                        reinterpret_cast<float2*>(e_loc), moments);
    return cudaGetLastError();
}

#endif // QMC_MAXW == 8 && QMC_HAM == 0

} // namespace qmc

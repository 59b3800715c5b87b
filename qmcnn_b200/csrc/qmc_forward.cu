// qmc_forward.cu - K1: model.factors(pad(x)) and log psi for N samples.
// Replaces models.py:31-67 (CRBM.factors) / models.py:95-131 (DCRBM.factors)
// plus helpers.py:73-91 (pad) - the periodic halo is index arithmetic.
//
// One CTA per sample.  Layer by layer, each warp takes BHxBW blocks of the
// lattice, stages the block's input tile (block + halo p) in shared memory
// from the spins / the previous layer's plane in the per-sample cache, runs
// the shared conv_region routine, and writes tanh outputs back to the cache
// (hidden layers) or the per-site complex factor (last layer).  The cache is a
// by-product the sweep, energy and backward kernels consume.
#include "qmc_host.h"

namespace qmc {

constexpr int kFwdBlock = 8;   // output block side per warp task

__global__ void __launch_bounds__(256, 1)
k_forward(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins, int N,
          float* __restrict__ cache_all, float2* __restrict__ factors, float2* __restrict__ logpsi,
          int buf_in_floats, int buf_out_floats, int allow_tiled, ImageStrides is) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    // symmetry images (blockIdx.y): own parameter block, cache and outputs, the same spins
    params += (size_t)blockIdx.y * is.params;
    cache_all += (size_t)blockIdx.y * is.cache;
    if (factors) factors += (size_t)blockIdx.y * N * m.n;
    if (logpsi) logpsi += (size_t)blockIdx.y * N;
    // broadcast from lane 0 so the compiler knows the warp index (and everything derived from it:
    // chain, task, loop bounds) is warp-uniform and may use the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* wbase = smem_f + m.smem_param_floats + warp * (buf_in_floats + buf_out_floats);
    float* tin = wbase;
    float* tout = wbase + buf_in_floats;
    float* red = smem_f + m.smem_param_floats + nwarps * (buf_in_floats + buf_out_floats); // 2*blockDim floats
    int8_t* spins_s = reinterpret_cast<int8_t*>(red + 2 * blockDim.x);
    load_params_to_smem(m, params, smem_f);
    const float* sp = smem_f;

    const int p = m.p, Ly = m.Ly, Lx = m.Lx, n = m.n;
    const int nby = (Ly + kFwdBlock - 1) / kFwdBlock, nbx = (Lx + kFwdBlock - 1) / kFwdBlock;
    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        float* cache = cache_all + (size_t)s * m.cache_floats;
        for (int i = threadIdx.x; i < n; i += blockDim.x) spins_s[i] = spins[(size_t)s * n + i];
        __syncthreads();
        for (int l = 0; l < m.D; ++l) {
            const LayerInfo& L = m.layer[l];
            const bool last = (l == m.D - 1);
            for (int b = warp; b < nby * nbx; b += nwarps) {
                const int by = b / nbx, bx = b - by * nbx;
                const int ry = by * kFwdBlock, rx = bx * kFwdBlock;
                const int rh = min(kFwdBlock, Ly - ry), rw = min(kFwdBlock, Lx - rx);
                const int th = rh + 2 * p, tw = rw + 2 * p, tarea = th * tw;
                if (l == 0) {
                    for (int idx = lane; idx < tarea; idx += kWarp) {
                        const int ty = idx / tw, tx = idx - ty * tw;
                        tin[idx] = (float)spins_s[wrapi(ry - p + ty, Ly) * Lx + wrapi(rx - p + tx, Lx)];
                    }
                } else {
                    const float* plane = cache + m.layer[l - 1].act_off;
                    const int ncg = L.cinp >> 2;
                    for (int idx = lane; idx < ncg * tarea; idx += kWarp) {
                        const int cg = idx / tarea, pos = idx - cg * tarea;
                        const int ty = pos / tw, tx = pos - ty * tw;
                        const int site = wrapi(ry - p + ty, Ly) * Lx + wrapi(rx - p + tx, Lx);
                        reinterpret_cast<float4*>(tin)[idx] = ldcg4(plane + (size_t)(cg * n + site) * 4);
                    }
                }
                __syncwarp();
                if (!last) {
                    float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
                    conv_region<64>(m, l, sp, tin, tw, tarea, rh, rw, lane, allow_tiled,
                                [&](int, int y, int x, int cog, float4 a) {
                                    a.x = tanhf(a.x); a.y = tanhf(a.y); a.z = tanhf(a.z); a.w = tanhf(a.w);
                                    plane4[cog * n + (ry + y) * Lx + rx + x] = a;
                                });
                } else {
                    float4* tout4 = reinterpret_cast<float4*>(tout);
                    const int rarea = rh * rw;
                    conv_region<64>(m, l, sp, tin, tw, tarea, rh, rw, lane, allow_tiled,
                                [&](int pos, int, int, int cog, float4 a) { tout4[cog * rarea + pos] = a; });
                    __syncwarp();
                    for (int pos = lane; pos < rarea; pos += kWarp) {
                        const int y = pos / rw, x = pos - y * rw;
                        const int site = (ry + y) * Lx + rx + x;
                        float re, im;
                        site_factor<true>(m, sp, tout, rarea, pos, (float)spins_s[site], re, im);
                        cache[m.fre_off + site] = re;
                        cache[m.fim_off + site] = im;
                        if (factors) factors[(size_t)s * n + site] = make_float2(re, im);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
        }
        if (logpsi) {   // deterministic site sum
            float re = 0.f, im = 0.f;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                re += __ldcg(cache + m.fre_off + i);
                im += __ldcg(cache + m.fim_off + i);
            }
            red[threadIdx.x] = re;
            red[blockDim.x + threadIdx.x] = im;
            __syncthreads();
            for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
                if (threadIdx.x < o) {
                    red[threadIdx.x] += red[threadIdx.x + o];
                    red[blockDim.x + threadIdx.x] += red[blockDim.x + threadIdx.x + o];
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) logpsi[s] = make_float2(red[0], red[blockDim.x]);
        }
        __syncthreads();
    }
}

cudaError_t launch_forward(const qmc_handle* h, const int8_t* spins, int N, float* cache,
                           float* factors, float* logpsi, cudaStream_t st, std::string& err) {
    return launch_forward_images(h, 1, h->d_params_padded, spins, N, cache, factors, logpsi, st, err);
}

// the forward of `nimg` parameter images (qmc_set_image_params) of the same samples in ONE launch:
// caches [nimg, N, cache_floats], factors [nimg, N, n], logpsi [nimg, N]
cudaError_t launch_forward_images(const qmc_handle* h, int nimg, const float* padded_blocks, const int8_t* spins, int N,
                                  float* cache, float* factors, float* logpsi, cudaStream_t st, std::string& err) {
    if (h->d_plane_tab && forward_plane_supported(h))          // the lattice's planes fit in shared memory
        return launch_forward_plane(h, nimg, padded_blocks, spins, N, cache, factors, logpsi, st);
    const DevModel& m = h->m;
    const int p = m.p, side = kFwdBlock + 2 * p;
    int bin = side * side, bout = 0;
    for (int l = 0; l < m.D; ++l) {
        const int t = side * side * m.layer[l].cinp;
        bin = bin > t ? bin : t;
    }
    bout = kFwdBlock * kFwdBlock * m.layer[m.D - 1].coutp;
    bin = round4(bin); bout = round4(bout);
    const size_t per_warp = (size_t)(bin + bout) * 4;
    const int nblocks = ((m.Ly + kFwdBlock - 1) / kFwdBlock) * ((m.Lx + kFwdBlock - 1) / kFwdBlock);
    int warps = nblocks < 8 ? nblocks : 8;
    // power-of-two thread count for the tree reduction
    int w2 = 1; while (w2 * 2 <= warps) w2 *= 2; warps = w2;
    size_t smem = 0;
    for (;; warps >>= 1) {
        smem = (size_t)m.smem_param_floats * 4 + per_warp * warps + (size_t)2 * warps * 32 * 4 + ((m.n + 15) & ~15);
        if (smem <= h->max_smem) break;
        if (warps == 1) { err = "forward: model does not fit in shared memory"; return cudaErrorInvalidValue; }
    }
    cudaError_t e = cudaFuncSetAttribute(k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_img = (h->num_sms * 4 + nimg - 1) / nimg;
    int grid = N < per_img ? N : per_img;
    ++g_launches;
    const ImageStrides is{(size_t)m.smem_param_floats, (size_t)N * m.cache_floats};
    k_forward<<<dim3(grid, nimg), warps * 32, smem, st>>>(m, padded_blocks, spins, N, cache,
                                             reinterpret_cast<float2*>(factors),
                                             reinterpret_cast<float2*>(logpsi), bin, bout, h->allow_tiled ? 1 : 0, is);
    return cudaGetLastError();
}

// log psi_g - log psi_0 of every image (double; per-site factor differences first, so that the ~1e2-sized totals
// never meet in fp32) and log psi_sym = log psi_0 + log mean_g exp(log psi_g - log psi_0), from the images' caches
__global__ void k_sym_logrel(DevModel m, int nsym, int N, const float* __restrict__ caches, double* __restrict__ log_rel,
                             float2* __restrict__ logpsi_sym) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const size_t cimg = (size_t)N * m.cache_floats;
    const float* c0 = caches + (size_t)s * m.cache_floats;
    double l0r = 0, l0i = 0, sr = 0, si = 0;
    if (logpsi_sym)
        for (int i = 0; i < m.n; ++i) { l0r += (double)c0[m.fre_off + i]; l0i += (double)c0[m.fim_off + i]; }
    for (int g = 0; g < nsym; ++g) {
        const float* cg = c0 + g * cimg;
        double dre = 0, dim = 0;
        if (g > 0)
            for (int i = 0; i < m.n; ++i) {
                dre += (double)cg[m.fre_off + i] - (double)c0[m.fre_off + i];
                dim += (double)cg[m.fim_off + i] - (double)c0[m.fim_off + i];
            }
        if (log_rel) { log_rel[((size_t)s * nsym + g) * 2] = dre; log_rel[((size_t)s * nsym + g) * 2 + 1] = dim; }
        const double a = exp(dre);
        sr += a * cos(dim); si += a * sin(dim);
    }
    if (logpsi_sym) {
        sr /= nsym; si /= nsym;
        logpsi_sym[s] = make_float2((float)(l0r + 0.5 * log(sr * sr + si * si)), (float)(l0i + atan2(si, sr)));
    }
}

cudaError_t launch_sym_logrel(const qmc_handle* h, int nsym, int N, const float* caches, double* log_rel,
                              float* logpsi_sym, cudaStream_t st) {
    ++g_launches;
    k_sym_logrel<<<(N + 127) / 128, 128, 0, st>>>(h->m, nsym, N, caches, log_rel, reinterpret_cast<float2*>(logpsi_sym));
    return cudaGetLastError();
}

} // namespace qmc

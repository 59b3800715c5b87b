// qmc_backward.cu - K4: hand-written log-psi gradient for the VMC update.
// Replaces TF autodiff of loss_op through model.factors
// (mcmc_tf.py:35-56, 172-177; models.py:51-67, 110-131):
//   grad[p] += sum_n Re[ w_n * conj(d log psi_n / d p) ]
// d f / d theta_c = tanh(theta_c) (complex), so the cotangent entering the last
// layer is  G[c] = Re(w conj t_c),  G[c+half] = Im(w conj t_c); below that it is
// ordinary real backprop through (1 - tanh^2) and the circular cross-correlations.
//
// One CTA walks samples grid-stride.  Activations come from the K1 cache; the
// per-layer cotangent planes ping-pong in an L2-resident per-CTA scratch; the
// parameter gradient is accumulated per CTA in shared memory and the CTAs'
// partial vectors are summed in a fixed order by k_backward_reduce, so the
// result is deterministic for a given launch geometry.
#include "qmc_host.h"

namespace qmc {

// flat caller-order parameters -> padded block layout (the image shared memory holds)
__global__ void k_repack_params(DevModel m, const float* __restrict__ params, float* __restrict__ padded) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.smem_param_floats) return;
    params += (size_t)blockIdx.y * m.P;                        // parameter images (qmc_set_image_params)
    padded += (size_t)blockIdx.y * m.smem_param_floats;
    float v = 0.f;
    for (int l = 0; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        const int rows = m.k * m.k * L.cin;
        if (i >= L.sw_off && i < L.sw_off + rows * L.coutp) {
            const int row = (i - L.sw_off) / L.coutp, co = (i - L.sw_off) - row * L.coutp;
            if (co < L.cout) v = params[L.w_off + row * L.cout + co];
        } else if (i >= L.sb_off && i < L.sb_off + L.cout) {
            v = params[L.b_off + i - L.sb_off];
        }
    }
    if (m.bias_vis_off >= 0 && i >= m.sp_vis_off && i < m.sp_vis_off + 2) v = params[m.bias_vis_off + i - m.sp_vis_off];
    padded[i] = v;
}

// `nimg` flat vectors [nimg, P] -> `nimg` padded blocks [nimg, smem_param_floats] in one launch
cudaError_t repack_params_to(const qmc_handle* h, const float* flat, float* padded, cudaStream_t st, int nimg) {
    const int nthr = 256, nblk = (h->m.smem_param_floats + nthr - 1) / nthr;
    ++g_launches;
    k_repack_params<<<dim3(nblk, nimg), nthr, 0, st>>>(h->m, flat, padded);
    return cudaGetLastError();
}

cudaError_t repack_params(const qmc_handle* h, cudaStream_t st) {
    return repack_params_to(h, h->d_params, h->d_params_padded, st);
}

constexpr int kBwdThreads = 256;

__device__ __forceinline__ float2 ctanh_stable(float a, float b) {
    const float A = fabsf(a), e = expf(-2.f * A);
    float sb, cb;
    sincosf(b, &sb, &cb);
    const float om = 1.f - e;
    const float den = fmaf(om, om, 4.f * e * cb * cb);
    const float re = copysignf((1.f - e * e) / den, a);
    const float im = 4.f * e * sb * cb / den;
    return make_float2(re, im);
}

__global__ void __launch_bounds__(kBwdThreads)
k_backward(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins,
           const float2* __restrict__ weights, int N, const float* __restrict__ cache_all,
           float* __restrict__ gscratch, float* __restrict__ partial, int gplane_floats, ImageStrides is) {
    extern __shared__ float4 smem4[];
    float* sp = reinterpret_cast<float*>(smem4);
    // symmetry images (blockIdx.y): own parameter block, cache, per-sample weights, scratch and partial sums
    params += (size_t)blockIdx.y * is.params;
    cache_all += (size_t)blockIdx.y * is.cache;
    weights += (size_t)blockIdx.y * N;
    gscratch += (size_t)blockIdx.y * gridDim.x * 2 * gplane_floats;
    partial += (size_t)blockIdx.y * gridDim.x * m.P;
    float* acc = sp + m.smem_param_floats;            // P floats, caller's flat order
    load_params_to_smem(m, params, sp);
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int n = m.n, Ly = m.Ly, Lx = m.Lx, k = m.k, p = m.p, D = m.D;
    float* G0 = gscratch + (size_t)blockIdx.x * 2 * gplane_floats;   // [site][C] channel-last
    float* G1 = G0 + gplane_floats;

    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        const float* cache = cache_all + (size_t)s * m.cache_floats;
        const int8_t* sx = spins + (size_t)s * n;
        const float2 w = weights[s];
        // ---- head: theta of the last layer, cotangent G_D -----------------------
        {
            const LayerInfo& L = m.layer[D - 1];
            const int half = L.cout >> 1;
            const float* inplane = D > 1 ? cache + m.layer[D - 2].act_off : nullptr;
            for (int t = threadIdx.x; t < n * half; t += blockDim.x) {
                const int site = t / half, c = t - site * half;
                const int y = site / Lx, x = site - y * Lx;
                float th[2];
                for (int part = 0; part < 2; ++part) {
                    const int co = c + part * half;
                    float a = sp[L.sb_off + co];
                    for (int dy = 0; dy < k; ++dy)
                        for (int dx = 0; dx < k; ++dx) {
                            const int q = wrapi(y + dy - p, Ly) * Lx + wrapi(x + dx - p, Lx);
                            const float* wr = sp + L.sw_off + (dy * k + dx) * L.cin * L.coutp + co;
                            if (D == 1) {
                                a = fmaf((float)sx[q], wr[0], a);
                            } else {
                                for (int ci = 0; ci < L.cin; ++ci)
                                    a = fmaf(__ldcg(inplane + (size_t)((ci >> 2) * n + q) * 4 + (ci & 3)),
                                             wr[ci * L.coutp], a);
                            }
                        }
                    th[part] = a;
                }
                const float2 tc = ctanh_stable(th[0], th[1]);
                // w * conj(t)
                G0[site * L.cout + c] = w.x * tc.x + w.y * tc.y;
                G0[site * L.cout + c + half] = w.y * tc.x - w.x * tc.y;
            }
            if (m.bias_vis_off >= 0 && threadIdx.x < 32) {     // visible bias: Re(w) sum s, Im(w) sum s
                int ssum = 0;
                for (int i = threadIdx.x; i < n; i += 32) ssum += sx[i];
                for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if (threadIdx.x == 0) {
                    acc[m.bias_vis_off] += w.x * (float)ssum;
                    acc[m.bias_vis_off + 1] += w.y * (float)ssum;
                }
            }
        }
        __syncthreads();
        float* G = G0;
        float* Gn = G1;
        for (int l = D - 1; l >= 0; --l) {
            const LayerInfo& L = m.layer[l];
            const float* inplane = l > 0 ? cache + m.layer[l - 1].act_off : nullptr;
            // ---- parameter gradients of layer l ---------------------------------
            const int ncog = (L.cout + 3) >> 2;
            const int ntask = k * k * L.cin * ncog;
            for (int t = threadIdx.x; t < ntask + L.cout; t += blockDim.x) {
                if (t >= ntask) {                                  // bias
                    const int co = t - ntask;
                    float sum = 0.f;
                    for (int site = 0; site < n; ++site) sum += __ldcg(G + site * L.cout + co);
                    acc[L.b_off + co] += sum;
                    continue;
                }
                const int cog = t % ncog, rest = t / ncog;
                const int ci = rest % L.cin, d = rest / L.cin;
                const int dy = d / k - p, dx = d % k - p;
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                const int nco = min(4, L.cout - cog * 4);
                for (int y = 0; y < Ly; ++y) {
                    const int qy = wrapi(y + dy, Ly) * Lx;
                    for (int x = 0; x < Lx; ++x) {
                        const int q = qy + wrapi(x + dx, Lx);
                        const float a = l == 0 ? (float)sx[q]
                                               : __ldcg(inplane + (size_t)((ci >> 2) * n + q) * 4 + (ci & 3));
                        const float* g = G + (y * Lx + x) * L.cout + cog * 4;
                        for (int j = 0; j < nco; ++j) s4[j] = fmaf(a, __ldcg(g + j), s4[j]);
                    }
                }
                float* dst = acc + L.w_off + (d * L.cin + ci) * L.cout + cog * 4;
                for (int j = 0; j < nco; ++j) dst[j] += s4[j];
            }
            // ---- cotangent of the layer below -----------------------------------
            if (l > 0) {
                for (int t = threadIdx.x; t < n * L.cin; t += blockDim.x) {
                    const int site = t / L.cin, ci = t - site * L.cin;
                    const int y = site / Lx, x = site - y * Lx;
                    float sum = 0.f;
                    for (int dy = 0; dy < k; ++dy)
                        for (int dx = 0; dx < k; ++dx) {
                            // output site q = site - (d - p) reads input site through filter tap d
                            const int q = wrapi(y - dy + p, Ly) * Lx + wrapi(x - dx + p, Lx);
                            const float* wr = sp + L.sw_off + ((dy * k + dx) * L.cin + ci) * L.coutp;
                            const float* g = G + q * L.cout;
                            for (int co = 0; co < L.cout; ++co) sum = fmaf(wr[co], __ldcg(g + co), sum);
                        }
                    const float a = __ldcg(inplane + (size_t)((ci >> 2) * n + site) * 4 + (ci & 3));
                    Gn[site * L.cin + ci] = (1.f - a * a) * sum;
                }
            }
            __syncthreads();
            float* tmp = G; G = Gn; Gn = tmp;
        }
    }
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) partial[(size_t)blockIdx.x * m.P + i] = acc[i];
}


// ---------------------------------------------------------------------------------------------
// k_backward_smem: the same computation with everything a layer touches staged in shared memory.
// One CTA per SM walks samples grid-stride; per sample and layer the input activation plane
// (from the K1 cache), the cotangent plane G and the cotangent below Gn live in shared memory
// next to the padded parameter block and the CTA's gradient accumulator, so the 9 x C_in x C_out
// re-reads of every site hit shared memory instead of L2 (the first version ran at <1% of the
// FP32 roofline).  Planes are channel-group planar float4 [cg][site] like the cache: consecutive
// threads (consecutive sites) read consecutive 16-byte words.  Every sum keeps the order of
// k_backward (taps, then channels, then sites ascending), one owner thread per gradient element:
// deterministic.  Used when the planes fit (lattices up to ~24 x 24 at 16 channels); larger
// lattices run k_backward.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdSmemThreads = 576;      // = 9 taps x 16 C_in x 4 channel groups: one weight-gradient round

__global__ void __launch_bounds__(kBwdSmemThreads, 1)
k_backward_smem(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins,
                const float2* __restrict__ weights, int N, const float* __restrict__ cache_all,
                float* __restrict__ partial, int plane_floats, ImageStrides is) {
    extern __shared__ float4 smem4[];
    float* sp = reinterpret_cast<float*>(smem4);
    params += (size_t)blockIdx.y * is.params;                  // symmetry images, as in k_backward
    cache_all += (size_t)blockIdx.y * is.cache;
    weights += (size_t)blockIdx.y * N;
    partial += (size_t)blockIdx.y * gridDim.x * m.P;
    float* acc = sp + m.smem_param_floats;                  // P floats, caller's flat order
    float* A = acc + round_up4(m.P);                        // input plane of the current layer
    float* G = A + plane_floats;                            // cotangent of the current layer's output
    float* Gn = G + plane_floats;                           // cotangent of its input
    load_params_to_smem(m, params, sp);
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int n = m.n, Ly = m.Ly, Lx = m.Lx, k = m.k, p = m.p, D = m.D;
    const int tid = threadIdx.x, nthr = blockDim.x;

    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        const float* cache = cache_all + (size_t)s * m.cache_floats;
        const int8_t* sx = spins + (size_t)s * n;
        const float2 w = weights[s];
        // input plane of the last layer
        auto stage_input = [&](int l) {
            if (l == 0) {
                for (int i = tid; i < n; i += nthr) A[i] = (float)sx[i];
            } else {
                const LayerInfo& Lp = m.layer[l - 1];
                const float4* src = reinterpret_cast<const float4*>(cache + Lp.act_off);
                float4* dst = reinterpret_cast<float4*>(A);
                for (int i = tid; i < (Lp.coutp >> 2) * n; i += nthr) dst[i] = __ldcg(src + i);
            }
        };
        stage_input(D - 1);
        __syncthreads();
        // ---- head: theta of the last layer, cotangent G_D = (Re, Im)(w conj tanh theta) ----------
        {
            const LayerInfo& L = m.layer[D - 1];
            const int half = L.cout >> 1;
            for (int i = tid; i < (L.coutp >> 2) * n * 4; i += nthr) G[i] = 0.f;      // padded channels stay zero
            __syncthreads();
            for (int t = tid; t < n * half; t += nthr) {
                const int c = t / n, site = t - c * n;            // consecutive threads: consecutive sites
                const int y = site / Lx, x = site - y * Lx;
                float th[2];
                for (int part = 0; part < 2; ++part) {
                    const int co = c + part * half;
                    float a = sp[L.sb_off + co];
                    for (int dy = 0; dy < k; ++dy) {
                        const int qy = wrap1(y + dy - p, Ly) * Lx;
                        for (int dx = 0; dx < k; ++dx) {
                            const int q = qy + wrap1(x + dx - p, Lx);
                            const float* wr = sp + L.sw_off + (dy * k + dx) * L.cin * L.coutp + co;
                            if (D == 1) {
                                a = fmaf(A[q], wr[0], a);
                            } else {
                                for (int ci = 0; ci < L.cin; ++ci)
                                    a = fmaf(A[((ci >> 2) * n + q) * 4 + (ci & 3)], wr[ci * L.coutp], a);
                            }
                        }
                    }
                    th[part] = a;
                }
                const float2 tc = ctanh_stable(th[0], th[1]);
                const int c2 = c + half;
                G[((c >> 2) * n + site) * 4 + (c & 3)] = w.x * tc.x + w.y * tc.y;       // w * conj(t)
                G[((c2 >> 2) * n + site) * 4 + (c2 & 3)] = w.y * tc.x - w.x * tc.y;
            }
            if (m.bias_vis_off >= 0 && tid < 32) {                 // visible bias: Re(w) sum s, Im(w) sum s
                int ssum = 0;
                for (int i = tid; i < n; i += 32) ssum += sx[i];
                for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if (tid == 0) {
                    acc[m.bias_vis_off] += w.x * (float)ssum;
                    acc[m.bias_vis_off + 1] += w.y * (float)ssum;
                }
            }
        }
        __syncthreads();
        for (int l = D - 1; l >= 0; --l) {
            const LayerInfo& L = m.layer[l];
            const int ncog = L.coutp >> 2;
            const float4* G4 = reinterpret_cast<const float4*>(G);
            // ---- parameter gradients of layer l: one thread per (tap, C_in, 4 C_out) ----------------
            const int ntask = k * k * L.cin * ncog;
            for (int t = tid; t < ntask + L.cout; t += nthr) {
                if (t >= ntask) {                                  // bias
                    const int co = t - ntask;
                    float sum = 0.f;
                    for (int site = 0; site < n; ++site) sum += G[((co >> 2) * n + site) * 4 + (co & 3)];
                    acc[L.b_off + co] += sum;
                    continue;
                }
                const int cog = t % ncog, rest = t / ncog;
                const int ci = rest % L.cin, d = rest / L.cin;
                const int dy = d / k - p, dx = d % k - p;
                const float* Ap = l == 0 ? A : A + (size_t)(ci >> 2) * n * 4 + (ci & 3);
                const int astride = l == 0 ? 1 : 4;
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4* g = G4 + cog * n;
                for (int y = 0; y < Ly; ++y) {
                    const int qy = wrap1(y + dy, Ly) * Lx;
                    for (int x = 0; x < Lx; ++x) {
                        const float a = Ap[(qy + wrap1(x + dx, Lx)) * astride];
                        const float4 gv = g[y * Lx + x];
                        s4.x = fmaf(a, gv.x, s4.x); s4.y = fmaf(a, gv.y, s4.y);
                        s4.z = fmaf(a, gv.z, s4.z); s4.w = fmaf(a, gv.w, s4.w);
                    }
                }
                float* dst = acc + L.w_off + (d * L.cin + ci) * L.cout + cog * 4;
                const int nco = min(4, L.cout - cog * 4);
                if (nco > 0) dst[0] += s4.x;
                if (nco > 1) dst[1] += s4.y;
                if (nco > 2) dst[2] += s4.z;
                if (nco > 3) dst[3] += s4.w;
            }
            // ---- cotangent of the layer below: one thread per (4 C_in, site) --------------------------
            if (l > 0) {
                const int ncig = L.cinp >> 2;
                float4* Gn4 = reinterpret_cast<float4*>(Gn);
                const float4* A4 = reinterpret_cast<const float4*>(A);
                for (int t = tid; t < ncig * n; t += nthr) {
                    const int cig = t / n, site = t - cig * n;
                    const int y = site / Lx, x = site - y * Lx;
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int dy = 0; dy < k; ++dy) {
                        const int qy = wrap1(y - dy + p, Ly) * Lx;
                        for (int dx = 0; dx < k; ++dx) {
                            // output site q = site - (d - p) reads this input site through filter tap d
                            const int q = qy + wrap1(x - dx + p, Lx);
                            const float* wr = sp + L.sw_off + ((dy * k + dx) * L.cin + cig * 4) * L.coutp;
                            for (int cg = 0; cg < ncog; ++cg) {
                                const float4 gv = G4[cg * n + q];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (cig * 4 + j >= L.cin) break;
                                    const float4 wv = *reinterpret_cast<const float4*>(wr + j * L.coutp + cg * 4);
                                    sum[j] = fmaf(wv.x, gv.x, sum[j]); sum[j] = fmaf(wv.y, gv.y, sum[j]);
                                    sum[j] = fmaf(wv.z, gv.z, sum[j]); sum[j] = fmaf(wv.w, gv.w, sum[j]);
                                }
                            }
                        }
                    }
                    const float4 a = A4[cig * n + site];
                    Gn4[cig * n + site] = make_float4((1.f - a.x * a.x) * sum[0], (1.f - a.y * a.y) * sum[1],
                                                      (1.f - a.z * a.z) * sum[2], (1.f - a.w * a.w) * sum[3]);
                }
            }
            __syncthreads();
            if (l > 0) {
                stage_input(l - 1);           // A is free: every thread is past the reads of this layer
                float* tmp = G; G = Gn; Gn = tmp;
                __syncthreads();
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) partial[(size_t)blockIdx.x * m.P + i] = acc[i];
}

__global__ void k_backward_reduce(const float* __restrict__ partial, int nparts, int P,
                                  float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    partial += (size_t)blockIdx.y * nparts * P;                // image blockIdx.y -> grad[image][P]
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * P + i];
    grad[(size_t)blockIdx.y * P + i] += s;
}

// per-image sample weights of the symmetrised gradient: d/dp sum_n Re[w_n conj(log psi_sym,n)] =
// sum_g sum_n Re[(w_n conj(p_gn)) conj(d log psi_gn / dp)], p_g = psi_g / sum_h psi_h from the caches' per-site factor
// planes (differences to image 0 first, in double - k_energy_finish_sym's weights)
__global__ void k_sym_weights(DevModel m, int nsym, int N, const float* __restrict__ caches,
                              const float2* __restrict__ w, float2* __restrict__ wimg) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const size_t cimg = (size_t)N * m.cache_floats;
    const float* c0 = caches + (size_t)s * m.cache_floats;
    double wr[8], wi[8], sr = 0, si = 0;
    for (int g = 0; g < nsym; ++g) {
        const float* cg = c0 + g * cimg;
        double dre = 0, dim = 0;
        if (g > 0)
            for (int i = 0; i < m.n; ++i) {
                dre += (double)cg[m.fre_off + i] - (double)c0[m.fre_off + i];
                dim += (double)cg[m.fim_off + i] - (double)c0[m.fim_off + i];
            }
        const double a = exp(dre);
        wr[g] = a * cos(dim); wi[g] = a * sin(dim);
        sr += wr[g]; si += wi[g];
    }
    const double den = sr * sr + si * si;
    const float2 ws = w[s];
    for (int g = 0; g < nsym; ++g) {
        const double pr = (wr[g] * sr + wi[g] * si) / den, pi = (wi[g] * sr - wr[g] * si) / den;   // p_g
        // w * conj(p_g)
        wimg[(size_t)g * N + s] = make_float2((float)(ws.x * pr + ws.y * pi), (float)(ws.y * pr - ws.x * pi));
    }
}

static int gplane(const DevModel& m);

// shared memory of k_backward_smem, or 0 when the planes do not fit / QMC_FLAG_BACKWARD_GENERIC
static size_t backward_smem_bytes(const qmc_handle* h) {
    if (h->backward_generic) return 0;                     // force the L2-resident k_backward (cross-check)
    const DevModel& m = h->m;
    int c = 1;
    for (int l = 0; l < m.D; ++l) c = c > m.layer[l].coutp ? c : m.layer[l].coutp;
    const size_t bytes = (size_t)(m.smem_param_floats + round4(m.P) + 3 * round4(m.n * c)) * 4;
    return bytes <= h->max_smem ? bytes : 0;
}

static int backward_ctas(const qmc_handle* h, int N) {
    const int cap = backward_smem_bytes(h) ? h->num_sms : h->num_sms * 2;
    return N < cap ? (N > 0 ? N : 1) : cap;
}

static int gplane(const DevModel& m) {
    int c = 1;
    for (int l = 0; l < m.D; ++l) c = c > m.layer[l].cout ? c : m.layer[l].cout;
    return round4(m.n * c);
}

size_t backward_workspace_floats(const qmc_handle* h, int N) {
    return backward_images_workspace_floats(h, 1, N);
}

static int backward_image_ctas(const qmc_handle* h, int nimg, int N) {
    int c = backward_ctas(h, N);
    if (nimg > 1) { c = c / nimg; if (c < 1) c = 1; }          // the images share the SMs
    return c;
}

size_t backward_images_workspace_floats(const qmc_handle* h, int nimg, int N) {
    if (h->d_bwd_tab && backward_plane_supported(h)) return backward_plane_workspace_floats(h, nimg, N);
    const DevModel& m = h->m;
    const size_t ctas = (size_t)backward_image_ctas(h, nimg, N) * nimg;
    return (size_t)nimg * N * m.cache_floats + ctas * 2 * gplane(m) + ctas * round4(m.P) + (nimg > 1 ? (size_t)nimg * N * 2 : 0);
}

// gradient for `nimg` parameter images in one launch sequence: grad [nimg, P] (+=).  nimg == 1: weights [N] are the
// per-sample cotangents; nimg > 1 (symmetry average): weights [N] are those of log psi_sym and the per-image ones
// are formed by k_sym_weights
cudaError_t launch_backward_images(const qmc_handle* h, int nimg, const float* blocks, const int8_t* spins,
                                   const float* weights, int N, float* workspace, float* grad, cudaStream_t st,
                                   std::string& err) {
    const DevModel& m = h->m;
    float* cache = workspace;
    if (h->d_bwd_tab && backward_plane_supported(h)) {
        // per-layer band kernels: [caches | 2 x cotangent planes | CTA partials | image weights]
        if (nimg > 8) { err = "backward: at most 8 images"; return cudaErrorInvalidValue; }
        const int pc = backward_plane_ctas(h, nimg, N);
        const size_t gfl = (backward_plane_workspace_floats(h, nimg, N) - (size_t)nimg * N * m.cache_floats -
                            (size_t)pc * nimg * round4(m.P) - (nimg > 1 ? (size_t)nimg * N * 2 : 0));
        float* gbuf = workspace + (size_t)nimg * N * m.cache_floats;
        float* part = gbuf + gfl;
        float* wim = part + (size_t)pc * nimg * round4(m.P);
        cudaError_t e2 = launch_forward_images(h, nimg, blocks, spins, N, cache, nullptr, nullptr, st, err);
        if (e2 != cudaSuccess) return e2;
        const float2* wv = reinterpret_cast<const float2*>(weights);
        if (nimg > 1) {
            ++g_launches;
            k_sym_weights<<<(N + 127) / 128, 128, 0, st>>>(m, nimg, N, cache, wv, reinterpret_cast<float2*>(wim));
            if ((e2 = cudaGetLastError()) != cudaSuccess) return e2;
            wv = reinterpret_cast<const float2*>(wim);
        }
        e2 = launch_backward_plane(h, nimg, blocks, spins, wv, N, cache, gbuf, part, grad, st);
        if (e2 != cudaSuccess) return e2;
        ++g_launches;
        k_backward_reduce<<<dim3((m.P + 127) / 128, nimg), 128, 0, st>>>(part, pc, m.P, grad);
        return cudaGetLastError();
    }
    const int ctas = backward_image_ctas(h, nimg, N);
    float* gscratch = workspace + (size_t)nimg * N * m.cache_floats;
    float* partial = gscratch + (size_t)ctas * nimg * 2 * gplane(m);
    float* wimg = partial + (size_t)ctas * nimg * round4(m.P);
    cudaError_t e = launch_forward_images(h, nimg, blocks, spins, N, cache, nullptr, nullptr, st, err);
    if (e != cudaSuccess) return e;
    const float2* w2 = reinterpret_cast<const float2*>(weights);
    if (nimg > 1) {
        if (nimg > 8) { err = "backward: at most 8 images"; return cudaErrorInvalidValue; }
        ++g_launches;
        k_sym_weights<<<(N + 127) / 128, 128, 0, st>>>(m, nimg, N, cache, w2, reinterpret_cast<float2*>(wimg));
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        w2 = reinterpret_cast<const float2*>(wimg);
    }
    const ImageStrides is{(size_t)m.smem_param_floats, (size_t)N * m.cache_floats};
    if (const size_t sb = backward_smem_bytes(h)) {
        int c = 1;
        for (int l = 0; l < m.D; ++l) c = c > m.layer[l].coutp ? c : m.layer[l].coutp;
        e = cudaFuncSetAttribute(k_backward_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
        if (e != cudaSuccess) return e;
        g_launches += 2;
        k_backward_smem<<<dim3(ctas, nimg), kBwdSmemThreads, sb, st>>>(m, blocks, spins, w2, N, cache, partial,
                                                                     round4(m.n * c), is);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        k_backward_reduce<<<dim3((m.P + 127) / 128, nimg), 128, 0, st>>>(partial, ctas, m.P, grad);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)(m.smem_param_floats + round4(m.P)) * 4;
    if (smem > h->max_smem) { err = "backward: parameters do not fit in shared memory"; return cudaErrorInvalidValue; }
    e = cudaFuncSetAttribute(k_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    g_launches += 2;
    k_backward<<<dim3(ctas, nimg), kBwdThreads, smem, st>>>(m, blocks, spins, w2, N, cache, gscratch, partial, gplane(m), is);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_backward_reduce<<<dim3((m.P + 127) / 128, nimg), 128, 0, st>>>(partial, ctas, m.P, grad);
    return cudaGetLastError();
}

cudaError_t launch_backward(const qmc_handle* h, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad, cudaStream_t st, std::string& err) {
    return launch_backward_images(h, 1, h->d_params_padded, spins, weights, N, workspace, grad, st, err);
}

} // namespace qmc

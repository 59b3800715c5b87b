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

// flat caller-order parameters -> padded block layout (the image c_params / shared memory hold)
__global__ void k_repack_params(DevModel m, const float* __restrict__ params, float* __restrict__ padded) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.smem_param_floats) return;
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

cudaError_t repack_params_to(const qmc_handle* h, const float* flat, float* padded, cudaStream_t st) {
    const int nthr = 256, nblk = (h->m.smem_param_floats + nthr - 1) / nthr;
    ++g_launches;
    k_repack_params<<<nblk, nthr, 0, st>>>(h->m, flat, padded);
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
           float* __restrict__ gscratch, float* __restrict__ partial, int gplane_floats) {
    extern __shared__ float4 smem4[];
    float* sp = reinterpret_cast<float*>(smem4);
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

__global__ void k_backward_reduce(const float* __restrict__ partial, int nparts, int P,
                                  float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * P + i];
    grad[i] += s;
}

static int backward_ctas(const qmc_handle* h, int N) {
    const int cap = h->num_sms * 2;
    return N < cap ? (N > 0 ? N : 1) : cap;
}

static int gplane(const DevModel& m) {
    int c = 1;
    for (int l = 0; l < m.D; ++l) c = c > m.layer[l].cout ? c : m.layer[l].cout;
    return round4(m.n * c);
}

size_t backward_workspace_floats(const qmc_handle* h, int N) {
    const DevModel& m = h->m;
    const size_t ctas = backward_ctas(h, N);
    return (size_t)N * m.cache_floats + ctas * 2 * gplane(m) + ctas * round4(m.P);
}

cudaError_t launch_backward(const qmc_handle* h, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad, cudaStream_t st, std::string& err) {
    const DevModel& m = h->m;
    float* cache = workspace;
    const int ctas = backward_ctas(h, N);
    float* gscratch = workspace + (size_t)N * m.cache_floats;
    float* partial = gscratch + (size_t)ctas * 2 * gplane(m);
    cudaError_t e = launch_forward(h, spins, N, cache, nullptr, nullptr, st, err);
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)(m.smem_param_floats + round4(m.P)) * 4;
    if (smem > h->max_smem) { err = "backward: parameters do not fit in shared memory"; return cudaErrorInvalidValue; }
    e = cudaFuncSetAttribute(k_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    g_launches += 2;
    k_backward<<<ctas, kBwdThreads, smem, st>>>(m, h->d_params, spins,
                                               reinterpret_cast<const float2*>(weights), N, cache,
                                               gscratch, partial, gplane(m));
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_backward_reduce<<<(m.P + 127) / 128, 128, 0, st>>>(partial, ctas, m.P, grad);
    return cudaGetLastError();
}

} // namespace qmc

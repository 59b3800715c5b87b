// qmc_nd.cu - 1-D and 3-D lattices (the conv1d / conv3d branches of models.py:56-61, 118-123 and the
// n_dims-generic sampler.py / mcmc_tf.py).  A separate, deliberately simple code path: one CTA per chain
// or sample runs the REFERENCE'S OWN ALGORITHM - a full network evaluation of the flipped configuration per
// proposal (sampler.py:117-133) and per connected configuration (mcmc_tf.py:72-90, 106-141) - with every
// operand in L2.  It restores API coverage for n_dims = 1, 3 at parity with the oracle; the tuned
// incremental kernels are 2-D.  Same arithmetic contract as the 2-D path: acc = bias, then taps ascending
// (row-major over the lattice axes), then input channels ascending, fmaf; tanhf between layers;
// log2cosh_c at the head.  Stateless C entry points (no handle): parameters are the caller's flat vector
// in the reference's variable order.
#include <string>
#include "qmc_host.h"

namespace qmc {

struct NdModel {
    int kind, k, D, nd, n, P;
    int L[3], str[3];                  // lattice sides (leading axes padded with 1) and site strides
    int cin[QMC_MAX_LAYERS], cout[QMC_MAX_LAYERS], w_off[QMC_MAX_LAYERS], b_off[QMC_MAX_LAYERS];
    int bias_vis_off, cmax, ktaps;
};

static bool nd_build(const qmc_nd_desc* d, NdModel& m, std::string& err) {
    if (!d) { err = "null descriptor"; return false; }
    if (d->n_dims < 1 || d->n_dims > 3) { err = "n_dims must be 1, 2 or 3"; return false; }
    if (d->kind != QMC_MODEL_CRBM && d->kind != QMC_MODEL_DCRBM) { err = "unknown model kind"; return false; }
    if (d->k < 1 || d->k % 2 == 0) { err = "filter side k must be odd"; return false; }
    if (d->n_layers < 1 || d->n_layers > QMC_MAX_LAYERS) { err = "n_layers out of range"; return false; }
    if (d->kind == QMC_MODEL_CRBM && d->n_layers != 1) { err = "CRBM has exactly one layer"; return false; }
    if (d->channels[d->n_layers - 1] % 2) { err = "last layer needs an even channel count"; return false; }
    m.kind = d->kind; m.k = d->k; m.D = d->n_layers; m.nd = d->n_dims;
    for (int a = 0; a < 3; ++a) m.L[a] = 1;
    for (int a = 0; a < m.nd; ++a) {
        if (d->L[a] < 1) { err = "lattice sides must be positive"; return false; }
        m.L[3 - m.nd + a] = d->L[a];
    }
    m.str[2] = 1; m.str[1] = m.L[2]; m.str[0] = m.L[1] * m.L[2];
    m.n = m.L[0] * m.L[1] * m.L[2];
    m.ktaps = 1;
    for (int a = 0; a < m.nd; ++a) m.ktaps *= m.k;
    int off = 0, cin = 1;
    m.bias_vis_off = -1; m.cmax = 1;
    for (int l = 0; l < m.D; ++l) {
        if (d->channels[l] < 1) { err = "channel counts must be positive"; return false; }
        m.cin[l] = cin; m.cout[l] = d->channels[l];
        m.w_off[l] = off; off += m.ktaps * cin * m.cout[l];
        if (d->kind == QMC_MODEL_CRBM) { m.bias_vis_off = off; off += 2; }      // models.py:19-28 order
        m.b_off[l] = off; off += m.cout[l];
        cin = m.cout[l];
        if (m.cout[l] > m.cmax) m.cmax = m.cout[l];
    }
    m.P = off;
    return true;
}

// flat site index of (site + tap - p) on the periodic lattice; tap is row-major over the nd trailing axes
__device__ __forceinline__ int nd_neighbour(const NdModel& m, int site, int tap) {
    const int p = (m.k - 1) >> 1;
    int q = 0, s = site, t = tap;
#pragma unroll
    for (int a = 2; a >= 0; --a) {
        const int La = m.L[a];
        const int c = s % La; s /= La;
        int d = 0;
        if (a >= 3 - m.nd) { d = t % m.k - p; t /= m.k; }
        q += wrapi(c + d, La) * m.str[a];
    }
    return q;
}

// Full network on one configuration by one CTA: spins (+-1, with up to two sites negated) -> per-site complex
// factors fac[n] (float2).  act0/act1: ping-pong activation planes [n][C] in global memory.
__device__ void nd_forward_cta(const NdModel& m, const float* __restrict__ params, const int8_t* spins, int f0,
                               int f1, float* act0, float* act1, float2* fac) {
    const int n = m.n;
    float* in = act0;
    float* out = act1;
    for (int l = 0; l < m.D; ++l) {
        const int cin = m.cin[l], cout = m.cout[l];
        const bool last = l == m.D - 1;
        const float* w = params + m.w_off[l];
        const float* b = params + m.b_off[l];
        for (int idx = threadIdx.x; idx < n * cout; idx += blockDim.x) {
            const int site = idx / cout, co = idx - site * cout;
            float acc = __ldg(b + co);
            for (int tap = 0; tap < m.ktaps; ++tap) {
                const int q = nd_neighbour(m, site, tap);
                const float* wr = w + (size_t)tap * cin * cout + co;
                if (l == 0) {
                    int s = spins[q];
                    if (q == f0) s = -s;
                    if (q == f1) s = -s;
                    acc = fmaf((float)s, __ldg(wr), acc);
                } else {
                    for (int ci = 0; ci < cin; ++ci) acc = fmaf(in[(size_t)q * cin + ci], __ldg(wr + ci * cout), acc);
                }
            }
            out[(size_t)site * cout + co] = last ? acc : tanhf(acc);
        }
        __syncthreads();
        float* t = in; in = out; out = t;
    }
    // head: sum_c log 2cosh(theta_c + i theta_{c+half}) (+ visible bias x spin for CRBM), models.py:64-67, 128-131
    const int C = m.cout[m.D - 1], half = C >> 1;
    for (int site = threadIdx.x; site < n; site += blockDim.x) {
        float re = 0.f, im = 0.f;
        for (int c = 0; c < half; ++c) {
            float r1, i1;
            log2cosh_c<true>(in[(size_t)site * C + c], in[(size_t)site * C + c + half], r1, i1);
            re += r1; im += i1;
        }
        if (m.bias_vis_off >= 0) {
            int s = spins[site];
            if (site == f0) s = -s;
            if (site == f1) s = -s;
            re = fmaf(__ldg(params + m.bias_vis_off), (float)s, re);
            im = fmaf(__ldg(params + m.bias_vis_off + 1), (float)s, im);
        }
        fac[site] = make_float2(re, im);
    }
    __syncthreads();
}

// deterministic block sum of per-site complex values (site order fixed per thread, tree over threads)
__device__ float2 nd_block_sum(const float2* a, const float2* b, int n, float2* red) {
    float re = 0.f, im = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        re += a[i].x - (b ? b[i].x : 0.f);
        im += a[i].y - (b ? b[i].y : 0.f);
    }
    red[threadIdx.x] = make_float2(re, im);
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) { red[threadIdx.x].x += red[threadIdx.x + o].x; red[threadIdx.x].y += red[threadIdx.x + o].y; }
        __syncthreads();
    }
    const float2 r = red[0];
    __syncthreads();
    return r;
}

constexpr int kNdThreads = 256;

__global__ void __launch_bounds__(kNdThreads)
k_nd_forward(NdModel m, const float* __restrict__ params, const int8_t* __restrict__ spins, int N, float* scratch,
             float2* factors, float2* logpsi) {
    __shared__ float2 red[kNdThreads];
    const size_t per = (size_t)2 * m.n * m.cmax + 2 * (size_t)m.n;
    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        float* base = scratch + (size_t)blockIdx.x * per;
        float2* fac = factors ? factors + (size_t)s * m.n : reinterpret_cast<float2*>(base + (size_t)2 * m.n * m.cmax);
        nd_forward_cta(m, params, spins + (size_t)s * m.n, -1, -1, base, base + (size_t)m.n * m.cmax, fac);
        if (logpsi) {
            const float2 t = nd_block_sum(fac, nullptr, m.n, red);
            if (threadIdx.x == 0) logpsi[s] = t;
        }
    }
}

// Sampler.mcmc_step x n_steps (sampler.py:104-155), one CTA per chain, full forward per proposal
__global__ void __launch_bounds__(kNdThreads)
k_nd_sweep(NdModel m, const float* __restrict__ params, SweepArgs a, float2* cur_fac, float* scratch) {
    __shared__ float2 red[kNdThreads];
    __shared__ int sh_accept;
    const int n = m.n;
    const size_t per = (size_t)2 * n * m.cmax + 2 * (size_t)n;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    for (int chain = blockIdx.x; chain < a.S; chain += gridDim.x) {
        int8_t* spins = a.spins + (size_t)chain * n;
        float2* fac = cur_fac + (size_t)chain * n;
        float* base = scratch + (size_t)blockIdx.x * per;
        float2* nfac = reinterpret_cast<float2*>(base + (size_t)2 * n * m.cmax);
        const unsigned long long gchain = (unsigned long long)(a.chain_id0 + chain);
        unsigned long long accepted = 0;
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
                const uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)((unsigned long long)step >> 32),
                                                         (uint32_t)gchain, (uint32_t)(gchain >> 32)), key);
                f0 = (int)__umulhi(r.x, (uint32_t)n);
                if (a.num_flips > 1) f1 = (int)__umulhi(r.y, (uint32_t)n);
                u = (float)(r.w >> 8) * 5.9604644775390625e-8f;
            }
            bool accept;
            float dre = 0.f;
            if (a.num_flips > 1 && f0 == f1) {
                accept = 1.0f > u;                       // the two flips cancel (sampler.py:114-115)
            } else {
                nd_forward_cta(m, params, spins, f0, f1, base, base + (size_t)n * m.cmax, nfac);
                const float2 d = nd_block_sum(nfac, fac, n, red);            // per-site differences first, :124
                dre = d.x;
                const float amp = expf(d.x);
                if (threadIdx.x == 0) sh_accept = amp * amp > u ? 1 : 0;    // strict, :125
                __syncthreads();
                accept = sh_accept != 0;
                if (accept) {
                    for (int i = threadIdx.x; i < n; i += blockDim.x) fac[i] = nfac[i];
                    if (threadIdx.x == 0) {
                        spins[f0] = -spins[f0];
                        if (a.num_flips > 1) spins[f1] = -spins[f1];
                    }
                }
                __syncthreads();
            }
            if (accept) ++accepted;
            if (threadIdx.x == 0) {
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = dre;
            }
            if (a.samples && step >= a.therm_its && (step - a.therm_its) % a.its_per_sample == 0) {
                const long long j = (step - a.therm_its) / a.its_per_sample;
                if (j < a.n_sample_slots) {
                    int8_t* dst = a.samples + ((size_t)j * a.S + chain) * n;
                    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = spins[i];
                }
                __syncthreads();
            }
        }
        if (a.n_accept && threadIdx.x == 0 && accepted) atomicAdd(a.n_accept, accepted);
    }
}

// ising_energy / heisenberg_energy (mcmc_tf.py:59-141), one CTA per sample, one full forward per connected
// configuration; bonds s_i s_{i+e_d} over the nd lattice axes (helpers.py:171-195)
__global__ void __launch_bounds__(kNdThreads)
k_nd_energy(NdModel m, const float* __restrict__ params, const int8_t* __restrict__ spins_all, int N, int hamiltonian,
            float field_h, float* scratch, float2* e_loc) {
    __shared__ float2 red[kNdThreads];
    const int n = m.n;
    const size_t per = (size_t)2 * n * m.cmax + 4 * (size_t)n;
    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        const int8_t* spins = spins_all + (size_t)s * n;
        float* base = scratch + (size_t)blockIdx.x * per;
        float2* fac = reinterpret_cast<float2*>(base + (size_t)2 * n * m.cmax);
        float2* nfac = fac + n;
        nd_forward_cta(m, params, spins, -1, -1, base, base + (size_t)n * m.cmax, fac);
        float are = 0.f, aim = 0.f;        // identical in every thread (block sums are broadcast)
        int aligned = 0;
        for (int i = 0; i < n; ++i) {
            if (hamiltonian == QMC_HAMILTONIAN_TFIM) {
                nd_forward_cta(m, params, spins, i, -1, base, base + (size_t)n * m.cmax, nfac);
                const float2 d = nd_block_sum(nfac, fac, n, red);
                float sn, cn;
                const float amp = expf(d.x);
                sincosf(d.y, &sn, &cn);
                are += amp * cn; aim += amp * sn;                            // exp(log_pop), :87
            }
            int rem = i;
            for (int ax = 2; ax >= 3 - m.nd; --ax) {
                const int La = m.L[ax], c = rem % La;
                rem /= La;
                const int j = i + ((c + 1 == La ? 0 : c + 1) - c) * m.str[ax];   // neighbour along +e_d
                const int ss = spins[i] * spins[j];
                if (hamiltonian == QMC_HAMILTONIAN_TFIM) { aligned += ss; continue; }
                if (j == i || ss > 0) { are += 1.f; continue; }                 // -(1-1) exp + 1, :137
                nd_forward_cta(m, params, spins, i, j, base, base + (size_t)n * m.cmax, nfac);
                const float2 d = nd_block_sum(nfac, fac, n, red);
                float sn, cn;
                const float amp = expf(d.x);
                sincosf(d.y, &sn, &cn);
                are += -2.f * amp * cn - 1.f;                                   // -(1+1) exp(log_pop) - 1
                aim += -2.f * amp * sn;
            }
        }
        if (threadIdx.x == 0) {
            if (hamiltonian == QMC_HAMILTONIAN_TFIM) { are = -field_h * are - (float)aligned; aim = -field_h * aim; }
            e_loc[s] = make_float2(are / (float)n, aim / (float)n);            // :89 / :140
        }
        __syncthreads();
    }
}


// flat site index of (site - tap + p): the output site that reads `site` through filter tap `tap`
__device__ __forceinline__ int nd_neighbour_rev(const NdModel& m, int site, int tap) {
    const int p = (m.k - 1) >> 1;
    int q = 0, s = site, t = tap;
#pragma unroll
    for (int a = 2; a >= 0; --a) {
        const int La = m.L[a];
        const int c = s % La; s /= La;
        int d = 0;
        if (a >= 3 - m.nd) { d = p - t % m.k; t /= m.k; }
        q += wrapi(c + d, La) * m.str[a];
    }
    return q;
}

__device__ __forceinline__ float2 nd_ctanh(float a, float b) {      // tanh(a + ib), stable (as qmc_backward.cu)
    const float A = fabsf(a), e = expf(-2.f * A);
    float sb, cb;
    sincosf(b, &sb, &cb);
    const float om = 1.f - e;
    const float den = fmaf(om, om, 4.f * e * cb * cb);
    return make_float2(copysignf((1.f - e * e) / den, a), 4.f * e * sb * cb / den);
}

// gradient of loss_op (mcmc_tf.py:35-56, 172-177) on 1-D / 3-D lattices: grad[p] += sum_n Re[w_n conj(d log psi_n / d p)].
// One CTA per sample (grid-stride): forward with every layer's activations kept (global scratch), cotangent of
// the last layer (Re, Im)(w conj tanh theta), real backprop; per-CTA accumulator in shared memory, one owner
// thread per gradient element, fixed-order reduction over CTAs: deterministic.
__global__ void __launch_bounds__(kNdThreads)
k_nd_backward(NdModel m, const float* __restrict__ params, const int8_t* __restrict__ spins_all,
              const float2* __restrict__ weights, int N, float* scratch, float* partial) {
    extern __shared__ float acc[];
    const int n = m.n, D = m.D;
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const size_t plane = (size_t)n * m.cmax;
    float* base = scratch + (size_t)blockIdx.x * ((size_t)(D + 2) * plane);
    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        const int8_t* spins = spins_all + (size_t)s * n;
        const float2 w = weights[s];
        // forward, keeping act[l] = output of layer l (post-tanh for hidden layers, theta for the last)
        for (int l = 0; l < D; ++l) {
            const int cin = m.cin[l], cout = m.cout[l];
            const float* in = l ? base + (size_t)(l - 1) * plane : nullptr;
            float* out = base + (size_t)l * plane;
            const float* wt = params + m.w_off[l];
            for (int idx = threadIdx.x; idx < n * cout; idx += blockDim.x) {
                const int site = idx / cout, co = idx - site * cout;
                float a = __ldg(params + m.b_off[l] + co);
                for (int tap = 0; tap < m.ktaps; ++tap) {
                    const int q = nd_neighbour(m, site, tap);
                    const float* wr = wt + (size_t)tap * cin * cout + co;
                    if (l == 0) a = fmaf((float)spins[q], __ldg(wr), a);
                    else for (int ci = 0; ci < cin; ++ci) a = fmaf(in[(size_t)q * cin + ci], __ldg(wr + ci * cout), a);
                }
                out[(size_t)site * cout + co] = l == D - 1 ? a : tanhf(a);
            }
            __syncthreads();
        }
        float* G = base + (size_t)D * plane;
        float* Gn = G + plane;
        {   // head cotangent
            const int C = m.cout[D - 1], half = C >> 1;
            const float* th = base + (size_t)(D - 1) * plane;
            for (int t = threadIdx.x; t < n * half; t += blockDim.x) {
                const int site = t / half, c = t - site * half;
                const float2 tc = nd_ctanh(th[(size_t)site * C + c], th[(size_t)site * C + c + half]);
                G[(size_t)site * C + c] = w.x * tc.x + w.y * tc.y;              // w * conj(t)
                G[(size_t)site * C + c + half] = w.y * tc.x - w.x * tc.y;
            }
            if (m.bias_vis_off >= 0 && threadIdx.x == 0) {
                int ssum = 0;
                for (int i = 0; i < n; ++i) ssum += spins[i];
                acc[m.bias_vis_off] += w.x * (float)ssum;
                acc[m.bias_vis_off + 1] += w.y * (float)ssum;
            }
            __syncthreads();
        }
        for (int l = D - 1; l >= 0; --l) {
            const int cin = m.cin[l], cout = m.cout[l];
            const float* in = l ? base + (size_t)(l - 1) * plane : nullptr;
            const int ntask = m.ktaps * cin * cout;
            for (int t = threadIdx.x; t < ntask + cout; t += blockDim.x) {
                if (t >= ntask) {
                    const int co = t - ntask;
                    float sum = 0.f;
                    for (int site = 0; site < n; ++site) sum += G[(size_t)site * cout + co];
                    acc[m.b_off[l] + co] += sum;
                    continue;
                }
                const int co = t % cout, rest = t / cout, ci = rest % cin, tap = rest / cin;
                float sum = 0.f;
                for (int site = 0; site < n; ++site) {
                    const int q = nd_neighbour(m, site, tap);
                    const float a = l ? in[(size_t)q * cin + ci] : (float)spins[q];
                    sum = fmaf(a, G[(size_t)site * cout + co], sum);
                }
                acc[m.w_off[l] + (tap * cin + ci) * cout + co] += sum;
            }
            if (l > 0) {
                const float* wt = params + m.w_off[l];
                for (int t = threadIdx.x; t < n * cin; t += blockDim.x) {
                    const int site = t / cin, ci = t - site * cin;
                    float sum = 0.f;
                    for (int tap = 0; tap < m.ktaps; ++tap) {
                        const int q = nd_neighbour_rev(m, site, tap);
                        const float* wr = wt + ((size_t)tap * cin + ci) * cout;
                        for (int co = 0; co < cout; ++co) sum = fmaf(__ldg(wr + co), G[(size_t)q * cout + co], sum);
                    }
                    const float a = in[(size_t)site * cin + ci];
                    Gn[(size_t)site * cin + ci] = (1.f - a * a) * sum;
                }
            }
            __syncthreads();
            float* tmp = G; G = Gn; Gn = tmp;
        }
    }
    for (int i = threadIdx.x; i < m.P; i += blockDim.x) partial[(size_t)blockIdx.x * m.P + i] = acc[i];
}

__global__ void k_nd_reduce(const float* __restrict__ partial, int nparts, int P, float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * P + i];
    grad[i] += s;
}

static thread_local std::string g_nd_err;

static int nd_grid(int device, int units) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int cap = prop.multiProcessorCount * 4;
    return units < cap ? (units > 0 ? units : 1) : cap;
}

} // namespace qmc

using namespace qmc;

extern "C" {

const char* qmc_nd_last_error(void) { return g_nd_err.c_str(); }

size_t qmc_nd_num_params(const qmc_nd_desc* d) {
    NdModel m;
    std::string err;
    return nd_build(d, m, err) ? (size_t)m.P : 0;
}

size_t qmc_nd_scratch_floats(const qmc_nd_desc* d, int device, int units) {
    NdModel m;
    std::string err;
    if (!nd_build(d, m, err) || units < 1) return 0;
    return (size_t)nd_grid(device, units) * ((size_t)2 * m.n * m.cmax + 4 * (size_t)m.n);
}

#define ND_ENTER()                                                                 \
    NdModel m;                                                                     \
    if (!nd_build(d, m, g_nd_err)) return QMC_ERR_BAD_ARGUMENT;                    \
    int ndev = 0;                                                                  \
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {                   \
        g_nd_err = "no CUDA device: the qmcnn_b200 hot path has no CPU fallback";  \
        return QMC_ERR_NO_DEVICE;                                                  \
    }                                                                              \
    int prev = 0;                                                                  \
    cudaGetDevice(&prev);                                                          \
    if (prev != device) cudaSetDevice(device);
#define ND_LEAVE(what)                                                             \
    cudaError_t e_ = cudaGetLastError();                                           \
    if (prev != device) cudaSetDevice(prev);                                       \
    if (e_ != cudaSuccess) { g_nd_err = std::string(what) + ": " + cudaGetErrorString(e_); return QMC_ERR_CUDA; } \
    return QMC_OK;

int qmc_nd_forward(const qmc_nd_desc* d, int device, const float* params, const int8_t* spins, int N,
                   float* scratch, float* factors, float* logpsi, void* stream) {
    ND_ENTER();
    if (N > 0) {
        if (!params || !spins || !scratch) { g_nd_err = "nd_forward: null argument"; if (prev != device) cudaSetDevice(prev); return QMC_ERR_BAD_ARGUMENT; }
        ++g_launches;
        k_nd_forward<<<nd_grid(device, N), kNdThreads, 0, (cudaStream_t)stream>>>(
            m, params, spins, N, scratch, reinterpret_cast<float2*>(factors), reinterpret_cast<float2*>(logpsi));
    }
    ND_LEAVE("nd_forward");
}

int qmc_nd_sweep(const qmc_nd_desc* d, int device, const float* params, int8_t* spins, float* cur_factors,
                 float* scratch, int S, int num_flips, int64_t step0, int64_t n_steps, const int32_t* flip_pos,
                 const float* uniforms, uint64_t seed, int64_t chain_id0, int64_t therm_its, int64_t its_per_sample,
                 int8_t* samples, int64_t n_sample_slots, uint8_t* accept_trace, float* logratio_trace,
                 unsigned long long* n_accept, void* stream) {
    ND_ENTER();
    if (S > 0 && n_steps > 0) {
        if (!params || !spins || !cur_factors || !scratch || num_flips < 1 || num_flips > QMC_MAX_FLIPS ||
            ((flip_pos == nullptr) != (uniforms == nullptr))) {
            g_nd_err = "nd_sweep: bad argument";
            if (prev != device) cudaSetDevice(prev);
            return QMC_ERR_BAD_ARGUMENT;
        }
        SweepArgs a{spins, nullptr, nullptr, S, num_flips, step0, n_steps, flip_pos, uniforms, seed, chain_id0,
                    therm_its, its_per_sample > 0 ? its_per_sample : 1, samples, n_sample_slots, accept_trace,
                    logratio_trace, n_accept};
        ++g_launches;
        k_nd_sweep<<<nd_grid(device, S), kNdThreads, 0, (cudaStream_t)stream>>>(
            m, params, a, reinterpret_cast<float2*>(cur_factors), scratch);
    }
    ND_LEAVE("nd_sweep");
}

int qmc_nd_local_energy(const qmc_nd_desc* d, int device, int hamiltonian, float field_h, const float* params,
                        const int8_t* spins, int N, float* scratch, float* e_loc, void* stream) {
    ND_ENTER();
    if (N > 0) {
        if (!params || !spins || !scratch || !e_loc ||
            (hamiltonian != QMC_HAMILTONIAN_TFIM && hamiltonian != QMC_HAMILTONIAN_HEISENBERG)) {
            g_nd_err = "nd_local_energy: bad argument";
            if (prev != device) cudaSetDevice(prev);
            return QMC_ERR_BAD_ARGUMENT;
        }
        ++g_launches;
        k_nd_energy<<<nd_grid(device, N), kNdThreads, 0, (cudaStream_t)stream>>>(
            m, params, spins, N, hamiltonian, field_h, scratch, reinterpret_cast<float2*>(e_loc));
    }
    ND_LEAVE("nd_local_energy");
}

size_t qmc_nd_backward_scratch_floats(const qmc_nd_desc* d, int device, int N) {
    NdModel m;
    std::string err;
    if (!nd_build(d, m, err) || N < 1) return 0;
    const size_t ctas = (size_t)nd_grid(device, N);
    return ctas * ((size_t)(m.D + 2) * m.n * m.cmax + (size_t)m.P);
}

int qmc_nd_logpsi_backward(const qmc_nd_desc* d, int device, const float* params, const int8_t* spins,
                           const float* weights, int N, float* scratch, float* grad, void* stream) {
    ND_ENTER();
    if (N > 0) {
        if (!params || !spins || !weights || !scratch || !grad) {
            g_nd_err = "nd_backward: null argument";
            if (prev != device) cudaSetDevice(prev);
            return QMC_ERR_BAD_ARGUMENT;
        }
        const int ctas = nd_grid(device, N);
        float* partial = scratch + (size_t)ctas * ((size_t)(m.D + 2) * m.n * m.cmax);
        const size_t smem = (size_t)m.P * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(k_nd_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            g_nd_err = "nd_backward: parameters do not fit in shared memory";
            if (prev != device) cudaSetDevice(prev);
            return QMC_ERR_UNSUPPORTED;
        }
        g_launches += 2;
        k_nd_backward<<<ctas, kNdThreads, smem, (cudaStream_t)stream>>>(
            m, params, spins, reinterpret_cast<const float2*>(weights), N, scratch, partial);
        k_nd_reduce<<<(m.P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, ctas, m.P, grad);
    }
    ND_LEAVE("nd_backward");
}

} // extern "C"

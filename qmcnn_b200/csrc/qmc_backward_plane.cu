// qmc_backward_plane.cu - K4, layer by layer over ALL samples, with row bands of the lattice resident in shared memory.
// Replaces the gradient half of mcmc_tf.py:172-177 (tf.gradients of loss_op through models.py:95-131) for DCRBM k = 3.
//
// k_backward_smem (qmc_backward.cu) walks one sample through all layers inside one CTA; its weight gradient is one
// thread per (tap, C_in, 4 C_out) looping over the sites with two shared-memory loads per four FMAs (0.07 of the FP32
// roofline at C3), and its planes (3 x 25.6 kB at 20 x 20) stop fitting at 24 x 24, where the L2-resident k_backward
// takes over at 0.01 (819 ms for 8192 samples of 40 x 40).  Here the loop nest is turned inside out:
//
//   k_bwd_head          G_D[n] = (Re, Im)(w_n conj tanh theta_D[n])         (recomputes the last layer's theta)
//   k_bwd_layer  l = D-1 .. 1:   dW_l += sum_n A_{l-1}[n]^T (*) G_l[n],  db_l += sum G_l[n],
//                                 G_{l-1}[n] = (1 - A_{l-1}[n]^2) . (W_l^T (*) G_l[n])
//   k_bwd_layer0        dW_0, db_0 from the spins
//   k_backward_reduce   fixed-order sum of the per-CTA partials
//
// with the cotangent planes G_l[n] of all samples in global memory (2 x N x n x C floats, ping-pong: ~0.2 GB at C3,
// read and written once per layer - nothing next to the arithmetic).  What that buys:
//  * the weight-gradient accumulators of a layer live in REGISTERS for the whole kernel: a group of C_in/4 x C_out/2
//    lanes owns all 9 x C_in x C_out outputs (72 accumulators per lane at 16 -> 16), walks rows of the band with a
//    sliding 3 x 3 window of input float4s (three new LDS.128 + one LDS.64 per 36 FFMA2) and adds into shared memory
//    once, at the end of the kernel, in warp order; CTA partials are summed in CTA order: deterministic;
//  * the cotangent convolution is the forward's split-channel register tile (conv_region_split) over the band with
//    the transposed, tap-reversed weights of ONE layer in shared memory;
//  * a task is (sample, band of BH rows): planes of (BH + 2) x (Lx + 2) sites, so 40 x 40 lattices (C5) run the same
//    kernel in 4 bands; 2-3 CTAs per SM;
//  * the rows of a band arrive by TMA bulk copies (cp.async.bulk, one per channel group and row, completed on an
//    mbarrier) - the copy is linear because cache and G planes are [channel group][site] float4.
#include <vector>
#include "qmc_host.h"
#include "qmc_ip.cuh"

namespace qmc {

constexpr int kBwdPlaneWarps = 6;         // x 2 CTAs per SM: 168 registers (72 weight-gradient accumulators stay live across the conv)

struct BwdPlan {
    int ok;
    int BH, nbands;             // rows per band, bands per sample
    int PW, PA;                 // padded pitch (Lx + 2), padded band area (BH + 2) * PW  [float4 per channel group]
    int plane_floats;           // one padded band, all channel groups (+ one row of slack for the sliding window)
    int P, chunk;               // conv tile: sites per slot, sites per warp
    int tab_entries;
    int gfloats;                // floats of one sample's cotangent plane in global memory: cmaxp * n
    size_t smem;
};

__device__ __forceinline__ float2 ctanh_stable_p(float a, float b) {      // qmc_backward.cu: ctanh_stable
    const float A = fabsf(a), e = expf(-2.f * A);
    float sb, cb;
    sincosf(b, &sb, &cb);
    const float om = 1.f - e;
    const float den = fmaf(om, om, 4.f * e * cb * cb);
    const float re = copysignf((1.f - e * e) / den, a);
    const float im = 4.f * e * sb * cb / den;
    return make_float2(re, im);
}

// ---- TMA: rows of a [channel group][site] float4 plane -> the interior of a wrap-padded band in shared memory ------
// One thread issues one bulk copy per (channel group, band row incl. the two halo rows); everybody waits on the
// mbarrier; then the two halo columns are filled from the interior.  `phase` is the barrier's parity bit.
__device__ __forceinline__ void band_load_issue(float* dst, const float* __restrict__ plane, int ncg, int n, int Ly, int Lx,
                                                int y0, int BH, int PW, int PA, unsigned bar_a) {
    const unsigned row_bytes = (unsigned)Lx * 16u;
    const int nrows = BH + 2;
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(row_bytes * (unsigned)(ncg * nrows)) : "memory");
    __syncthreads();                                   // the expectation is armed before any copy can complete
    for (int i = threadIdx.x; i < ncg * nrows; i += blockDim.x) {
        const int cg = i / nrows, r = i - cg * nrows;
        const int gy = wrap1(y0 - 1 + r, Ly);
        QMC_ASSERT(gy >= 0 && gy < Ly && r * PW + 1 + Lx <= PA, "band row inside lattice and buffer");
        const unsigned d = smem_addr_u32(dst + ((size_t)cg * PA + r * PW + 1) * 4);
        const float* src = plane + ((size_t)cg * n + (size_t)gy * Lx) * 4;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src), "r"(row_bytes), "r"(bar_a) : "memory");
    }
}

__device__ __forceinline__ void band_wait(unsigned bar_a, unsigned phase) {
    const long long t0 = clock64();
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) asm volatile("trap;");
    }
}

__device__ __forceinline__ void band_fill_halo_cols(float* buf, int ncg, int BH, int Lx, int PW, int PA) {
    float4* b4 = reinterpret_cast<float4*>(buf);
    const int nrows = BH + 2;
    for (int i = threadIdx.x; i < ncg * nrows; i += blockDim.x) {
        const int cg = i / nrows, r = i - cg * nrows;
        float4* row = b4 + (size_t)cg * PA + r * PW;
        row[0] = row[Lx];
        row[Lx + 1] = row[1];
    }
}

// ---- weight gradient of one band: a group of NCIG x NCP lanes owns all 9 x CIN x COUT sums --------------------------
//   A4: padded band of the layer's input (float4 = 4 input channels), G: padded band of the cotangent
//   acc[d][j]: (tap d, input channel 4 cig + j, output channels 2 cp, 2 cp + 1)
template <int CIN, int COUT>
struct DwTile {
    static constexpr int NCIG = CIN / 4, NCP = COUT / 2, LG = NCIG * NCP, RP = kWarp / LG;
    static_assert(LG <= kWarp && kWarp % LG == 0, "lane group");
    float2 acc[9][4];
    float2 bacc;

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int d = 0; d < 9; ++d)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[d][j] = make_float2(0.f, 0.f);
        bacc = make_float2(0.f, 0.f);
    }

    __device__ __forceinline__ void site(const float4 (&c0)[3], const float4 (&c1)[3], const float4 (&c2)[3], float2 g) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float4 a[3] = {c0[dy], c1[dy], c2[dy]};
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int d = dy * 3 + dx;
                acc[d][0] = __ffma2_rn(make_float2(a[dx].x, a[dx].x), g, acc[d][0]);
                acc[d][1] = __ffma2_rn(make_float2(a[dx].y, a[dx].y), g, acc[d][1]);
                acc[d][2] = __ffma2_rn(make_float2(a[dx].z, a[dx].z), g, acc[d][2]);
                acc[d][3] = __ffma2_rn(make_float2(a[dx].w, a[dx].w), g, acc[d][3]);
            }
        }
        bacc.x += g.x;
        bacc.y += g.y;
    }

    // rows worker, worker + nworkers, ... of the band; sliding window along x, unrolled by three columns
    __device__ __forceinline__ void band(const float* A, const float* G, int BH, int Lx, int PW, int PA, int lane,
                                         int warp, int nwarps) {
        const int wl = lane % LG, cig = wl / NCP, cp = wl % NCP, worker = warp * RP + lane / LG, nworkers = nwarps * RP;
        const float4* A4 = reinterpret_cast<const float4*>(A) + (size_t)cig * PA;
        const float* Gc = G + ((size_t)(cp >> 1) * PA) * 4 + (cp & 1) * 2;
        for (int y = worker; y < BH; y += nworkers) {
            QMC_ASSERT((y + 2) * PW + ((Lx + 2) / 3) * 3 + 4 < PA + PW, "sliding window inside the band (+ slack row)");
            const float4* ar = A4 + y * PW;                    // padded rows y, y + 1, y + 2 are taps dy = 0, 1, 2
            const float* gr = Gc + (size_t)((y + 1) * PW + 1) * 4;
            float4 c0[3], c1[3], c2[3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) { c0[dy] = ar[dy * PW]; c1[dy] = ar[dy * PW + 1]; }
            for (int x = 0; x < Lx; x += 3) {
                // (the window may run up to two columns past the row: the words are finite - the next row or the
                // slack row - and their cotangent is zeroed)
                float2 g;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) c2[dy] = ar[dy * PW + x + 2];
                g = *reinterpret_cast<const float2*>(gr + (size_t)x * 4);
                site(c0, c1, c2, g);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) c0[dy] = ar[dy * PW + x + 3];
                g = x + 1 < Lx ? *reinterpret_cast<const float2*>(gr + (size_t)(x + 1) * 4) : make_float2(0.f, 0.f);
                site(c1, c2, c0, g);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) c1[dy] = ar[dy * PW + x + 4];
                g = x + 2 < Lx ? *reinterpret_cast<const float2*>(gr + (size_t)(x + 2) * 4) : make_float2(0.f, 0.f);
                site(c2, c0, c1, g);
            }
        }
    }

    // kernel end: lane groups of a warp (fixed order), then the warps of the CTA (index order) into shared memory,
    // then the CTA's partial in the caller's flat parameter order
    __device__ __forceinline__ void flush(float* red, float* __restrict__ partial, const LayerInfo& L, int lane, int warp,
                                          int nwarps) {
        const int wl = lane % LG, cig = wl / NCP, cp = wl % NCP;
#pragma unroll
        for (int o = LG; o < kWarp; o <<= 1) {                 // lanes wl, wl + LG, ... hold the same outputs
#pragma unroll
            for (int d = 0; d < 9; ++d)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[d][j].x += __shfl_down_sync(0xffffffffu, acc[d][j].x, o);
                    acc[d][j].y += __shfl_down_sync(0xffffffffu, acc[d][j].y, o);
                }
            bacc.x += __shfl_down_sync(0xffffffffu, bacc.x, o);
            bacc.y += __shfl_down_sync(0xffffffffu, bacc.y, o);
        }
        for (int i = threadIdx.x; i < 9 * CIN * COUT + COUT; i += blockDim.x) red[i] = 0.f;
        __syncthreads();
        for (int w = 0; w < nwarps; ++w) {
            if (warp == w && lane < LG) {
#pragma unroll
                for (int d = 0; d < 9; ++d)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float* r = red + (d * CIN + cig * 4 + j) * COUT + cp * 2;
                        r[0] += acc[d][j].x;
                        r[1] += acc[d][j].y;
                    }
                if (cig == 0) {
                    red[9 * CIN * COUT + cp * 2] += bacc.x;
                    red[9 * CIN * COUT + cp * 2 + 1] += bacc.y;
                }
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) partial[L.w_off + i] = red[i];
        for (int i = threadIdx.x; i < COUT; i += blockDim.x) partial[L.b_off + i] = red[9 * CIN * COUT + i];
    }
};

// shared-memory carve-up common to the kernels below
struct BwdSmem {
    float* wt;          // transposed, tap-reversed weights of the layer + 16 zeros (the "bias" of the cotangent conv)
    float* A;           // padded band of the layer's input
    float* G;           // padded band of the cotangent
    float* red;         // reduction scratch
    site_t* tab;
    unsigned long long* bar;
};

__device__ __forceinline__ BwdSmem bwd_carve(float* base, const BwdPlan& bp, int wt_floats, int red_floats) {
    BwdSmem s;
    s.bar = reinterpret_cast<unsigned long long*>(base);
    s.wt = base + 4;
    s.A = s.wt + wt_floats;
    s.G = s.A + bp.plane_floats;
    s.red = s.G + bp.plane_floats;
    s.tab = reinterpret_cast<site_t*>(s.red + red_floats);
    return s;
}

__device__ __forceinline__ void bwd_bar_init(unsigned long long* bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// zero a plane buffer once (halo rows / slack must hold finite numbers before the first band arrives)
__device__ __forceinline__ void bwd_zero(float* p, int nfloats) {
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) p[i] = 0.f;
}

template <int K, int CI, int CO, typename OutF>
__device__ __forceinline__ void bwd_conv(int P, int wbase, int bbase, const float* wsm, const float* tin, int PW, int PA,
                                         int Lx, int lane, OutF out, const site_t* tab, int cap4) {
    NoMid mid;
    if (P == 4) conv_region_split<K, CI, CO, 2, 4>(wbase, bbase, wsm, tin, PW, PA, Lx, lane, out, mid, tab, cap4);
    else conv_region_split<K, CI, CO, 2, 6>(wbase, bbase, wsm, tin, PW, PA, Lx, lane, out, mid, tab, cap4);
}

// ---- head: theta of the last layer over the band, G_D = (Re, Im)(w conj tanh theta) --------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(kBwdPlaneWarps * 32, 2)
k_bwd_head(DevModel m, const float* __restrict__ params, const float2* __restrict__ weights, int N,
           const float* __restrict__ cache_all, float* __restrict__ g_out, BwdPlan bp,
           const site_t* __restrict__ tab_g, ImageStrides is, size_t gimg) {
    extern __shared__ float4 smem4[];
    params += (size_t)blockIdx.y * is.params;
    cache_all += (size_t)blockIdx.y * is.cache;
    weights += (size_t)blockIdx.y * N;
    g_out += (size_t)blockIdx.y * gimg;
    const LayerInfo& L = m.layer[m.D - 1];
    constexpr int WF = 9 * CIN * COUT + COUT;                      // this layer's weights + bias, forward layout
    BwdSmem sm = bwd_carve(reinterpret_cast<float*>(smem4), bp, round4(WF), COUT * 0 + 4);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) sm.wt[i] = params[L.sw_off + i];     // coutp == COUT
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sm.wt[9 * CIN * COUT + i] = params[L.sb_off + i];
    for (int i = threadIdx.x; i < bp.tab_entries; i += blockDim.x) sm.tab[i] = tab_g[i];
    bwd_zero(sm.A, bp.plane_floats);
    bwd_zero(sm.G, bp.plane_floats);
    bwd_bar_init(sm.bar);
    const unsigned bar_a = smem_addr_u32(sm.bar);
    const site_t* tab = sm.tab + warp * (bp.P + 1) * 16;
    const int n = m.n, Ly = m.Ly, Lx = m.Lx, PW = bp.PW, PA = bp.PA, half = COUT / 2;
    const float* inoff = cache_all + m.layer[m.D - 2].act_off;
    unsigned phase = 0;
    float4* th4 = reinterpret_cast<float4*>(sm.G);               // theta of the band: [cog][band site]
    const long long ntasks = (long long)N * bp.nbands;
    for (long long t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const int s = (int)(t / bp.nbands), band = (int)(t - (long long)s * bp.nbands);
        const int y0 = band * bp.BH, bh = min(bp.BH, Ly - y0), bsites = bh * Lx;
        band_load_issue(sm.A, inoff + (size_t)s * m.cache_floats, CIN / 4, n, Ly, Lx, y0, bp.BH, PW, PA, bar_a);
        band_wait(bar_a, phase);
        phase ^= 1u;
        band_fill_halo_cols(sm.A, CIN / 4, bp.BH, Lx, PW, PA);
        __syncthreads();
        bwd_conv<3, CIN, COUT>(bp.P, 0, 9 * CIN * COUT, sm.wt, sm.A, PW, PA, Lx, lane,
                               [&](int pos, int, int, int cog, float4 a) { if (pos < bsites) th4[cog * bsites + pos] = a; }, tab,
                               bp.plane_floats >> 2);
        __syncthreads();
        const float2 w = weights[s];
        float* gs = g_out + (size_t)s * bp.gfloats;
        const float* th = sm.G;
        for (int i = threadIdx.x; i < bsites * half; i += blockDim.x) {
            const int c = i / bsites, pos = i - c * bsites, c2 = c + half;
            const float2 tc = ctanh_stable_p(th[((c >> 2) * bsites + pos) * 4 + (c & 3)], th[((c2 >> 2) * bsites + pos) * 4 + (c2 & 3)]);
            const int site = y0 * Lx + pos;
            gs[((c >> 2) * n + site) * 4 + (c & 3)] = w.x * tc.x + w.y * tc.y;              // w * conj(t)
            gs[((c2 >> 2) * n + site) * 4 + (c2 & 3)] = w.y * tc.x - w.x * tc.y;
        }
        __syncthreads();
    }
}

// ---- layer l >= 1: weight / bias gradient and the cotangent of the layer below --------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(kBwdPlaneWarps * 32, 2)
k_bwd_layer(DevModel m, int l, const float* __restrict__ params, int N, const float* __restrict__ cache_all,
            const float* __restrict__ g_in, float* __restrict__ g_out, float* __restrict__ partial, BwdPlan bp,
            const site_t* __restrict__ tab_g, ImageStrides is, size_t gimg) {
    extern __shared__ float4 smem4[];
    params += (size_t)blockIdx.y * is.params;
    cache_all += (size_t)blockIdx.y * is.cache;
    g_in += (size_t)blockIdx.y * gimg;
    g_out += (size_t)blockIdx.y * gimg;
    partial += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * m.P;
    const LayerInfo& L = m.layer[l];
    constexpr int WF = 9 * CIN * COUT;
    BwdSmem sm = bwd_carve(reinterpret_cast<float*>(smem4), bp, round4(WF + 16), round4(WF + COUT));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // cotangent conv: G (COUT channels) -> CIN channels through wt[(d' * COUT + co) * CIN + ci] = W[(d, ci, co)], d' = 8 - d
    for (int i = threadIdx.x; i < WF; i += blockDim.x) {
        const int d = i / (CIN * COUT), r = i - d * CIN * COUT, ci = r / COUT, co = r - ci * COUT;
        sm.wt[((8 - d) * COUT + co) * CIN + ci] = params[L.sw_off + i];
    }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) sm.wt[WF + i] = 0.f;
    for (int i = threadIdx.x; i < bp.tab_entries; i += blockDim.x) sm.tab[i] = tab_g[i];
    bwd_zero(sm.A, bp.plane_floats);
    bwd_zero(sm.G, bp.plane_floats);
    bwd_bar_init(sm.bar);
    const unsigned bar_a = smem_addr_u32(sm.bar);
    const site_t* tab = sm.tab + warp * (bp.P + 1) * 16;
    const int n = m.n, Ly = m.Ly, Lx = m.Lx, PW = bp.PW, PA = bp.PA;
    const float* inoff = cache_all + m.layer[l - 1].act_off;
    unsigned phase = 0;
    DwTile<CIN, COUT> dw;
    dw.clear();
    const float4* A4 = reinterpret_cast<const float4*>(sm.A);
    const long long ntasks = (long long)N * bp.nbands;
    for (long long t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const int s = (int)(t / bp.nbands), band = (int)(t - (long long)s * bp.nbands);
        const int y0 = band * bp.BH, bh = min(bp.BH, Ly - y0), bsites = bh * Lx;
        // both planes of the band on one barrier phase
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                         "r"((unsigned)Lx * 16u * (unsigned)((CIN / 4 + COUT / 4) * (bp.BH + 2))) : "memory");
        __syncthreads();
        {
            const unsigned row_bytes = (unsigned)Lx * 16u;
            const int nrows = bp.BH + 2, na = (CIN / 4) * nrows, ng = (COUT / 4) * nrows;
            const float* pa = inoff + (size_t)s * m.cache_floats;
            const float* pg = g_in + (size_t)s * bp.gfloats;
            for (int i = threadIdx.x; i < na + ng; i += blockDim.x) {
                const bool isa = i < na;
                const int k = isa ? i : i - na, cg = k / nrows, r = k - cg * nrows;
                const int gy = wrap1(y0 - 1 + r, Ly);
                QMC_ASSERT(gy >= 0 && gy < Ly && ((size_t)cg * PA + r * PW + 1 + Lx) * 4 <= (size_t)bp.plane_floats, "band row inside lattice and buffer");
                const unsigned d = smem_addr_u32((isa ? sm.A : sm.G) + ((size_t)cg * PA + r * PW + 1) * 4);
                const float* src = (isa ? pa : pg) + ((size_t)cg * n + (size_t)gy * Lx) * 4;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(d), "l"(src), "r"(row_bytes), "r"(bar_a) : "memory");
            }
        }
        band_wait(bar_a, phase);
        phase ^= 1u;
        band_fill_halo_cols(sm.A, CIN / 4, bp.BH, Lx, PW, PA);
        band_fill_halo_cols(sm.G, COUT / 4, bp.BH, Lx, PW, PA);
        __syncthreads();
        // (a) weight and bias gradient of this layer over the band's own rows
        dw.band(sm.A, sm.G, bh, Lx, PW, PA, lane, warp, nwarps);
        // (b) cotangent of the layer below over the band's sites
        float4* go4 = reinterpret_cast<float4*>(g_out + (size_t)s * bp.gfloats);
        bwd_conv<3, COUT, CIN>(bp.P, 0, WF, sm.wt, sm.G, PW, PA, Lx, lane,
                               [&](int pos, int y, int x, int cog, float4 a) {
                                   if (pos >= bsites) return;
                                   const float4 act = A4[cog * PA + (y + 1) * PW + x + 1];
                                   go4[cog * n + y0 * Lx + pos] = make_float4((1.f - act.x * act.x) * a.x, (1.f - act.y * act.y) * a.y,
                                                                              (1.f - act.z * act.z) * a.z, (1.f - act.w * act.w) * a.w);
                               }, tab, bp.plane_floats >> 2);
        __syncthreads();
    }
    dw.flush(sm.red, partial, L, lane, warp, nwarps);
}

// ---- layer 0: dW_0[d][co] = sum s[site + d] G_0[site][co], db_0 -----------------------------------------------------
__global__ void __launch_bounds__(256)
k_bwd_layer0(DevModel m, const int8_t* __restrict__ spins, int N, const float* __restrict__ g_in,
             float* __restrict__ partial, int gfloats, size_t gimg) {
    // one thread per (sample slice, tap or bias, 4 output channels); per-CTA sums in shared memory in thread order
    extern __shared__ float4 smem4[];
    float* red = reinterpret_cast<float*>(smem4);
    g_in += (size_t)blockIdx.y * gimg;
    partial += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * m.P;
    const LayerInfo& L = m.layer[0];
    const int n = m.n, Ly = m.Ly, Lx = m.Lx, p = m.p, ncog = L.coutp >> 2, ntap = m.k * m.k;
    const int ntask = (ntap + 1) * ncog;                       // tap index ntap = bias
    const int slices = blockDim.x / ntask, slice = threadIdx.x / ntask, task = threadIdx.x - slice * ntask;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slice < slices) {
        const int d = task / ncog, cog = task - d * ncog;
        const int dy = d / m.k - p, dx = d - (d / m.k) * m.k - p;
        for (int s = blockIdx.x * slices + slice; s < N; s += gridDim.x * slices) {
            const int8_t* sx = spins + (size_t)s * n;
            const float4* g4 = reinterpret_cast<const float4*>(g_in + (size_t)s * gfloats) + (size_t)cog * n;
            for (int y = 0; y < Ly; ++y) {
                const int qy = wrap1(y + dy, Ly) * Lx;
                for (int x = 0; x < Lx; ++x) {
                    const float a = d < ntap ? (float)sx[qy + wrap1(x + dx, Lx)] : 1.f;
                    const float4 gv = __ldcg(g4 + y * Lx + x);
                    acc.x = fmaf(a, gv.x, acc.x); acc.y = fmaf(a, gv.y, acc.y);
                    acc.z = fmaf(a, gv.z, acc.z); acc.w = fmaf(a, gv.w, acc.w);
                }
            }
        }
    }
    reinterpret_cast<float4*>(red)[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < ntask) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sl = 0; sl < slices; ++sl) {
            const float4 v = reinterpret_cast<const float4*>(red)[sl * ntask + threadIdx.x];
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const int d = threadIdx.x / ncog, cog = threadIdx.x - d * ncog;
        const float v[4] = {sum.x, sum.y, sum.z, sum.w};
        for (int j = 0; j < 4; ++j) {
            const int co = cog * 4 + j;
            if (co >= L.cout) break;
            if (d < ntap) partial[L.w_off + d * L.cout + co] = v[j];      // cin == 1
            else partial[L.b_off + co] = v[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static BwdPlan bwd_plan(const qmc_handle* h) {
    BwdPlan bp{};
    const DevModel& m = h->m;
    if (!h->allow_tiled || h->backward_generic || h->backward_smem_only || m.kind != QMC_MODEL_DCRBM || m.D < 2 || m.k != 3) return bp;
    if (m.Lx > 255 || m.Ly > 255 || m.Lx < 3 || m.Ly < 3) return bp;
    int cmax = 0;
    for (int l = 0; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        if (l >= 1) {
            const bool hidden_ok = (L.cin == 16 && L.cout == 16) || (L.cin == 8 && L.cout == 8);
            const bool last_ok = hidden_ok || (L.cin == 16 && L.cout == 8);
            if (l < m.D - 1 ? !hidden_ok : !last_ok) return bp;
        }
        if (L.coutp > cmax) cmax = L.coutp;
    }
    bp.PW = m.Lx + 2;
    // band height: the tallest band with at most 8 warps x 6 x 16 conv slots whose two planes leave room for 2 CTAs per SM
    const size_t budget = ((size_t)h->max_smem - 2048) / 2;
    for (int nb = 1; nb <= m.Ly; ++nb) {
        const int BH = (m.Ly + nb - 1) / nb;
        if (BH * m.Lx > kBwdPlaneWarps * 6 * 16) continue;
        const int PA = (BH + 2) * bp.PW;
        const int plane = round4((PA + bp.PW) * cmax);
        const size_t smem = ((size_t)4 + round4(9 * 16 * 16 + 16) + 2 * (size_t)plane + round4(9 * 16 * 16 + 16)) * 4 + sizeof(site_t) * kBwdPlaneWarps * 7 * 16;
        if (smem > budget) continue;
        bp.BH = BH; bp.nbands = nb; bp.PA = PA; bp.plane_floats = plane;
        break;
    }
    if (!bp.BH) return bp;
    const int bs = bp.BH * m.Lx;
    double best = -1;
    int warps = 0;
    for (int P = 6; P >= 4; P -= 2)
        for (int w = 1; w <= kBwdPlaneWarps; ++w) {
            const int chunk = (bs + w - 1) / w;
            if (chunk > P * 16) continue;
            const double util = (double)bs / (double)(w * P * 16);
            if (util > best + 1e-9) { best = util; warps = w; bp.P = P; bp.chunk = chunk; }
            break;
        }
    if (best < 0) return bp;
    // the conv needs `warps`; the weight gradient likes all eight: tables cover 8 warps, surplus ones hold no sites
    bp.tab_entries = (kBwdPlaneWarps * (bp.P + 1) * 16 + 7) & ~7;
    bp.gfloats = cmax * m.n;
    bp.smem = ((size_t)4 + round4(9 * 16 * 16 + 16) + 2 * (size_t)bp.plane_floats + round4(9 * 16 * 16 + 16)) * 4 + (size_t)bp.tab_entries * sizeof(site_t);
    (void)warps;
    bp.ok = 1;
    return bp;
}

bool backward_plane_supported(const qmc_handle* h) { return bwd_plan(h).ok != 0; }

// plane_site_table of qmc_plane.cu, for the band tile
static void band_site_table(int s0, int s1, int Lx, int PW, int P, site_t* tab) {
    const int cnt = s1 > s0 ? s1 - s0 : 0, G = (cnt + P - 1) / P;
    std::vector<char> taken(cnt > 0 ? cnt : 1, 0);
    for (int i = 0; i < (P + 1) * 16; ++i) tab[i] = kNoSite;
    int left = cnt;
    for (int j = 0; j < P; ++j)
        for (int h0 = 0; h0 < 16; h0 += 8) {
            unsigned used = 0;
            for (int slot = h0; slot < h0 + 8 && slot < G && left > 0; ++slot) {
                int pick = -1, fallback = -1;
                for (int i = 0; i < cnt; ++i) {
                    if (taken[i]) continue;
                    if (fallback < 0) fallback = i;
                    const int y = (s0 + i) / Lx, x = (s0 + i) % Lx;
                    if (!((used >> ((y * PW + x) & 7)) & 1u)) { pick = i; break; }
                }
                if (pick < 0) pick = fallback;
                const int y = (s0 + pick) / Lx, x = (s0 + pick) % Lx;
                used |= 1u << ((y * PW + x) & 7);
                taken[pick] = 1;
                --left;
                tab[j * 16 + slot] = make_site(y, x, PW);
            }
        }
    finish_site_table(tab, P, 16);
}

cudaError_t bwd_plane_upload_tables(qmc_handle* h) {
    h->d_bwd_tab = nullptr;
    const BwdPlan bp = bwd_plan(h);
    if (!bp.ok) return cudaSuccess;
    std::vector<site_t> tab(bp.tab_entries, kNoSite);
    const int bs = bp.BH * h->m.Lx;
    for (int w = 0; w < kBwdPlaneWarps; ++w) {
        const int s0 = w * bp.chunk < bs ? w * bp.chunk : bs, s1 = s0 + bp.chunk < bs ? s0 + bp.chunk : bs;
        band_site_table(s0, s1, h->m.Lx, bp.PW, bp.P, tab.data() + (size_t)w * (bp.P + 1) * 16);
    }
    cudaError_t e = cudaMalloc(&h->d_bwd_tab, tab.size() * sizeof(site_t));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(h->d_bwd_tab, tab.data(), tab.size() * sizeof(site_t), cudaMemcpyHostToDevice);
}

static int bwd_plane_ctas(const qmc_handle* h, int nimg, long long ntasks) {
    long long c = (long long)h->num_sms * 2;
    if (nimg > 1) c = c / nimg > 0 ? c / nimg : 1;
    return (int)(ntasks < c ? (ntasks > 0 ? ntasks : 1) : c);
}

size_t backward_plane_workspace_floats(const qmc_handle* h, int nimg, int N) {
    const BwdPlan bp = bwd_plan(h);
    const size_t ctas = (size_t)bwd_plane_ctas(h, nimg, (long long)N * bp.nbands) * nimg;
    return (size_t)nimg * N * h->m.cache_floats + 2 * (size_t)nimg * N * bp.gfloats + ctas * round4(h->m.P) +
           (nimg > 1 ? (size_t)nimg * N * 2 : 0);
}

// grad [nimg, P] (+=) for the per-sample cotangents w [nimg, N] (complex); `cache` holds the images' forward caches
cudaError_t launch_backward_plane(const qmc_handle* h, int nimg, const float* blocks, const int8_t* spins, const float2* w,
                                  int N, const float* cache, float* gbuf, float* partial, float* grad, cudaStream_t st) {
    const DevModel& m = h->m;
    const BwdPlan bp = bwd_plan(h);
    const int ctas = bwd_plane_ctas(h, nimg, (long long)N * bp.nbands);
    const ImageStrides is{(size_t)m.smem_param_floats, (size_t)N * m.cache_floats};
    const size_t gimg = (size_t)N * bp.gfloats;
    float* g0 = gbuf;
    float* g1 = gbuf + (size_t)nimg * gimg;
    const dim3 grid(ctas, nimg);
    const int thr = kBwdPlaneWarps * 32;
    cudaError_t e;
#define QMC_SMEM(kern)                                                                                                  \
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem)) != cudaSuccess) return e
    {
        const LayerInfo& L = m.layer[m.D - 1];
        ++g_launches;
        if (L.cin == 16 && L.cout == 16) {
            QMC_SMEM((k_bwd_head<16, 16>));
            k_bwd_head<16, 16><<<grid, thr, bp.smem, st>>>(m, blocks, w, N, cache, g0, bp, h->d_bwd_tab, is, gimg);
        } else if (L.cin == 16) {
            QMC_SMEM((k_bwd_head<16, 8>));
            k_bwd_head<16, 8><<<grid, thr, bp.smem, st>>>(m, blocks, w, N, cache, g0, bp, h->d_bwd_tab, is, gimg);
        } else {
            QMC_SMEM((k_bwd_head<8, 8>));
            k_bwd_head<8, 8><<<grid, thr, bp.smem, st>>>(m, blocks, w, N, cache, g0, bp, h->d_bwd_tab, is, gimg);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    for (int l = m.D - 1; l >= 1; --l) {
        const LayerInfo& L = m.layer[l];
        ++g_launches;
        if (L.cin == 16 && L.cout == 16) {
            QMC_SMEM((k_bwd_layer<16, 16>));
            k_bwd_layer<16, 16><<<grid, thr, bp.smem, st>>>(m, l, blocks, N, cache, g0, g1, partial, bp, h->d_bwd_tab, is, gimg);
        } else if (L.cin == 16) {
            QMC_SMEM((k_bwd_layer<16, 8>));
            k_bwd_layer<16, 8><<<grid, thr, bp.smem, st>>>(m, l, blocks, N, cache, g0, g1, partial, bp, h->d_bwd_tab, is, gimg);
        } else {
            QMC_SMEM((k_bwd_layer<8, 8>));
            k_bwd_layer<8, 8><<<grid, thr, bp.smem, st>>>(m, l, blocks, N, cache, g0, g1, partial, bp, h->d_bwd_tab, is, gimg);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        float* t = g0; g0 = g1; g1 = t;
    }
#undef QMC_SMEM
    {
        const LayerInfo& L = m.layer[0];
        const int ntask = (m.k * m.k + 1) * (L.coutp >> 2);
        if (ntask > 256) return cudaErrorInvalidValue;
        ++g_launches;
        k_bwd_layer0<<<grid, 256, 256 * 16, st>>>(m, spins, N, g0, partial, bp.gfloats, gimg);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    (void)grad;
    return cudaSuccess;
}

int backward_plane_ctas(const qmc_handle* h, int nimg, int N) {
    const BwdPlan bp = bwd_plan(h);
    return bwd_plane_ctas(h, nimg, (long long)N * bp.nbands);
}

} // namespace qmc

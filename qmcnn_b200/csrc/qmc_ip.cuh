// qmc_ip.cuh - the in-place window evaluator of the persistent sweep kernel (k_sweep_ip).
// Kept out of qmc_device.cuh so that tuning it rebuilds one translation unit only.
// IpPlan (the shared-memory plan) lives in qmc_device.cuh; the design note is above it.
#pragma once
#include "qmc_device.cuh"

namespace qmc {

// Phase timers (variant builds only: scripts/build_variant.sh out.so "-DQMC_IP_PROFILE=1", read with
// qmc_diag_ip_profile): clock64 at the phase boundaries of a proposal, summed per warp, added to a global table at
// the end of the kernel.  Phases: 0 draw + barrier wait at the top of the proposal, 1 spin tile + frame gathers,
// 2 layer 0, 3 barrier waits in front of the layers, 4 conv accumulation loops, 5 tanh epilogues + frame waits,
// 6 head, 7 accept + commit + sample write-out.
#ifndef QMC_IP_PROFILE
#define QMC_IP_PROFILE 0
#endif
constexpr int kIpProfPhases = 8;
#if QMC_IP_PROFILE
__shared__ int s_ip_conv[4];          // warps of scheduler (warp & 3) that are inside a conv accumulation loop right now
__device__ unsigned long long g_ip_conc[4];   // histogram: tap iterations that saw 1, 2, 3, 4+ warps of the scheduler in conv
#endif
struct IpProf {
#if QMC_IP_PROFILE
    long long last, acc[kIpProfPhases];

    __device__ __forceinline__ void start() { last = clock64(); }
    __device__ __forceinline__ void mark(int i) { const long long now = clock64(); acc[i] += now - last; last = now; }
#else
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void mark(int) {}
#endif
};

// Out-of-line epilogue pieces.  The kernel runs 12 warps per SM at unrelated program counters and
// the instruction caches are small (L1.5: 32 KB = 2048 instructions), so code that is executed
// once per proposal must be compact: 64 inlined tanhf per register tile were 18 KB per tile shape
// and the first version of this kernel stalled 34% of its cycles on instruction fetch.  Same
// library tanhf / log2cosh_c as every other path, so the results stay bit-identical.
// (tanh_small / kTanhSmall / ip_tanh4_any: qmc_device.cuh - the classic evaluator uses the same fast path)
#ifndef QMC_IP_TANH_INLINE
#define QMC_IP_TANH_INLINE 0      // 1: the small-argument path inlined at every call site (variant build; measured, profiles/r02_summary.md)
#endif
#if QMC_IP_TANH_INLINE
__device__ __forceinline__ float4 ip_tanh4(float4 a) {
#else
static __device__ __noinline__ float4 ip_tanh4(float4 a) {
#endif
    const float mx = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
    if (!(mx < kTanhSmall)) return ip_tanh4_any(a);          // (NaN goes to tanhf too)
    a.x = tanh_small(a.x); a.y = tanh_small(a.y); a.z = tanh_small(a.z); a.w = tanh_small(a.w);
    return a;
}

// sum_c Re log 2cosh(theta_c + i theta_{c+half}) of one site (site_factor<false> without the CRBM term)
static __device__ __noinline__ float ip_site_re(const float* th, int npos, int pos, int half) {
    float re = 0.f;
    // four channels at a time: the loads first, then four independent exp / sincos / log chains the
    // scheduler can interleave (one channel per iteration left this latency-bound at ILP 1); the sum
    // keeps the order c = 0, 1, 2, ... of site_factor
    int c = 0;
    for (; c + 4 <= half; c += 4) {
        float a[4], b[4], r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c1 = c + j, c2 = c1 + half;
            a[j] = th[((c1 >> 2) * npos + pos) * 4 + (c1 & 3)];
            b[j] = th[((c2 >> 2) * npos + pos) * 4 + (c2 & 3)];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float i1 = 0.f;
            log2cosh_c<false>(a[j], b[j], r[j], i1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) re += r[j];
    }
    for (; c < half; ++c) {
        const int c2 = c + half;
        const float a = th[((c >> 2) * npos + pos) * 4 + (c & 3)];
        const float b = th[((c2 >> 2) * npos + pos) * 4 + (c2 & 3)];
        float r1, i1 = 0.f;
        log2cosh_c<false>(a, b, r1, i1);
        re += r1;
    }
    return re;
}

// (Re, Im) sum_c log 2cosh(theta_c + i theta_{c+half}) of one site: site_factor<true> without the CRBM term,
// one channel at a time in site_factor's order (the local-energy kernel needs the complex ratio)
static __device__ __noinline__ float2 ip_site_reim(const float* th, int npos, int pos, int half) {
    float re = 0.f, im = 0.f;
    for (int c = 0; c < half; ++c) {
        const int c2 = c + half;
        const float a = th[((c >> 2) * npos + pos) * 4 + (c & 3)];
        const float b = th[((c2 >> 2) * npos + pos) * 4 + (c2 & 3)];
        float r1, i1;
        log2cosh_c<true>(a, b, r1, i1);
        re += r1;
        im += i1;
    }
    return make_float2(re, im);
}

// cp.async the cache values of the frame of half-widths (a, a+p] around the centre into the arena
__device__ __forceinline__ void ip_gather_frame(float4* arena4, const IpPlan& ip, const float* __restrict__ plane,
                                                int ncg, int n, int a, int p, unsigned mgW, int y0, int x0,
                                                int Ly, int Lx, int lane) {
    const int b = a + p, W = 2 * b + 1, nt2 = 2 * p * W, total = nt2 + (2 * a + 1) * 2 * p;
    const FastDiv dW(mgW, W), d2p(ip.mg2p, 2 * p);
    for (int idx = lane; idx < total; idx += kWarp) {
        int dy, dx;
        if (idx < nt2) {                       // p rows above and p rows below, full width
            const int r = dW.div(idx);
            dx = idx - r * W - b;
            dy = r < p ? r - b : a + 1 + (r - p);
        } else {                               // the 2a+1 middle rows: p columns left, p columns right
            const int j = idx - nt2, r = d2p.div(j), cc = j - r * 2 * p;
            dy = r - a;
            dx = cc < p ? cc - b : a + 1 + (cc - p);
        }
        const int site = wrap1(y0 + dy, Ly) * Lx + wrap1(x0 + dx, Lx);
        QMC_ASSERT(site >= 0 && site < n, "frame gather: lattice site");
        QMC_ASSERT(ip.c + dy >= 0 && ip.c + dy < ip.T && ip.c + dx >= 0 && ip.c + dx < ip.T, "frame gather: arena position");
        QMC_ASSERT((ncg * ip.tarea) * 4 <= ip.arena_floats, "frame gather: arena planes");
        float4* dst = arena4 + (ip.c + dy) * ip.T + (ip.c + dx);
        const float* src = plane + (size_t)site * 4;
        for (int cg = 0; cg < ncg; ++cg) cp_async16(dst + cg * ip.tarea, src + (size_t)cg * n * 4);
    }
}

// The chain's lattice in shared memory, ONE BIT per spin (bit = spin up): 200 bytes instead of 1600 at 40 x 40, which is
// what lets a 12th warp fit next to the weight block at C5 (11 warps before: 17.7 / 19.2 M proposals/s).
__host__ __device__ inline int ip_spin_words(int n) { return (n + 31) >> 5; }
__device__ __forceinline__ int ip_spin(const unsigned* spw, int site) { return ((spw[site >> 5] >> (site & 31)) & 1u) ? 1 : -1; }
__device__ __forceinline__ void ip_pack_spins(unsigned* spw, const int8_t* __restrict__ g, int n, int lane) {
    for (int base = 0; base < n; base += kWarp) {
        const int i = base + lane;
        const unsigned w = __ballot_sync(0xffffffffu, i < n && g[i] > 0);
        if (lane == 0) spw[base >> 5] = w;
    }
}
__device__ __forceinline__ void ip_unpack_spins(int8_t* __restrict__ g, const unsigned* spw, int n, int lane) {
    for (int i = lane; i < n; i += kWarp) g[i] = ((spw[i >> 5] >> lane) & 1u) ? 1 : -1;     // i & 31 == lane
}

// ---------------------------------------------------------------------------------------------------------------
// Split-channel register tile.  conv_region_tiled gives a lane P sites x ALL output channels, so every weight is an
// all-lane broadcast LDS.128 - two shared-memory wavefronts for 16 bytes - and the conv loops of k_sweep_ip were
// co-limited by the shared-memory pipe (r01 capture: 10.96k wavefronts per proposal, 5.2k of them weight broadcasts;
// 4 SMSPs x 10.96k = 75% of the SM's wavefront slots at the measured rate, next to 51% of the FMA pipe).
// scripts/proto/lds_patterns.cu measured what an LDS.128 costs on sm_100: 4 cycles for 32 different addresses, 2 when
// the 16 lanes of each half-warp ask for <= 8 different 16-byte words in different bank groups AND every aligned quad
// of lanes holds at most two different addresses (lane & 3 patterns cost 4; lane >> 1, lane & 1 cost 2).
// So here lane = (site slot, channel part): CS = 2 parts of COUT / 2 channels, slot = lane >> 1, part = lane & 1.
// A weight load (2 different addresses, interleaved) still costs 2 but now feeds half as many loads per lane; an
// input load (16 different sites, 8 per half-warp, dealt conflict-free by the site table) costs 2 instead of 4.
// Per (tap, 4 input channels): 8 x 2 + 2P x 2 wavefronts instead of 16 x 2 + P x 4 (P = sites per lane before): 20 / 24
// / 28 / 32 instead of 36 / 40 / 44 / 48 for the 5x5 ... 11x11 windows.  The fma chain of an output value is unchanged
// (bias, taps ascending, input channels ascending), so results stay bit-identical.
//   tab: P rows of NS = 32 / CS entries make_site(y, x, tile pitch) + one row of valid masks (qmc_device.cuh: site_t).
// ---------------------------------------------------------------------------------------------------------------
#ifndef QMC_IP_L0_FASTDIV
#define QMC_IP_L0_FASTDIV 0   // layer 0 (generic conv over the 3x3 window): site decode by integer division (0) or FastDiv (1).
                              // The FastDiv form executes fewer instructions and measured 0.32% SLOWER (22.336 vs 22.407 M proposals/s,
                              // twice each, alternating: profiles/r02_summary.md) - register allocation and code layout of the whole
                              // kernel move with it (145 vs 149 registers)
#endif
#ifndef QMC_IP_SPLIT
#define QMC_IP_SPLIT 4        // 0: one-part tiles (conv_region_tiled); 2: two channel parts; 4: + four parts for the 7x7 and 9x9
                              // windows (56 and 88 site slots instead of 64 and 96: 21.45 -> 21.99 M proposals/s).  Four parts
                              // for the 5x5 window as well (same 32 slots, fewer weight loads) measured 16.0 vs 19.8: not used.
#endif

//   TA > 0: the tile's plane stride (float4 words) is this compile-time constant (the in-place arena: kIpPlane), so the
//   second channel group of a pair is an immediate offset of the first one's address: one integer add per (site, tap,
//   PAIR of channel groups) instead of ~3 per (site, tap, channel group) - they were 5% of k_sweep_ip's instructions.
template <int K, int CIN, int COUT, int CS, int P, int TA = 0, typename OutF, typename MidF>
__device__ __forceinline__ void conv_region_split(int wbase, int bbase, const float* wsm, const float* tin, int tw,
                                                  int tarea, int side, int lane, OutF out, MidF mid,
                                                  const site_t* tab, int cap4 = 0x7fffffff) {
    QMC_ASSERT(TA == 0 || TA == tarea, "compile-time plane stride");
    // cap4 (debug builds): float4 words of the input tile buffer - every input read is checked against it
    static_assert(CS == 2 || CS == 4, "channel parts");
    constexpr int CL = COUT / CS;             // output channels of this lane
    constexpr int NS = kWarp / CS;            // site slots
    constexpr int NCG = CIN / 4;
    static_assert(CIN % 4 == 0 && CL % 4 == 0, "shape");
    // aligned lane quads hold two slots x two parts (lds_patterns.cu)
    const int part = CS == 2 ? (lane & 1) : ((lane >> 1) & 3);
    const int slot = CS == 2 ? (lane >> 1) : ((lane & 1) | ((lane >> 3) << 1));
    // the lane's P input positions as shared-memory byte addresses; the (tap, channel group) part of an address is
    // warp-uniform, so that the loads below can take the form LDS [R + UR] with no per-lane address arithmetic
    unsigned pin[P];
    unsigned valid = 0;                       // bit j: site j of this slot is a real output
    {
        const unsigned tin_a = smem_addr_u32(tin);
        valid = tab[P * NS + slot];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const site_t pk = tab[j * NS + slot];
            QMC_ASSERT((int)(pk >> 16) == ((int)((pk >> 8) & 255u) * tw + (int)(pk & 255u)) * 16, "site table: offset = (y * pitch + x) * 16");
            QMC_ASSERT((int)(pk >> 20) + (K - 1) * tw + (K - 1) + (NCG - 1) * tarea < cap4, "conv input window inside the tile buffer");
            pin[j] = tin_a + (pk >> 16);
        }
    }
    const int wlane = wbase + part * CL, blane = bbase + part * CL;
#if QMC_IP_PROFILE
    const int sched_ = (threadIdx.x >> 5) & 3;
    int conc_[4] = {0, 0, 0, 0};
    if (lane == 0) atomicAdd(&s_ip_conv[sched_], 1);
#endif
    // accumulators start from the bias: one broadcast load per accumulator pair instead of one load and P register moves
    // per pair (64 MOVs for the 8-site tile)
    float2 acc[P][CL / 2];
    {
        const unsigned b_a = smem_addr_u32(wsm + blane);
#pragma unroll
        for (int j = 0; j < P; ++j)
#pragma unroll
            for (int q2 = 0; q2 < CL / 2; ++q2)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(acc[j][q2].x), "=f"(acc[j][q2].y) : "r"(b_a + 8u * q2));
    }
#pragma unroll 1
    for (int d = 0; d < K * K; ++d) {
#if QMC_IP_PROFILE
        { const int v_ = *(volatile int*)&s_ip_conv[sched_]; ++conc_[v_ < 1 ? 0 : v_ > 4 ? 3 : v_ - 1]; }
#endif
        const int dy = d / K, dx = d - dy * K;
        const unsigned toff = (unsigned)(dy * tw + dx) * 16u;
        const int wrow = wlane + d * CIN * COUT;
        // one channel group: four input channels of every site against this lane's CL output channels
        auto fma_group = [&](int cg, const float4 (&in)[P]) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float2 w[CL / 2];
#pragma unroll
                for (int q4 = 0; q4 < CL / 4; ++q4) {
                    const float4 t = *reinterpret_cast<const float4*>(wsm + wrow + (cg * 4 + c4) * COUT + q4 * 4);
                    w[q4 * 2] = make_float2(t.x, t.y);
                    w[q4 * 2 + 1] = make_float2(t.z, t.w);
                }
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const float v = c4 == 0 ? in[j].x : c4 == 1 ? in[j].y : c4 == 2 ? in[j].z : in[j].w;
                    const float2 v2 = make_float2(v, v);
#pragma unroll
                    for (int q2 = 0; q2 < CL / 2; ++q2) acc[j][q2] = __ffma2_rn(v2, w[q2], acc[j][q2]);
                }
            }
        };
        if constexpr (TA > 0 && NCG % 2 == 0) {
#pragma unroll 1
            for (int cp = 0; cp < NCG / 2; ++cp) {             // pairs of channel groups: the loop body of kCgUnroll = 2
                unsigned a[P];
#pragma unroll
                for (int j = 0; j < P; ++j) a[j] = pin[j] + toff + (unsigned)cp * (2u * TA * 16u);
                float4 in0[P], in1[P];
#pragma unroll
                for (int j = 0; j < P; ++j)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(in0[j].x), "=f"(in0[j].y), "=f"(in0[j].z), "=f"(in0[j].w) : "r"(a[j]));
#pragma unroll
                for (int j = 0; j < P; ++j)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                                 : "=f"(in1[j].x), "=f"(in1[j].y), "=f"(in1[j].z), "=f"(in1[j].w) : "r"(a[j]), "n"(TA * 16));
                fma_group(2 * cp, in0);
                fma_group(2 * cp + 1, in1);
            }
        } else {
#pragma unroll(kCgUnroll)
            for (int cg = 0; cg < NCG; ++cg) {
                float4 in[P];
                const unsigned uoff = toff + (unsigned)(cg * tarea) * 16u;          // warp-uniform
#pragma unroll
                for (int j = 0; j < P; ++j)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(in[j].x), "=f"(in[j].y), "=f"(in[j].z), "=f"(in[j].w) : "r"(pin[j] + uoff));
                fma_group(cg, in);
            }
        }
    }
#if QMC_IP_PROFILE
    if (lane == 0) {
        atomicSub(&s_ip_conv[sched_], 1);
        for (int i = 0; i < 4; ++i) if (conc_[i]) atomicAdd(&g_ip_conc[i], (unsigned long long)conc_[i]);
    }
#endif
    mid();
#pragma unroll
    for (int j = 0; j < P; ++j) {
        if (!((valid >> j) & 1u)) continue;
        const site_t pk = tab[j * NS + slot];            // re-read: cheaper than P live registers across the loop
        const int y = (int)((pk >> 8) & 255u), x = (int)(pk & 255u);
        const int pos = y * side + x;
#pragma unroll
        for (int q4 = 0; q4 < CL / 4; ++q4)
            out(pos, y, x, part * (CL / 4) + q4,
                make_float4(acc[j][q4 * 2].x, acc[j][q4 * 2].y, acc[j][q4 * 2 + 1].x, acc[j][q4 * 2 + 1].y));
    }
}

// The register tile of a tiled layer of the in-place evaluator: cs channel parts, p sites per lane, ns = 32 / cs site
// slots (host and device; p == 0: the window does not fit one round, the model is outside the evaluator's coverage)
struct IpTile { int cs, p, ns; };
__host__ __device__ inline IpTile ip_tile(int acc, int cout, int npos) {
    IpTile t;
#if QMC_IP_SPLIT
    // 16 output channels, 7x7 and 9x9 windows: four channel parts x eight site slots (finer rounding of the window:
    // 56 instead of 64 site slots for 7x7, 88 instead of 96 for 9x9, and half the weight loads); else two parts
    if (QMC_IP_SPLIT >= 4 && cout == 16 && npos > 32 && npos <= 88) {
        t.cs = 4;
        t.ns = kWarp / t.cs;
        const int need = (npos + t.ns - 1) / t.ns;
        t.p = need <= 7 ? 7 : 11;
        return t;
    }
    t.cs = 2;
    t.ns = kWarp / t.cs;
    const int need = (npos + t.ns - 1) / t.ns;
    t.p = need <= 2 ? 2 : need <= 4 ? 4 : need <= 6 ? 6 : need <= 8 ? 8 : need <= 11 ? 11 : 0;   // instantiated heights
    if (t.p * (cout / t.cs) > acc) t.p = 0;
#else
    t.cs = 1;
    t.ns = kWarp;
    const int pmax = acc / cout;
    t.p = (pmax == 1 || npos <= 32) ? 1 : (pmax == 2 || npos <= 64) ? 2 : (pmax == 3 || npos <= 96) ? (pmax >= 3 ? 3 : 2)
        : (pmax == 4 || npos <= 128) ? (pmax >= 4 ? 4 : 2) : (pmax < 8 || npos <= 192) ? (pmax >= 6 ? 6 : 4)
        : (pmax >= 8 ? 8 : 4);
    if ((npos + t.p - 1) / t.p > kWarp) t.p = 0;
#endif
    return t;
}

template <int K, int CIN, int COUT, int ACC, int TA = 0, typename OutF, typename MidF>
__device__ __forceinline__ void conv_region_pick_ip(int wbase, int bbase, const float* wsm, const float* tin,
                                                    int tw, int tarea, int side, int lane, OutF out, MidF mid,
                                                    const site_t* tab, int cap4 = 0x7fffffff) {
    const int npos = side * side;
#if QMC_IP_SPLIT
    constexpr int CL = COUT / 2;
    if constexpr (QMC_IP_SPLIT >= 4 && COUT == 16) {
#define QMC_SPLIT4(PP) conv_region_split<K, CIN, COUT, 4, (PP), TA>(wbase, bbase, wsm, tin, tw, tarea, side, lane, out, mid, tab, cap4)
        if (npos > 32 && npos <= 7 * 8) return QMC_SPLIT4(7);
        if (npos > 7 * 8 && npos <= 11 * 8) return QMC_SPLIT4(11);
#undef QMC_SPLIT4
    }
#define QMC_SPLIT(PP) conv_region_split<K, CIN, COUT, 2, (PP), TA>(wbase, bbase, wsm, tin, tw, tarea, side, lane, out, mid, tab, cap4)
    if (npos <= 2 * 16) return QMC_SPLIT(2);
    if (npos <= 4 * 16) return QMC_SPLIT(4);
    if (npos <= 6 * 16) return QMC_SPLIT(6);
    if constexpr (8 * CL <= ACC) { if (npos <= 8 * 16) return QMC_SPLIT(8); }
    if constexpr (11 * CL <= ACC) return QMC_SPLIT(11);
#undef QMC_SPLIT
#else
    constexpr int PMAX = ACC / COUT;
    auto o = [&](int, int pos, int y, int x, int cog, float4 a) { out(pos, y, x, cog, a); };
#define QMC_TILED(PP) conv_region_tiled<K, CIN, COUT, (PP), 1>(wbase, bbase, wsm, tin, 0, tw, tarea, side, side, lane, o, mid, tab)
    if (PMAX == 1 || npos <= 32) return QMC_TILED(1);
    if (PMAX == 2 || npos <= 64) return QMC_TILED(2);
    if (PMAX == 3 || npos <= 96) return QMC_TILED(PMAX >= 3 ? 3 : 2);
    if (PMAX == 4 || npos <= 128) return QMC_TILED(PMAX >= 4 ? 4 : 2);
    if (PMAX < 8 || npos <= 192) return QMC_TILED(PMAX >= 6 ? 6 : 4);
    return QMC_TILED(PMAX >= 8 ? 8 : 4);
#undef QMC_TILED
#endif
}

// barrier of the warp's phase group: SYNC 2 = the whole CTA; SYNC 1, 3 = the four warps w/4 == g (one per
// scheduler), so that an SM runs three phases at a time: few enough for the instruction caches, different
// enough for the schedulers to overlap one group's memory latency with another group's arithmetic
template <int SYNC>
__device__ __forceinline__ void ip_barrier(int gid, int gthreads) {
    if (SYNC == 3 || SYNC == 1) asm volatile("bar.sync %0, %1;" ::"r"(gid), "r"(gthreads) : "memory");
    else if (SYNC) __syncthreads();
}

// SYNC == 2: a CTA barrier in front of every layer, so that all warps of the SM run the same loop
// body at the same time (instruction-cache locality; every warp of the CTA must call this the same
// number of times).
// Staging layout of a layer's window: plane-major [cog][pos] float4 (default), or - QMC_IP_STAGE_SITEMAJOR, measured
// and not adopted, profiles/r02_summary.md - site-major [pos][slot] with slot = ip_stage_slot(cog), so that the
// channel groups a warp-wide store of the split-channel tile writes for one site (one per channel part) are adjacent.
#ifndef QMC_IP_STAGE_SITEMAJOR
#define QMC_IP_STAGE_SITEMAJOR 0
#endif
// float4 index of (site pos, channel group cog) in a layer's staged window
__host__ __device__ inline int ip_stage_slot(int cog, int ncg, int cs);
__host__ __device__ inline int ip_stage_index(int pos, int cog, int rarea, int ncg, int cs) {
#if QMC_IP_STAGE_SITEMAJOR
    return pos * ncg + ip_stage_slot(cog, ncg, cs);
#else
    return cog * rarea + pos;
#endif
}
__host__ __device__ inline int ip_stage_slot(int cog, int ncg, int cs) {
    // channel groups per part nq = ncg / cs (all powers of two): cog = part * nq + q -> slot = q * cs + part
    const int lq = ncg >= 4 * cs ? 2 : ncg >= 2 * cs ? 1 : 0;
    return cs == 1 ? cog : (cog & ((1 << lq) - 1)) * cs + (cog >> lq);
}

// commit: scatter the staged window of layer l (src: [pos][slot] float4, in shared memory) into the cache plane
__device__ __forceinline__ void ip_scatter_layer(const DevModel& m, const LayerInfo& L, const float4* src, float* cache,
                                                 int side, unsigned mg_side, int ry, int rx, int lane, int cs) {
    const int rarea = side * side, ncg = L.coutp >> 2, n = m.n;
    float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
    const FastDiv dside(mg_side, side);
    for (int pos = lane; pos < rarea; pos += kWarp) {
        const int y = dside.div(pos), x = pos - y * side;
        const int site = wrap1(ry + y, m.Ly) * m.Lx + wrap1(rx + x, m.Lx);
        QMC_ASSERT(site >= 0 && site < n && L.act_off >= 0, "commit scatter: cache site");
        for (int cg = 0; cg < ncg; ++cg) plane4[cg * n + site] = src[ip_stage_index(pos, cg, rarea, ncg, cs)];
    }
}

// SWEEP = true: the Metropolis kernel - new windows are also written to `staging` (and start their
// speculative commit copy), only Re of the log-ratio is formed.  SWEEP = false: the local-energy kernel -
// nothing is staged (the cache is read-only), Re and Im are formed (dim).
template <int ACC, int SYNC, bool SWEEP = true>
__device__ __forceinline__ void warp_eval_flip_ip(const DevModel& m, const IpPlan& ip, const float* sp,
                                                  const unsigned* mg, float* arena, float* spt, const unsigned* spins_s,
                                                  const float* __restrict__ cache, float* staging,
                                                  int site_f, int lane, int gid, int gthreads, float& dre,
                                                  float* dim_out, const int* tabo, const site_t* tab_s,
                                                  IpProf& prof) {
    const int p = m.p, Ly = m.Ly, Lx = m.Lx, n = m.n, D = m.D;
    const int T = ip.T, TA = ip.tarea, c = ip.c;
    const int y0 = site_f / Lx, x0 = site_f - y0 * Lx;
    float4* arena4 = reinterpret_cast<float4*>(arena);
    // spin tile (side 1 + 4p) with the flip applied
    const int stw = 1 + 4 * p;
    {
        const FastDiv dtw(mg[0], stw);            // W_0 = 4p + 1
        for (int idx = lane; idx < stw * stw; idx += kWarp) {
            const int ty = dtw.div(idx), tx = idx - ty * stw;
            const int site = wrap1(y0 - 2 * p + ty, Ly) * Lx + wrap1(x0 - 2 * p + tx, Lx);
            QMC_ASSERT(site >= 0 && site < n && idx < ip.spt_floats, "spin tile");
            int s = ip_spin(spins_s, site);
            if (site == site_f) s = -s;
            spt[idx] = (float)s;
        }
    }
    // the arena is free: both frames of layer-0 activations around its window (half-widths (p, 3p])
    {
        const LayerInfo& L = m.layer[0];
        ip_gather_frame(arena4, ip, cache + L.act_off, L.coutp >> 2, n, p, p, mg[0], y0, x0, Ly, Lx, lane);
        ip_gather_frame(arena4, ip, cache + L.act_off, L.coutp >> 2, n, 2 * p, p, mg[1], y0, x0, Ly, Lx, lane);
    }
    __syncwarp();
    prof.mark(1);
    int stg = 0;
    {   // layer 0 (C_in = 1) reads the spin tile, not the arena
        const LayerInfo& L = m.layer[0];
        const int side = 1 + 2 * p, rarea = side * side, o0 = c - p;
        float4* stg4 = reinterpret_cast<float4*>(staging);
        conv_region_generic<QMC_IP_L0_FASTDIV != 0>(L, m.k, sp, spt, stw, stw * stw, side, side, lane,
                            [&](int pos, int y, int x, int cog, float4 a) {
                                a = ip_tanh4(a);
                                QMC_ASSERT((cog * TA + (y + o0) * T + (x + o0)) * 4 + 3 < ip.arena_floats && y + o0 < T && x + o0 < T,
                                           "layer 0 output inside the arena");
                                QMC_ASSERT(!SWEEP || ip_stage_index(pos, cog, rarea, L.coutp >> 2, 1) * 4 + 3 < ip.staging_floats,
                                           "layer 0 output inside the staging");
                                arena4[cog * TA + (y + o0) * T + (x + o0)] = a;
                                if (SWEEP) stg4[ip_stage_index(pos, cog, rarea, L.coutp >> 2, 1)] = a;
                            });
        stg += L.coutp * rarea;
        cp_async_wait_all();
        __syncwarp();
        prof.mark(2);
    }
    int side = 1 + 2 * p;
    for (int l = 1; l < D; ++l) {
        const LayerInfo& L = m.layer[l];
        const bool last = (l == D - 1);
        if (SYNC >= 2) ip_barrier<SYNC>(gid, gthreads);
        prof.mark(3);
        const int hn = (l + 1) * p;                  // half-width of this layer's window
        side += 2 * p;
        const int rarea = side * side;
        const float* tin = arena + (size_t)((c - hn - p) * T + (c - hn - p)) * 4;
        const int cap4 = (ip.arena_floats >> 2) - ((c - hn - p) * T + (c - hn - p));      // float4 words from tin to the arena's end
        QMC_ASSERT(c - hn - p >= 0, "layer input box inside the arena");
        const site_t* tab = (tabo && tabo[l] >= 0) ? tab_s + tabo[l] : nullptr;   // conflict-free site deal
        auto sync = [&] { __syncwarp(); prof.mark(4); };
        if (!last) {
            const float* plane = cache + L.act_off;
            const int ncg = L.coutp >> 2, o0 = c - hn;
            ip_gather_frame(arena4, ip, plane, ncg, n, hn + p, p, mg[l + 1], y0, x0, Ly, Lx, lane);   // outer: free now
            float4* stg4 = reinterpret_cast<float4*>(staging + stg);
            const int tcs = ip_tile(ACC, L.cout, rarea).cs;
            auto out = [&](int pos, int y, int x, int cog, float4 a) {
                a = ip_tanh4(a);
                QMC_ASSERT((cog * TA + (y + o0) * T + (x + o0)) * 4 + 3 < ip.arena_floats && y + o0 < T && x + o0 < T && pos < rarea,
                           "hidden-layer output inside the arena");
                QMC_ASSERT(!SWEEP || stg + ip_stage_index(pos, cog, rarea, ncg, tcs) * 4 + 3 < ip.staging_floats,
                           "hidden-layer output inside the staging");
                arena4[cog * TA + (y + o0) * T + (x + o0)] = a;
                if (SWEEP) stg4[ip_stage_index(pos, cog, rarea, ncg, tcs)] = a;
            };
            // all lanes have read the input tile: the inner frame may land now, under the tanh epilogue
            auto mid = [&] {
                __syncwarp();
                prof.mark(4);
                ip_gather_frame(arena4, ip, plane, ncg, n, hn, p, mg[l], y0, x0, Ly, Lx, lane);
            };
            if (L.cin == 16 && L.cout == 16)
                conv_region_pick_ip<3, 16, 16, ACC, kIpPlane>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, mid, tab, cap4);
            else
                conv_region_pick_ip<3, 8, 8, ACC, kIpPlane>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, mid, tab, cap4);
            stg += L.coutp * rarea;
            cp_async_wait_all();
        } else {
            auto out = [&](int pos, int, int, int cog, float4 a) {
                QMC_ASSERT((cog * rarea + pos) * 4 + 3 < ip.newf_off && pos < rarea, "last-layer theta below the new factors");
                arena4[cog * rarea + pos] = a;
            };
            if (L.cin == 16 && L.cout == 16)
                conv_region_pick_ip<3, 16, 16, ACC, kIpPlane>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab, cap4);
            else if (L.cin == 16 && L.cout == 8)
                conv_region_pick_ip<3, 16, 8, ACC, kIpPlane>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab, cap4);
            else
                conv_region_pick_ip<3, 8, 8, ACC, kIpPlane>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab, cap4);
        }
        __syncwarp();
        prof.mark(5);
    }
    if (SYNC >= 2) ip_barrier<SYNC>(gid, gthreads);
    prof.mark(3);
    // commit, part 1 (speculative): the staged windows of the first layers start their way from L2 into the
    // free part of the arena now, so that an accepted move finds them in shared memory after the head
    if (SWEEP) {
        QMC_ASSERT(ip.spec_off + ip.spec_floats <= ip.newf_off && ip.spec_floats <= ip.staging_floats, "speculative commit copy");
        for (int i = lane * 4; i < ip.spec_floats; i += kWarp * 4) cp_async16(arena + ip.spec_off + i, staging + i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // head over the last window: same lane ownership and order as warp_eval_flip
    const int npos = side * side, ry = y0 - D * p, rx = x0 - D * p;
    float* newf = arena + ip.newf_off;
    const FastDiv drw(mg[D - 2], side);          // side = 2Dp + 1 = W_{D-2}
    // old factors first (their L2 latency overlaps the transcendental work); lane k owns the window
    // sites == k (mod 32) in increasing order, like warp_eval_flip; the host checks npos <= 8 * 32
    const int half = m.layer[D - 1].cout >> 1;
    float ore[8], oim[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pos = j * kWarp + lane;
        ore[j] = oim[j] = 0.f;
        if (pos < npos) {
            const int y = drw.div(pos), x = pos - y * side;
            const int site = wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx);
            ore[j] = __ldcg(cache + m.fre_off + site);
            if (!SWEEP) oim[j] = __ldcg(cache + m.fim_off + site);
        }
    }
    float sre = 0.f, sim = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pos = j * kWarp + lane;
        if (pos < npos) {
            if (SWEEP) {
                const float re = ip_site_re(arena, npos, pos, half);
                QMC_ASSERT(ip.newf_off + pos < ip.arena_floats, "new factors inside the arena");
                newf[pos] = re;
                sre += re - ore[j];
            } else {
                const float2 f = ip_site_reim(arena, npos, pos, half);
                sre += f.x - ore[j];
                sim += f.y - oim[j];
            }
        }
    }
    if (!SWEEP) *dim_out = warp_sum(sim);
    dre = warp_sum(sre);
    __syncwarp();
    prof.mark(6);
}

} // namespace qmc

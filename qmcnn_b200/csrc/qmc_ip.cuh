// qmc_ip.cuh - the in-place window evaluator of the persistent sweep kernel (k_sweep_ip).
// Kept out of qmc_device.cuh so that tuning it rebuilds one translation unit only.
// IpPlan (the shared-memory plan) lives in qmc_device.cuh; the design note is above it.
#pragma once
#include "qmc_device.cuh"

namespace qmc {

// Out-of-line epilogue pieces.  The kernel runs 12 warps per SM at unrelated program counters and
// the instruction caches are small (L1.5: 32 KB = 2048 instructions), so code that is executed
// once per proposal must be compact: 64 inlined tanhf per register tile were 18 KB per tile shape
// and the first version of this kernel stalled 34% of its cycles on instruction fetch.  Same
// library tanhf / log2cosh_c as every other path, so the results stay bit-identical.
__device__ __noinline__ float4 ip_tanh4(float4 a) {
    a.x = tanhf(a.x); a.y = tanhf(a.y); a.z = tanhf(a.z); a.w = tanhf(a.w);
    return a;
}

// sum_c Re log 2cosh(theta_c + i theta_{c+half}) of one site (site_factor<false> without the CRBM term)
__device__ __noinline__ float ip_site_re(const float* th, int npos, int pos, int half) {
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
__device__ __noinline__ float2 ip_site_reim(const float* th, int npos, int pos, int half) {
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
        float4* dst = arena4 + (ip.c + dy) * ip.T + (ip.c + dx);
        const float* src = plane + (size_t)site * 4;
        for (int cg = 0; cg < ncg; ++cg) cp_async16(dst + cg * ip.tarea, src + (size_t)cg * n * 4);
    }
}

template <int K, int CIN, int COUT, int ACC, typename OutF, typename MidF>
__device__ __forceinline__ void conv_region_pick_ip(int wbase, int bbase, const float* wsm, const float* tin,
                                                    int tw, int tarea, int side, int lane, OutF out, MidF mid,
                                                    const unsigned short* tab) {
    const int npos = side * side;
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
}

// sites per lane conv_region_pick_ip uses (host mirror, for the single-round check)
__host__ __device__ inline int ip_sites_per_lane(int acc, int cout, int npos) {
    const int pmax = acc / cout;
    if (pmax == 1 || npos <= 32) return 1;
    if (pmax == 2 || npos <= 64) return 2;
    if (pmax == 3 || npos <= 96) return pmax >= 3 ? 3 : 2;
    if (pmax == 4 || npos <= 128) return pmax >= 4 ? 4 : 2;
    if (pmax < 8 || npos <= 192) return pmax >= 6 ? 6 : 4;
    return pmax >= 8 ? 8 : 4;
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
// commit: scatter the staged window of layer l (src: [cog][pos] float4, in shared memory) into the cache plane
__device__ __forceinline__ void ip_scatter_layer(const DevModel& m, const LayerInfo& L, const float4* src, float* cache,
                                                 int side, unsigned mg_side, int ry, int rx, int lane) {
    const int rarea = side * side, ncg = L.coutp >> 2, n = m.n;
    float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
    const FastDiv dside(mg_side, side);
    for (int pos = lane; pos < rarea; pos += kWarp) {
        const int y = dside.div(pos), x = pos - y * side;
        const int site = wrap1(ry + y, m.Ly) * m.Lx + wrap1(rx + x, m.Lx);
        for (int cg = 0; cg < ncg; ++cg) plane4[cg * n + site] = src[cg * rarea + pos];
    }
}

// SWEEP = true: the Metropolis kernel - new windows are also written to `staging` (and start their
// speculative commit copy), only Re of the log-ratio is formed.  SWEEP = false: the local-energy kernel -
// nothing is staged (the cache is read-only), Re and Im are formed (dim).
template <int ACC, int SYNC, bool SWEEP = true>
__device__ __forceinline__ void warp_eval_flip_ip(const DevModel& m, const IpPlan& ip, const float* sp,
                                                  const unsigned* mg, float* arena, float* spt, const int8_t* spins_s,
                                                  const float* __restrict__ cache, float* staging,
                                                  int site_f, int lane, int gid, int gthreads, float& dre,
                                                  float* dim_out = nullptr, const int* tabo = nullptr,
                                                  const unsigned short* tab_s = nullptr) {
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
            int s = spins_s[site];
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
    int stg = 0;
    {   // layer 0 (C_in = 1) reads the spin tile, not the arena
        const LayerInfo& L = m.layer[0];
        const int side = 1 + 2 * p, rarea = side * side, o0 = c - p;
        float4* stg4 = reinterpret_cast<float4*>(staging);
        conv_region_generic(L, m.k, sp, spt, stw, stw * stw, side, side, lane,
                            [&](int pos, int y, int x, int cog, float4 a) {
                                a = ip_tanh4(a);
                                arena4[cog * TA + (y + o0) * T + (x + o0)] = a;
                                if (SWEEP) stg4[cog * rarea + pos] = a;
                            });
        stg += L.coutp * rarea;
        cp_async_wait_all();
        __syncwarp();
    }
    int side = 1 + 2 * p;
    for (int l = 1; l < D; ++l) {
        const LayerInfo& L = m.layer[l];
        const bool last = (l == D - 1);
        if (SYNC >= 2) ip_barrier<SYNC>(gid, gthreads);
        const int hn = (l + 1) * p;                  // half-width of this layer's window
        side += 2 * p;
        const int rarea = side * side;
        const float* tin = arena + (size_t)((c - hn - p) * T + (c - hn - p)) * 4;
        const unsigned short* tab = (tabo && tabo[l] >= 0) ? tab_s + tabo[l] : nullptr;   // conflict-free site deal
        auto sync = [] { __syncwarp(); };
        if (!last) {
            const float* plane = cache + L.act_off;
            const int ncg = L.coutp >> 2, o0 = c - hn;
            ip_gather_frame(arena4, ip, plane, ncg, n, hn + p, p, mg[l + 1], y0, x0, Ly, Lx, lane);   // outer: free now
            float4* stg4 = reinterpret_cast<float4*>(staging + stg);
            auto out = [&](int pos, int y, int x, int cog, float4 a) {
                a = ip_tanh4(a);
                arena4[cog * TA + (y + o0) * T + (x + o0)] = a;
                if (SWEEP) stg4[cog * rarea + pos] = a;
            };
            // all lanes have read the input tile: the inner frame may land now, under the tanh epilogue
            auto mid = [&] {
                __syncwarp();
                ip_gather_frame(arena4, ip, plane, ncg, n, hn, p, mg[l], y0, x0, Ly, Lx, lane);
            };
            if (L.cin == 16 && L.cout == 16)
                conv_region_pick_ip<3, 16, 16, ACC>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, mid, tab);
            else
                conv_region_pick_ip<3, 8, 8, ACC>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, mid, tab);
            stg += L.coutp * rarea;
            cp_async_wait_all();
        } else {
            auto out = [&](int pos, int, int, int cog, float4 a) { arena4[cog * rarea + pos] = a; };
            if (L.cin == 16 && L.cout == 16)
                conv_region_pick_ip<3, 16, 16, ACC>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab);
            else if (L.cin == 16 && L.cout == 8)
                conv_region_pick_ip<3, 16, 8, ACC>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab);
            else
                conv_region_pick_ip<3, 8, 8, ACC>(L.sw_off, L.sb_off, sp, tin, T, TA, side, lane, out, sync, tab);
        }
        __syncwarp();
    }
    if (SYNC >= 2) ip_barrier<SYNC>(gid, gthreads);
    // commit, part 1 (speculative): the staged windows of the first layers start their way from L2 into the
    // free part of the arena now, so that an accepted move finds them in shared memory after the head
    if (SWEEP) {
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
}

} // namespace qmc

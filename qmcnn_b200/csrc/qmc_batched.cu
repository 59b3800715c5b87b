// qmc_batched.cu - layer-synchronous batched evaluation of single-flip proposals.
//
// Same arithmetic as the persistent kernels (qmc_sweep.cu / qmc_energy.cu), different
// decomposition: instead of one warp walking a chain through all layers inside one big
// kernel, every layer is ONE small kernel over all items (item = chain for the sweep,
// (sample, site) for the TFIM local energy):
//
//   k_b_first            RNG / flip site, 0-th layer (spins -> C1 channels) over the 3x3 window
//   k_b_hidden<shape>    layer l: tile = old ring (cache) + new inner window (staging of l-1)
//   k_b_last<shape,mode> last layer + log 2cosh head + log-ratio; then
//                          sweep : accept / reject, commit of the staged windows, sample write-out
//                          energy: exp(log_pop) term per (sample, site)
//
// Why: (1) each kernel is one conv shape with one register tile, so ptxas keeps the
// weights in UNIFORM registers (LDCU from constant memory, FFMA2 R, R.F32, UR.F32x2, R):
// no weight LDS, no weight registers, no shared memory for weights; (2) a warp needs one
// tile (<= 14 KB), not the persistent kernel's 26 KB, so 15-16 warps per SM hide latency
// instead of 7; (3) small windows put 2 or 4 items on one warp.  Measured on B200
// (scripts/proto/layer_proto.cu): 49 TFLOP/s = 69% of the FP32 peak on the 11x11 layer
// against 21 TFLOP/s for the persistent kernel.
//
// Results are bit-identical to the persistent path: same accumulation order, same tanh /
// head code, and the log-ratio is reduced over the same 32 "virtual lanes" butterfly.
#include "qmc_host.h"

namespace qmc {

struct BatchArgs {
    const int8_t* spins_ro;   // energy: samples [N, n]
    int8_t* spins;            // sweep: chains [S, n] (in/out)
    float* cache;             // [n_chains, cache_floats]
    float* stg;               // staging: layer l window of item i at stg + stg_base[l] + i * stg_item[l]
    int* item_site;           // [n_items] flipped site
    float* item_u;            // [n_items] acceptance uniform (sweep)
    const long long* it_base; // device counter: local iteration of node 0 of the current graph / launch
    int n_items;
    int chain0;               // sweep: chain index of item 0 (the chains are split over two streams)
    int mode;                 // 0 sweep, 1 energy (TFIM)
    long long item0;          // energy: global (sample * n + site) index of item 0
    // sweep
    int S, num_flips;
    long long step0;
    const int32_t* flip_pos; const float* uniforms;
    unsigned long long seed; long long chain_id0;
    long long therm_its, its_per_sample, n_sample_slots;
    int8_t* samples; uint8_t* accept_trace; float* logratio_trace;
    unsigned long long* n_accept;
    // energy
    float2* terms;            // [N * n] exp(log_pop) per (sample, site)
    size_t stg_base[QMC_MAX_LAYERS];
    int stg_item[QMC_MAX_LAYERS];
};

__device__ __forceinline__ void item_origin(const DevModel& m, const BatchArgs& a, int item, int& chain,
                                            int& site) {
    if (a.mode == 0) { chain = a.chain0 + item; site = a.item_site[item]; }
    else {
        const long long gi = a.item0 + item;
        chain = (int)(gi / m.n);
        site = (int)(gi - (long long)chain * m.n);
    }
}

// ------------------------------------------------------------------------------------------
// first kernel: flip site + uniform, layer 0 over the (1+2p)^2 window, one warp per item
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_b_first(DevModel m, const float* __restrict__ padded_params, BatchArgs a, int j) {
    extern __shared__ float4 smem4[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int p = m.p, Ly = m.Ly, Lx = m.Lx, n = m.n;
    const int rw = 1 + 2 * p, tw = rw + 2 * p, tarea = tw * tw, rarea = rw * rw;
    float* tile = reinterpret_cast<float*>(smem4) + warp * round4(tarea);
    const long long it = *a.it_base + j;
    const LayerInfo& L = m.layer[0];
    for (int item = blockIdx.x * nwarps + warp; item < a.n_items; item += gridDim.x * nwarps) {
        int chain, site;
        if (a.mode == 0) {
            chain = a.chain0 + item;
            float u;
            if (a.flip_pos) {
                site = a.flip_pos[(size_t)it * a.S + chain];
                u = a.uniforms[(size_t)it * a.S + chain];
            } else {
                const unsigned long long step = (unsigned long long)(a.step0 + it);
                const unsigned long long gchain = (unsigned long long)(a.chain_id0 + chain);
                const uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)gchain,
                                                         (uint32_t)(gchain >> 32)),
                                              make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                site = (int)__umulhi(r.x, (uint32_t)n);
                u = (float)(r.w >> 8) * 5.9604644775390625e-8f;
            }
            if (lane == 0) { a.item_site[item] = site; a.item_u[item] = u; }
        } else {
            item_origin(m, a, item, chain, site);
        }
        const int8_t* sp8 = (a.mode == 0 ? a.spins : a.spins_ro) + (size_t)chain * n;
        const int y0 = site / Lx, x0 = site - y0 * Lx;
        for (int idx = lane; idx < tarea; idx += kWarp) {
            const int ty = idx / tw, tx = idx - ty * tw;
            const int s2 = wrap1(y0 - 2 * p + ty, Ly) * Lx + wrap1(x0 - 2 * p + tx, Lx);
            int s = sp8[s2];
            if (s2 == site) s = -s;
            tile[idx] = (float)s;
        }
        __syncwarp();
        float4* out4 = reinterpret_cast<float4*>(a.stg + a.stg_base[0] + (size_t)item * a.stg_item[0]);
        conv_region_generic(L, m.k, padded_params, tile, tw, tarea, rw, rw, lane,
                            [&](int pos, int, int, int cog, float4 v) {
                                v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
                                out4[cog * rarea + pos] = v;
                            });
        __syncwarp();
    }
}

// tile of layer l for one item: inner window from the staging of layer l-1, ring from the cache
template <int CIN, int LPI>
__device__ __forceinline__ void gather_tile(const DevModel& m, const BatchArgs& a, int l, int item, int chain,
                                            int site, float* tile, int rh, int rw, int sub) {
    const int p = m.p, Ly = m.Ly, Lx = m.Lx, n = m.n;
    const int th = rh + 2 * p, tw = rw + 2 * p, tarea = th * tw;
    const int ih = rh - 2 * p, iw = rw - 2 * p, iarea = ih * iw;   // window of layer l-1
    const int y0 = site / Lx, x0 = site - y0 * Lx;
    const int oy = y0 - (l + 1) * p - p, ox = x0 - (l + 1) * p - p;   // tile origin on the lattice
    const float* plane = a.cache + (size_t)chain * m.cache_floats + m.layer[l - 1].act_off;
    const float* inner = a.stg + a.stg_base[l - 1] + (size_t)item * a.stg_item[l - 1];
    const FastDiv dtw(tw);
    for (int pos = sub; pos < tarea; pos += LPI) {
        const int ty = dtw.div(pos), tx = pos - ty * tw;
        const int iy = ty - 2 * p, ix = tx - 2 * p;
        if (iy >= 0 && iy < ih && ix >= 0 && ix < iw) {
            const float* src = inner + (size_t)(iy * iw + ix) * 4;
#pragma unroll
            for (int cg = 0; cg < CIN / 4; ++cg)
                cp_async16(reinterpret_cast<float4*>(tile) + cg * tarea + pos, src + (size_t)cg * iarea * 4);
        } else {
            const int s2 = wrap1(oy + ty, Ly) * Lx + wrap1(ox + tx, Lx);
#pragma unroll
            for (int cg = 0; cg < CIN / 4; ++cg)
                cp_async16(reinterpret_cast<float4*>(tile) + cg * tarea + pos, plane + (size_t)(cg * n + s2) * 4);
        }
    }
}

// ------------------------------------------------------------------------------------------
// conv layer l (1 <= l <= D-1): pure conv kernel - nothing else lives here on purpose: with
// any extra code (head, commit) in the same kernel ptxas stops keeping the weights in uniform
// registers.  TANH: hidden layers store tanh(z) to the staging of layer l; the last layer
// stores the raw theta to the staging slot D-1.
// ------------------------------------------------------------------------------------------
template <int CIN, int COUT, int P, int IPW, bool TANH>
__global__ void __launch_bounds__(512, 1)
k_b_conv(DevModel m, BatchArgs a, int l, int wbase, int bbase) {
    extern __shared__ float4 smem4[];
    constexpr int LPI = kWarp / IPW;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int p = m.p;
    const int rh = 1 + 2 * (l + 1) * p, rw = rh, th = rh + 2 * p, tarea = th * th, rarea = rh * rw;
    const int tile_floats = tarea * CIN;
    float* tiles = reinterpret_cast<float*>(smem4) + (size_t)warp * IPW * tile_floats;
    const int grp = lane / LPI, sub = lane - grp * LPI;
    const int ntask = (a.n_items + IPW - 1) / IPW;
    for (int task = blockIdx.x * nwarps + warp; task < ntask; task += gridDim.x * nwarps) {
        const int item_raw = task * IPW + grp;
        const int item = item_raw < a.n_items ? item_raw : a.n_items - 1;
        int chain, site;
        item_origin(m, a, item, chain, site);
        gather_tile<CIN, LPI>(m, a, l, item, chain, site, tiles + grp * tile_floats, rh, rw, sub);
        cp_async_wait_all();
        __syncwarp();
        float* stg_l = a.stg + a.stg_base[l];
        const int item0 = task * IPW;
        conv_region_tiled<3, CIN, COUT, P, true, IPW>(
            wbase, bbase, nullptr, tiles, tile_floats, th, tarea, rh, rw, lane,
            [&](int it_in_warp, int pos, int, int, int cog, float4 v) {
                if (item0 + it_in_warp >= a.n_items) return;
                if (TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
                reinterpret_cast<float4*>(stg_l + (size_t)(item0 + it_in_warp) * a.stg_item[l])[cog * rarea + pos] = v;
            });
        __syncwarp();
    }
}

// sum over the 32 "virtual lanes" of the persistent kernel's head loop: virtual lane k owns the
// sites k, k+32, ...; real lane `sub` of an LPI-lane group owns virtual lanes sub + i*LPI.
template <int IPW>
__device__ __forceinline__ float virtual_butterfly(float (&v)[IPW]) {
    constexpr int LPI = kWarp / IPW;
    float s;
    if (IPW == 1) s = v[0];
    else if (IPW == 2) s = v[0] + v[1];                       // k ^ 16
    else s = (v[0] + v[IPW > 2 ? 2 : 0]) + (v[1] + v[IPW > 2 ? 3 : 0]);   // k ^ 16, then k ^ 8
#pragma unroll
    for (int o = LPI / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// ------------------------------------------------------------------------------------------
// head: log 2cosh factors of the staged theta window, log-ratio, then
//   MODE 0 (sweep) : accept / reject, commit of the staged windows, traces, sample write-out
//   MODE 1 (energy): exp(log_pop) term of the (sample, site) item
// ------------------------------------------------------------------------------------------
template <int IPW, int MODE>
__global__ void __launch_bounds__(512)
k_b_head(DevModel m, BatchArgs a, int j) {
    extern __shared__ float4 smem4[];
    constexpr int LPI = kWarp / IPW;
    constexpr bool NEED_IM = MODE == 1;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int p = m.p, D = m.D, l = D - 1, n = m.n, Ly = m.Ly, Lx = m.Lx;
    const int rh = 1 + 2 * (l + 1) * p, rw = rh, rarea = rh * rw;
    const int newf_floats = round4(rarea);
    const int grp = lane / LPI, sub = lane - grp * LPI;
    float* newf = reinterpret_cast<float*>(smem4) + (size_t)(warp * IPW + grp) * newf_floats;
    const int ntask = (a.n_items + IPW - 1) / IPW;
    const long long it = *a.it_base + j;
    const long long step = a.step0 + it;
    const FastDiv drw(rw);
    for (int task = blockIdx.x * nwarps + warp; task < ntask; task += gridDim.x * nwarps) {
        const int item_raw = task * IPW + grp;
        const bool item_on = item_raw < a.n_items;
        const int item = item_on ? item_raw : a.n_items - 1;
        int chain, site;
        item_origin(m, a, item, chain, site);
        const float* cache = a.cache + (size_t)chain * m.cache_floats;
        const float* theta = a.stg + a.stg_base[l] + (size_t)item * a.stg_item[l];
        const int y0 = site / Lx, x0 = site - y0 * Lx;
        const int ry = y0 - (l + 1) * p, rx = x0 - (l + 1) * p;
        float vre[IPW], vim[IPW];
#pragma unroll
        for (int i = 0; i < IPW; ++i) { vre[i] = 0.f; vim[i] = 0.f; }
        for (int base = 0; base < rarea; base += kWarp) {
#pragma unroll
            for (int i = 0; i < IPW; ++i) {
                const int pos = base + i * LPI + sub;        // virtual lane = i * LPI + sub
                if (pos < rarea) {
                    const int y = drw.div(pos), x = pos - y * rw;
                    const int s2 = wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx);
                    float re, im;
                    site_factor<NEED_IM>(m, nullptr, theta, rarea, pos, 0.f, re, im);
                    newf[pos] = re;
                    vre[i] += re - __ldcg(cache + m.fre_off + s2);
                    if (NEED_IM) vim[i] += im - __ldcg(cache + m.fim_off + s2);
                }
            }
        }
        const float dre = virtual_butterfly<IPW>(vre);
        const float dim = NEED_IM ? virtual_butterfly<IPW>(vim) : 0.f;
        __syncwarp();
        if (MODE == 1) {
            if (item_on && sub == 0) {
                const float amp = expf(dre);
                float sn, cn;
                sincosf(dim, &sn, &cn);
                a.terms[a.item0 + item] = make_float2(amp * cn, amp * sn);
            }
        } else {
            const float u = a.item_u[item];
            const float amp = expf(dre);
            const bool accept = amp * amp > u;               // strict, sampler.py:125
            if (item_on && accept) {
                float* cache_w = a.cache + (size_t)chain * m.cache_floats;
                for (int ll = 0; ll < D - 1; ++ll) {
                    const int ch = 1 + 2 * (ll + 1) * p, carea = ch * ch;
                    const int ncg = m.layer[ll].coutp >> 2;
                    const int cy = y0 - (ll + 1) * p, cx = x0 - (ll + 1) * p;
                    const float* src = a.stg + a.stg_base[ll] + (size_t)item * a.stg_item[ll];
                    float4* plane4 = reinterpret_cast<float4*>(cache_w + m.layer[ll].act_off);
                    const FastDiv dcw(ch);
                    for (int pos = sub; pos < carea; pos += LPI) {
                        const int y = dcw.div(pos), x = pos - y * ch;
                        const int s2 = wrap1(cy + y, Ly) * Lx + wrap1(cx + x, Lx);
                        for (int cg = 0; cg < ncg; cg += 4) {            // up to 4 loads in flight
                            float4 v[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (cg + q < ncg) v[q] = ldcg4(src + (size_t)((cg + q) * carea + pos) * 4);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (cg + q < ncg) plane4[(cg + q) * n + s2] = v[q];
                        }
                    }
                }
                for (int pos = sub; pos < rarea; pos += LPI) {
                    const int y = drw.div(pos), x = pos - y * rw;
                    cache_w[m.fre_off + wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx)] = newf[pos];
                }
                if (sub == 0) a.spins[(size_t)chain * n + site] = -a.spins[(size_t)chain * n + site];
            }
            if (item_on && sub == 0) {
                if (accept && a.n_accept) atomicAdd(a.n_accept, 1ULL);
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = dre;
            }
            __syncwarp();
            if (a.samples && step >= a.therm_its && (step - a.therm_its) % a.its_per_sample == 0) {
                const long long js = (step - a.therm_its) / a.its_per_sample;
                if (js < a.n_sample_slots && item_on) {
                    int8_t* dst = a.samples + ((size_t)js * a.S + chain) * n;
                    const int8_t* srcs = a.spins + (size_t)chain * n;
                    for (int i = sub; i < n; i += LPI) dst[i] = srcs[i];
                }
            }
        }
        __syncwarp();
    }
}

__global__ void k_b_advance(long long* it_base, long long by) { *it_base += by; }
__global__ void k_b_set(long long* it_base, long long v) { *it_base = v; }

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TileChoice { int P, IPW; };

// (P, IPW) with the fewest FMA-pipe cycles per item for a window of npos sites.  Only the
// 64-accumulator register tiles are offered (P = 4 at 16 output channels, P = 8 at 8): with
// smaller tiles ptxas keeps the weights in vector registers for some instances (checked on the
// SASS of every instance: FFMA2 with a UR operand).
static TileChoice choose_tile(int npos, int cout, int tile_bytes) {
    // ordered by items per warp: at equal FMA cost the fewer-items tile wins (one tile per warp,
    // one round, more warps per SM)
    static const TileChoice cands16[] = {{4, 1}, {4, 2}, {4, 4}};
    static const TileChoice cands8[] = {{8, 1}, {8, 2}, {8, 4}};
    const TileChoice* c = cout == 16 ? cands16 : cands8;
    TileChoice best = c[0];
    double best_cost = 1e30;
    for (int i = 0; i < 3; ++i) {
        // several items per warp only while >= 14 warps still fit one SM's shared memory
        if (c[i].IPW > 1 && (size_t)c[i].IPW * tile_bytes * 14 > 220 * 1024) continue;
        const int lpi = 32 / c[i].IPW, G = (npos + c[i].P - 1) / c[i].P;
        const int rounds = (G + lpi - 1) / lpi;
        const double cost = (double)rounds * c[i].P / c[i].IPW;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = c[i]; }
    }
    return best;
}

bool batched_supported(const qmc_handle* h) {
    const DevModel& m = h->m;
    if (!h->allow_tiled || !m.use_const || m.kind != QMC_MODEL_DCRBM || m.k != 3 || m.D < 2) return false;
    if (m.r > m.Ly || m.r > m.Lx) return false;
    for (int l = 1; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        const bool ok = (L.cin == 16 && L.cout == 16) || (L.cin == 16 && L.cout == 8) || (L.cin == 8 && L.cout == 8);
        if (!ok) return false;
    }
    return true;
}

// staged windows of all D layers (hidden: tanh activations; last: theta)
size_t batched_staging_floats(const qmc_handle* h, int n_items) {
    const DevModel& m = h->m;
    size_t f = 0;
    for (int l = 0; l < m.D; ++l) {
        const int side = 1 + 2 * (l + 1) * m.p;
        f += (size_t)n_items * m.layer[l].coutp * side * side;
    }
    return f;
}

// scratch after the staging: item_site (int), item_u (float), it_base (2 words), padding
size_t batched_scratch_floats(int n_items) { return (((size_t)2 * n_items + 8) + 3) & ~(size_t)3; }

// warps per CTA / grid for `ntask` warp tasks needing per_warp_bytes of shared memory each
static cudaError_t layer_geometry(const qmc_handle* h, size_t per_warp_bytes, int ntask, int max_warps,
                                  int& grid, int& warps, size_t& smem) {
    int wmax = max_warps;
    while (wmax > 1 && per_warp_bytes * wmax > h->max_smem) --wmax;
    if (per_warp_bytes * wmax > h->max_smem) return cudaErrorInvalidValue;
    // no more warps than a single wave needs; otherwise the count with the smallest idle tail
    const long long slots_needed = ((long long)ntask + h->num_sms - 1) / h->num_sms;
    if (slots_needed <= wmax) warps = (int)(slots_needed > 0 ? slots_needed : 1);
    else {
        double best = -1;
        warps = wmax;
        for (int w = wmax; w >= (wmax + 1) / 2; --w) {
            const long long slots = (long long)h->num_sms * w, waves = (ntask + slots - 1) / slots;
            const double eff = (double)ntask / (double)(waves * slots);
            if (eff > best + 1e-9) { best = eff; warps = w; }
        }
    }
    smem = per_warp_bytes * warps;
    long long ctas = ((long long)ntask + warps - 1) / warps;
    grid = (int)(ctas < h->num_sms ? ctas : h->num_sms);
    return cudaSuccess;
}

template <typename Kern>
static cudaError_t launch_conv_kernel(Kern kern, const qmc_handle* h, const BatchArgs& a, int l, int ipw,
                                      int tile_floats, cudaStream_t st) {
    int grid, warps; size_t smem;
    cudaError_t e = layer_geometry(h, (size_t)ipw * tile_floats * 4, (a.n_items + ipw - 1) / ipw, 16, grid, warps, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const LayerInfo& L = h->m.layer[l];
    kern<<<grid, warps * 32, smem, st>>>(h->m, a, l, L.sw_off, L.sb_off);
    return cudaGetLastError();
}

#define QMC_CONV_CASE(CI, CO, PP, IP)                                                            \
    if (L.cin == CI && L.cout == CO && tc.P == PP && tc.IPW == IP)                               \
        return last ? launch_conv_kernel(k_b_conv<CI, CO, PP, IP, false>, h, a, l, IP, tarea * CI, st) \
                    : launch_conv_kernel(k_b_conv<CI, CO, PP, IP, true>, h, a, l, IP, tarea * CI, st);
#define QMC_CONV_SHAPE16(CI, CO) \
    QMC_CONV_CASE(CI, CO, 4, 4) QMC_CONV_CASE(CI, CO, 4, 2) QMC_CONV_CASE(CI, CO, 4, 1)
#define QMC_CONV_SHAPE8(CI, CO) \
    QMC_CONV_CASE(CI, CO, 8, 4) QMC_CONV_CASE(CI, CO, 8, 2) QMC_CONV_CASE(CI, CO, 8, 1)

static cudaError_t launch_conv(const qmc_handle* h, const BatchArgs& a, int l, cudaStream_t st) {
    const DevModel& m = h->m;
    const LayerInfo& L = m.layer[l];
    const bool last = l == m.D - 1;
    const int side = 1 + 2 * (l + 1) * m.p, tside = side + 2 * m.p, tarea = tside * tside;
    const TileChoice tc = choose_tile(side * side, L.cout, tarea * L.cin * 4);
    QMC_CONV_SHAPE16(16, 16)
    QMC_CONV_SHAPE8(16, 8)
    QMC_CONV_SHAPE8(8, 8)
    return cudaErrorInvalidValue;
}

static cudaError_t launch_head(const qmc_handle* h, const BatchArgs& a, int j, cudaStream_t st) {
    const DevModel& m = h->m;
    const int side = 1 + 2 * m.D * m.p, rarea = side * side;
    const int ipw = rarea > 64 ? 1 : (rarea > 32 ? 2 : 4);
    int grid, warps; size_t smem;
    cudaError_t e = layer_geometry(h, (size_t)ipw * round4(rarea) * 4, (a.n_items + ipw - 1) / ipw, 16, grid, warps, smem);
    if (e != cudaSuccess) return e;
#define QMC_HEAD(IP, MD) k_b_head<IP, MD><<<grid, warps * 32, smem, st>>>(m, a, j)
    if (a.mode == 0) { if (ipw == 1) QMC_HEAD(1, 0); else if (ipw == 2) QMC_HEAD(2, 0); else QMC_HEAD(4, 0); }
    else { if (ipw == 1) QMC_HEAD(1, 1); else if (ipw == 2) QMC_HEAD(2, 1); else QMC_HEAD(4, 1); }
#undef QMC_HEAD
    return cudaGetLastError();
}

static cudaError_t launch_first(const qmc_handle* h, const BatchArgs& a, int j, cudaStream_t st) {
    const DevModel& m = h->m;
    const int tw = 1 + 4 * m.p, warps = 8;
    const size_t smem = (size_t)warps * round4(tw * tw) * 4;
    long long ctas = ((long long)a.n_items + warps - 1) / warps;
    const int grid = (int)(ctas < h->num_sms * 8 ? ctas : h->num_sms * 8);
    k_b_first<<<grid, warps * 32, smem, st>>>(m, h->d_params_padded, a, j);
    return cudaGetLastError();
}

static void fill_staging(const qmc_handle* h, BatchArgs& a, float* stg, int n_items) {
    const DevModel& m = h->m;
    size_t off = 0;
    for (int l = 0; l < m.D; ++l) {
        const int side = 1 + 2 * (l + 1) * m.p;
        a.stg_base[l] = off;
        a.stg_item[l] = m.layer[l].coutp * side * side;
        off += (size_t)n_items * a.stg_item[l];
    }
    a.stg = stg;
    int* ip = reinterpret_cast<int*>(stg + off);
    a.item_site = ip;
    a.item_u = reinterpret_cast<float*>(ip + n_items);
    a.it_base = reinterpret_cast<const long long*>(ip + 2 * n_items + (((size_t)(ip + 2 * n_items) & 7) ? 1 : 0));
}

static cudaError_t upload_const(const qmc_handle* h, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_params, h->d_params_padded, (size_t)h->m.smem_param_floats * 4, 0,
                                   cudaMemcpyDeviceToDevice, st);
}

static cudaError_t enqueue_step(const qmc_handle* h, const BatchArgs& a, int j, cudaStream_t st, bool count = true) {
    if (count) g_launches += h->m.D + 1;
    cudaError_t e = launch_first(h, a, j, st);
    for (int l = 1; e == cudaSuccess && l < h->m.D; ++l) e = launch_conv(h, a, l, st);
    if (e == cudaSuccess) e = launch_head(h, a, j, st);
    return e;
}

constexpr int kGraphSteps = 32;   // Metropolis steps captured per CUDA graph

// one half of the chains: capture kGraphSteps steps into a graph and replay it, remainder direct
static cudaError_t run_sweep_part(const qmc_handle* h, BatchArgs a, long long n_steps, cudaStream_t st,
                                  std::string& err) {
    long long* it_base = const_cast<long long*>(a.it_base);
    k_b_set<<<1, 1, 0, st>>>(it_base, 0);
    cudaError_t e = cudaGetLastError();
    long long done = 0;
    if (e == cudaSuccess && n_steps >= 2 * kGraphSteps) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            cudaError_t e2 = cudaSuccess;
            for (int j = 0; j < kGraphSteps && e2 == cudaSuccess; ++j) e2 = enqueue_step(h, a, j, st, false);
            if (e2 == cudaSuccess) k_b_advance<<<1, 1, 0, st>>>(it_base, kGraphSteps);
            e = cudaStreamEndCapture(st, &graph);
            if (e2 != cudaSuccess) e = e2;
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (e == cudaSuccess) {
            const long long nblocks = n_steps / kGraphSteps;
            for (long long b = 0; b < nblocks && e == cudaSuccess; ++b) e = cudaGraphLaunch(exec, st);
            g_launches += (unsigned long long)nblocks * (kGraphSteps * (h->m.D + 1) + 1);
            done = nblocks * kGraphSteps;
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) { err = "sweep (batched): CUDA graph capture/launch failed"; return e; }
    }
    for (long long i = done; i < n_steps && e == cudaSuccess; ++i) e = enqueue_step(h, a, (int)(i - done), st);
    return e;
}

cudaError_t launch_sweep_batched(const qmc_handle* h, const SweepArgs& s, cudaStream_t caller, std::string& err) {
    // The step sequence runs on the handle's own non-blocking streams so that it can be captured
    // into CUDA graphs even when the caller's stream is the legacy default stream; events keep it
    // ordered after / before the caller's stream.  The chains are split over TWO streams: all warps
    // of one per-layer kernel are in the same phase (gather, FMA, store), so two independent
    // kernel sequences let the memory phase of one half overlap the FMA phase of the other.
    cudaError_t e = cudaEventRecord(h->ev_in, caller);
    const int nparts = s.S >= 512 ? 2 : 1;
    for (int part = 0; part < nparts && e == cudaSuccess; ++part) e = cudaStreamWaitEvent(h->side_stream[part], h->ev_in, 0);
    if (e != cudaSuccess) return e;
    e = upload_const(h, h->side_stream[0]);
    if (e == cudaSuccess && nparts > 1) {   // the constant upload must precede the other stream's kernels too
        e = cudaEventRecord(h->ev_mid, h->side_stream[0]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->side_stream[1], h->ev_mid, 0);
    }
    const int half = nparts > 1 ? (s.S + 1) / 2 : s.S;
    float* stg = s.staging;
    for (int part = 0; part < nparts && e == cudaSuccess; ++part) {
        const int c0 = part * half, cnt = part == 0 ? half : s.S - half;
        BatchArgs a{};
        a.spins = s.spins; a.spins_ro = s.spins; a.cache = s.cache; a.n_items = cnt; a.chain0 = c0; a.mode = 0;
        a.item0 = 0; a.S = s.S; a.num_flips = 1; a.step0 = s.step0; a.flip_pos = s.flip_pos; a.uniforms = s.uniforms;
        a.seed = s.seed; a.chain_id0 = s.chain_id0; a.therm_its = s.therm_its; a.its_per_sample = s.its_per_sample;
        a.n_sample_slots = s.n_sample_slots; a.samples = s.samples; a.accept_trace = s.accept_trace;
        a.logratio_trace = s.logratio_trace; a.n_accept = s.n_accept; a.terms = nullptr;
        fill_staging(h, a, stg, cnt);
        stg += batched_staging_floats(h, cnt) + batched_scratch_floats(cnt);
        e = run_sweep_part(h, a, s.n_steps, h->side_stream[part], err);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_out[part], h->side_stream[part]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(caller, h->ev_out[part], 0);
    }
    return e;
}

// TFIM local energy: terms[s * n + i] = exp(log_pop_i) for every (sample, site), in chunks
cudaError_t launch_energy_batched(const qmc_handle* h, const int8_t* spins, int N, const float* cache,
                                  float* scratch, int chunk_items, float2* terms, cudaStream_t st,
                                  std::string& err) {
    (void)err;
    cudaError_t e = upload_const(h, st);
    if (e != cudaSuccess) return e;
    const long long total = (long long)N * h->m.n;
    for (long long i0 = 0; i0 < total && e == cudaSuccess; i0 += chunk_items) {
        BatchArgs a{};
        a.spins_ro = spins; a.spins = nullptr; a.cache = const_cast<float*>(cache);
        a.n_items = (int)((total - i0) < chunk_items ? (total - i0) : chunk_items);
        a.mode = 1; a.chain0 = 0; a.item0 = i0; a.S = N; a.terms = terms;
        fill_staging(h, a, scratch, chunk_items);
        if (i0 == 0) k_b_set<<<1, 1, 0, st>>>(const_cast<long long*>(a.it_base), 0);
        e = enqueue_step(h, a, 0, st);
    }
    return e;
}

} // namespace qmc

// qmc_sweep_ip.cu - the persistent Metropolis kernel with the in-place evaluator.
// Same per-proposal protocol as k_sweep (qmc_sweep.cu; sampler.py:104-155) and bit-identical
// results; the difference is warp_eval_flip_ip: one tile arena per warp instead of two ping-pong
// tiles, so that ~12 instead of 7 chains are resident per SM at C3 (3 warps per scheduler
// instead of 1.75 - the classic kernel issues on only 44% of cycles because each scheduler has
// fewer than two warps to choose from).  Single-flip proposals of deep k = 3 models whose
// windows fit one round of the register tile; everything else runs the classic kernel.
#include <cstdio>
#include <vector>
#include "qmc_host.h"
#include "qmc_ip.cuh"

namespace qmc {

constexpr int kIpMaxWarps = 12;       // 384 threads -> 168 registers per thread
constexpr int kIpAcc = 64;            // accumulators per lane

// Time slicing (host side, launch_sweep_ip): the S x n_steps proposals are cut into tasks
// (chunk of `chunk_len` consecutive steps, chain), numbered chunk-major, and every launch hands
// ONE task to every warp slot.  A chain's chunks are in different launches (S >= slots), so the
// stream orders them; every launch is exactly one full wave, so there is no idle tail whatever S
// is (4096 chains on 148 x 12 slots would otherwise be 2.3 waves = 77% efficiency).
struct IpSlice { long long task0, n_tasks, chunk_len; int group_warps, stagger; };

// This file is compiled three times (Makefile): QMC_IP_PART 0 = the host side + k_sweep_ip<3> (phase groups, the default),
// 1 = k_sweep_ip<0> (free-running, diagnosis), 2 = k_energy_ip.  Each kernel instantiates every register-tile shape of
// the evaluator and takes about a minute of ptxas; three objects build in parallel.
#ifndef QMC_IP_PART
#define QMC_IP_PART 0
#endif
cudaError_t ip_launch_sweep_groups(int ctas, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                                   const SweepArgs& a, const IpPlan& ip, const IpSlice& sl, const site_t* tab);
cudaError_t ip_launch_sweep_free(int ctas, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                                 const SweepArgs& a, const IpPlan& ip, const IpSlice& sl, const site_t* tab);
cudaError_t ip_launch_energy(int grid, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                             const int8_t* spins, int N, const float* cache, float2* partial, int nchunks, const IpPlan& ip,
                             int group_warps, const site_t* tab);

#if QMC_IP_PROFILE
static __device__ unsigned long long g_ip_prof[kIpProfPhases + 1 + 12];  // cycles per phase summed over warps, [8] = proposals,
                                                                   // [9 + w] = task duration of warp w summed over CTAs
#endif

// per-CTA words after the parameter block: division magics, site-table offsets, then the site tables (uint16)
constexpr int kIpCtaWords = 2 * QMC_MAX_LAYERS;
__host__ __device__ inline size_t ip_cta_bytes(const IpPlan& ip) { return (size_t)kIpCtaWords * 4 + (size_t)ip.tab_entries * sizeof(site_t); }

// shared-memory images of the division magics, the table offsets and the site tables; returns the first per-warp byte
__device__ __forceinline__ char* ip_cta_setup(float* after_params, const IpPlan& ip, const site_t* __restrict__ tab_g,
                                              unsigned*& mg, int*& tabo, site_t*& tab_s) {
    mg = reinterpret_cast<unsigned*>(after_params);
    tabo = reinterpret_cast<int*>(mg + QMC_MAX_LAYERS);
    tab_s = reinterpret_cast<site_t*>(mg + kIpCtaWords);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < QMC_MAX_LAYERS; ++j) { mg[j] = ip.mgW[j]; tabo[j] = tab_g ? ip.tab_off[j] : -1; }
    }
    if (tab_g)
        for (int i = threadIdx.x; i < ip.tab_entries; i += blockDim.x) tab_s[i] = tab_g[i];
    __syncthreads();
    return reinterpret_cast<char*>(tab_s + ip.tab_entries);
}

// SYNC: 0 = warps run free; 1 = phase-group barrier per proposal only; 2 = CTA barrier per layer; 3 = per
// layer among the four warps of a phase group (ip_barrier).  With barriers the warps of a group
// stay in the same phase of the proposal, so the SM's instruction working set is a few loop
// bodies instead of twelve (the free-running version is instruction-fetch bound: 93% of the
// GPC instruction-cache request rate, profiles/r01_summary.md).  Warps without a task or past
// their chunk's end shadow a valid chain without writing anything, to keep barrier counts equal.
#if QMC_IP_PART != 2
template <int SYNC>
__global__ void __launch_bounds__(kIpMaxWarps * 32, 1)
k_sweep_ip(DevModel m, const float* __restrict__ params, SweepArgs a, IpPlan ip, IpSlice sl,
           const site_t* __restrict__ tab_g) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
#if QMC_IP_PROFILE
    if (threadIdx.x < 4) s_ip_conv[threadIdx.x] = 0;
#endif
    load_params_to_smem(m, params, smem_f);
    const float* sp = smem_f;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // division magics W_j = 2(j+2)p + 1 in shared memory: indexing the kernel-parameter array by layer would
    // make ptxas keep a local-memory copy, and those loads miss L1 (28 KB next to 220 KB of shared memory)
    unsigned* mg; int* tabo; site_t* tab_s;
    char* wmem = ip_cta_setup(smem_f + m.smem_param_floats, ip, tab_g, mg, tabo, tab_s) + (size_t)warp * ip.per_warp_bytes;
    float* arena = reinterpret_cast<float*>(wmem);
    float* spt = arena + ip.arena_floats;
    unsigned* spins_s = reinterpret_cast<unsigned*>(spt + ip.spt_floats);     // one bit per spin (ip_spin)

    const int n = m.n, p = m.p, Ly = m.Ly, Lx = m.Lx, D = m.D;
    const int slot = blockIdx.x * nwarps + warp;
    float* staging = a.staging + (size_t)slot * ip.staging_floats;
    const float* newf = arena + ip.newf_off;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const int lside = 1 + 2 * D * p;                       // window of the last layer
    unsigned long long accepted = 0;
    // phase group of this warp (SYNC == 3): warps 4g .. 4g+3, named barrier 1 + g
    const int gw = sl.group_warps, g = warp / gw;
    const int gid = 1 + g;
    const int gthreads = 32 * min(gw, nwarps - g * gw);

    const long long task = sl.task0 + slot;
    const bool has_task = task < sl.n_tasks;
    if (!SYNC && !has_task) return;
    const long long tt = has_task ? task : sl.n_tasks - 1;
    const long long chunk = tt / a.S;
    const int chain = (int)(tt - chunk * a.S);
    const long long it0 = chunk * sl.chunk_len;
    const long long it1 = min(it0 + sl.chunk_len, a.n_steps);
    {
        int8_t* gspins = a.spins + (size_t)chain * n;
        float* cache = a.cache + (size_t)chain * m.cache_floats;
        ip_pack_spins(spins_s, gspins, n, lane);
        __syncwarp();
        const unsigned long long gchain = (unsigned long long)(a.chain_id0 + chain);
        // (flip site, uniform) of iteration `it`: Philox-4x32-10 keyed by (seed; step, global chain), or fed in
        auto draw = [&](long long it, int& f, float& uu) {
            if (a.flip_pos) {
                f = a.flip_pos[(size_t)it * a.S + chain];
                uu = a.uniforms[(size_t)it * a.S + chain];
            } else {
                const long long step = a.step0 + it;
                const uint4 r = philox4x32_10(
                    make_uint4((uint32_t)step, (uint32_t)((unsigned long long)step >> 32),
                               (uint32_t)gchain, (uint32_t)(gchain >> 32)), key);
                f = (int)__umulhi(r.x, (uint32_t)n);
                uu = (float)(r.w >> 8) * 5.9604644775390625e-8f;   // 2^-24
            }
            f = __shfl_sync(0xffffffffu, f, 0);
            uu = __shfl_sync(0xffffffffu, uu, 0);
        };
        int f_next = 0;
        float u_next = 0.f;
        // Phase groups start g * stagger cycles apart.  All warps run the same code on equally sized windows, so groups
        // that start together stay in lockstep: the three warps of a scheduler are then in the conv loops at the same
        // time (73% of the tap iterations see all three there, profiles/r02_summary.md) and in the latency-bound
        // epilogue / head / commit at the same time.  An offset de-phases them for the rest of the launch.
        if (sl.stagger > 0) {
            const long long t_go = clock64() + (long long)g * sl.stagger;
            while (clock64() < t_go) __nanosleep(200);
        }
        IpProf prof;
#if QMC_IP_PROFILE
        for (int i = 0; i < kIpProfPhases; ++i) prof.acc[i] = 0;
#endif
        draw(it0 < it1 ? it0 : it1 - 1, f_next, u_next);
        prof.start();
#if QMC_IP_PROFILE
        const long long t_task0 = clock64();
#endif
        for (long long itx = it0; itx < it0 + sl.chunk_len; ++itx) {
            const bool active = has_task && itx < it1;
            if (!SYNC && !active) break;
            const long long it = active ? itx : it1 - 1;      // shadow steps re-evaluate the last one
            const long long step = a.step0 + it;
            const int f0 = f_next;
            const float u = u_next;
            if (itx + 1 < it1) draw(itx + 1, f_next, u_next);   // one step ahead: the loads / Philox rounds overlap the evaluation
            float dre;
            cp_async_wait_all();                              // (a rejected move's speculative commit copy)
            ip_barrier<SYNC>(gid, gthreads);
            prof.mark(0);
            QMC_ASSERT(f0 >= 0 && f0 < n, "flip site on the lattice");
            warp_eval_flip_ip<kIpAcc, SYNC>(m, ip, sp, mg, arena, spt, spins_s, cache, staging, f0, lane, gid, gthreads, dre,
                                            nullptr, tabo, tab_s, prof);
            const float amp = expf(dre);                      // |exp(z)| = exp(Re z)
            const bool accept = __shfl_sync(0xffffffffu, (int)(amp * amp > u), 0) != 0;   // strict, sampler.py:125
            if (SYNC && !active) { prof.mark(7); continue; }  // shadow: nothing is written
            if (accept) {
                // commit: new hidden activations, new factors, the spin.  The staged windows travel L2 -> shared
                // memory in bulk (linear cp.async: the first layers went ahead during the head, the rest follow in
                // as few batches as fit the arena below the new factors) and are scattered from there, instead of
                // one L2 round trip per 2 sites x 2 channel groups (that loop was 10% of the kernel's time).
                const int y0 = f0 / Lx, x0 = f0 - y0 * Lx;
                cp_async_wait_all();
                __syncwarp();
                int stg = 0, side = 1, l = 0;
                for (; l < ip.spec_layers; ++l) {
                    const LayerInfo& L = m.layer[l];
                    side += 2 * p;
                    ip_scatter_layer(m, L, reinterpret_cast<const float4*>(arena + ip.spec_off + stg), cache, side,
                                     l ? mg[l - 1] : ip.mg2p1, y0 - (l + 1) * p, x0 - (l + 1) * p, lane,
                                     l ? ip_tile(kIpAcc, L.cout, side * side).cs : 1);
                    stg += L.coutp * side * side;
                }
                while (l < D - 1) {
                    // batch [l, l1): consecutive layers whose staged windows fit below the new factors
                    int l1 = l, fl = 0, sd = side;
                    while (l1 < D - 1) {
                        const int s2 = sd + 2 * p, add = m.layer[l1].coutp * s2 * s2;
                        if (l1 > l && fl + add > ip.newf_off) break;
                        fl += add; sd = s2; ++l1;
                    }
                    __syncwarp();                                  // earlier scatter reads of the arena are done
                    QMC_ASSERT(fl <= ip.newf_off && stg + fl <= ip.staging_floats, "commit batch inside arena and staging");
                    for (int i = lane * 4; i < fl; i += kWarp * 4) cp_async16(arena + i, staging + stg + i);
                    cp_async_wait_all();
                    __syncwarp();
                    int off = 0;
                    for (; l < l1; ++l) {
                        const LayerInfo& L = m.layer[l];
                        side += 2 * p;
                        ip_scatter_layer(m, L, reinterpret_cast<const float4*>(arena + off), cache, side,
                                         l ? mg[l - 1] : ip.mg2p1, y0 - (l + 1) * p, x0 - (l + 1) * p, lane,
                                         l ? ip_tile(kIpAcc, L.cout, side * side).cs : 1);
                        off += L.coutp * side * side;
                    }
                    stg += fl;
                }
                const int ry = y0 - D * p, rx = x0 - D * p;
                const FastDiv dls(mg[D - 2], lside);
                for (int pos = lane; pos < lside * lside; pos += kWarp) {
                    const int y = dls.div(pos), x = pos - y * lside;
                    cache[m.fre_off + wrap1(ry + y, Ly) * Lx + wrap1(rx + x, Lx)] = newf[pos];
                }
                if (lane == 0) spins_s[f0 >> 5] ^= 1u << (f0 & 31);
                __syncwarp();
                ++accepted;
            }
            if (lane == 0) {
                if (a.accept_trace) a.accept_trace[(size_t)it * a.S + chain] = accept ? 1 : 0;
                if (a.logratio_trace) a.logratio_trace[(size_t)it * a.S + chain] = dre;
            }
            // sample write-out AFTER the update (sampler.py:135-152)
            if (a.samples && step >= a.therm_its && (step - a.therm_its) % a.its_per_sample == 0) {
                const long long j = (step - a.therm_its) / a.its_per_sample;
                if (j < a.n_sample_slots) {
                    int8_t* dst = a.samples + ((size_t)j * a.S + chain) * n;
                    ip_unpack_spins(dst, spins_s, n, lane);
                }
            }
            prof.mark(7);
        }
#if QMC_IP_PROFILE
        if (lane == 0 && has_task) {
            for (int i = 0; i < kIpProfPhases; ++i) atomicAdd(&g_ip_prof[i], (unsigned long long)prof.acc[i]);
            atomicAdd(&g_ip_prof[kIpProfPhases], (unsigned long long)(it1 - it0));
            if (warp < 12) atomicAdd(&g_ip_prof[kIpProfPhases + 1 + warp], (unsigned long long)(clock64() - t_task0));
        }
#endif
        if (has_task)
            ip_unpack_spins(gspins, spins_s, n, lane);
        __syncwarp();
    }
    if (a.n_accept && lane == 0 && accepted) atomicAdd(a.n_accept, accepted);
}

#if QMC_IP_PART == 0
cudaError_t ip_launch_sweep_groups(int ctas, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                                   const SweepArgs& a, const IpPlan& ip, const IpSlice& sl, const site_t* tab) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep_ip<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_sweep_ip<3><<<ctas, threads, smem, st>>>(m, params, a, ip, sl, tab);
    return cudaGetLastError();
}
#else
cudaError_t ip_launch_sweep_free(int ctas, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                                 const SweepArgs& a, const IpPlan& ip, const IpSlice& sl, const site_t* tab) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep_ip<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_sweep_ip<0><<<ctas, threads, smem, st>>>(m, params, a, ip, sl, tab);
    return cudaGetLastError();
}
#endif
#endif // QMC_IP_PART != 2

#if QMC_IP_PART == 0
// Is the model inside the in-place evaluator's coverage, and what does a warp need?
IpPlan ip_plan(const qmc_handle* h) {
    IpPlan ip{};
    const DevModel& m = h->m;
    if (!h->allow_tiled || m.kind != QMC_MODEL_DCRBM || m.D < 2 || m.k != 3) return ip;
    if (1 + 2 * m.D * m.p > m.Ly || 1 + 2 * m.D * m.p > m.Lx) return ip;      // box_supported(1, 1)
    const int p = m.p;
    int cmax = 0, staging = 0;
    for (int l = 0; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        const int side = 1 + 2 * (l + 1) * p, npos = side * side;
        if (l >= 1) {
            const bool hidden_ok = (L.cin == 16 && L.cout == 16) || (L.cin == 8 && L.cout == 8);
            const bool last_ok = hidden_ok || (L.cin == 16 && L.cout == 8);
            if (l < m.D - 1 ? !hidden_ok : !last_ok) return ip;
            if (ip_tile(kIpAcc, L.cout, npos).p == 0) return ip;              // must be ONE round of the register tile
        }
        if (l < m.D - 1) {
            if (L.coutp > cmax) cmax = L.coutp;
            staging += L.coutp * npos;
        }
    }
    if ((1 + 2 * m.D * p) * (1 + 2 * m.D * p) > 8 * kWarp) return ip;      // head: at most 8 sites per lane
    ip.T = 1 + 2 * (m.D + 1) * p;
    if (ip.T * ip.T > kIpPlane) return ip;                                    // (cannot happen inside the coverage above)
    ip.tarea = kIpPlane;
    ip.c = (m.D + 1) * p;
    const int lside = 1 + 2 * m.D * p, theta = m.layer[m.D - 1].coutp * lside * lside;
    int arena = ip.tarea * cmax;
    const int nf = round4(lside * lside);
    if (round4(theta) + nf > arena) arena = round4(theta) + nf;
    ip.arena_floats = round4(arena);
    ip.newf_off = ip.arena_floats - nf;                          // new factors at the end of the arena
    // a single layer's staged window must fit below them (commit batches), else not covered
    for (int l = 0; l < m.D - 1; ++l) {
        const int side = 1 + 2 * (l + 1) * p;
        if (m.layer[l].coutp * side * side > ip.newf_off) return ip;
    }
    ip.spec_off = round4(theta);
    ip.spec_layers = 0;
    ip.spec_floats = 0;
    for (int l = 0; l < m.D - 1; ++l) {
        const int side = 1 + 2 * (l + 1) * p, add = m.layer[l].coutp * side * side;
        if (ip.spec_off + ip.spec_floats + add > ip.newf_off) break;
        ip.spec_floats += add;
        ++ip.spec_layers;
    }
    ip.spt_floats = round4((1 + 4 * p) * (1 + 4 * p));
    ip.staging_floats = round4(staging);
    ip.spins_bytes = (ip_spin_words(m.n) * 4 + 15) & ~15;          // one bit per spin
    ip.per_warp_bytes = (ip.arena_floats + ip.spt_floats) * 4 + ip.spins_bytes;
    ip.mg2p = fastdiv_magic(2 * p);
    ip.mg2p1 = fastdiv_magic(2 * p + 1);
    for (int j = 0; j < m.D; ++j) ip.mgW[j] = fastdiv_magic(2 * (j + 2) * p + 1);
    // site tables of the tiled layers (ip_site_table): p rows of ns site slots each
    int entries = 0;
    for (int l = 0; l < QMC_MAX_LAYERS; ++l) ip.tab_off[l] = -1;
    for (int l = 1; l < m.D; ++l) {
        const int side = 1 + 2 * (l + 1) * p;
        ip.tab_off[l] = entries;
        const IpTile t = ip_tile(kIpAcc, m.layer[l].cout, side * side);
        entries += (t.p + 1) * t.ns;
    }
    ip.tab_entries = (entries + 7) & ~7;
    ip.ok = 1;
    return ip;
}

// Deal the side x side window of layer l to (round j, site slot) so that the eight slots the hardware serves together
// (a quarter warp of the one-part tile; the 16 lanes of a half-warp of the split-channel tile) read eight different
// 16-byte bank groups of the arena (pitch T float4): walk the sites in row-major order and give each group of eight
// the first unassigned sites whose (y * T + x) mod 8 it does not hold yet.  Row-major preference keeps the staging
// stores of a group nearly contiguous.  Slots >= G = ceil(npos / P) stay idle.  conflict_free = false
// (QMC_FLAG_IP_ROWMAJOR_SITES): plain row-major deal, for the before / after measurement.
static void ip_site_table(int side, int T, int P, int NS, bool conflict_free, site_t* tab) {
    const int npos = side * side, G = (npos + P - 1) / P;
    std::vector<char> taken(npos, 0);
    for (int i = 0; i < (P + 1) * NS; ++i) tab[i] = kNoSite;
    int left = npos;
    for (int j = 0; j < P; ++j)
        for (int s0 = 0; s0 < NS; s0 += 8) {
            unsigned used = 0;
            for (int slot = s0; slot < s0 + 8 && slot < NS && slot < G && left > 0; ++slot) {
                int pick = -1, fallback = -1;
                for (int pos = 0; pos < npos; ++pos) {
                    if (taken[pos]) continue;
                    if (fallback < 0) fallback = pos;
                    const int y = pos / side, x = pos - y * side;
                    if (!conflict_free || !((used >> ((y * T + x) & 7)) & 1u)) { pick = pos; break; }
                }
                if (pick < 0) pick = fallback;          // no conflict-free site is left for this group
                const int y = pick / side, x = pick - y * side;
                used |= 1u << ((y * T + x) & 7);
                taken[pick] = 1;
                --left;
                tab[j * NS + slot] = make_site(y, x, T);
            }
        }
    finish_site_table(tab, P, NS);
}

// device image of the site tables of a handle (built once, qmc_create)
cudaError_t ip_upload_tables(qmc_handle* h) {
    h->d_ip_tab = nullptr;
    const IpPlan ip = ip_plan(h);
    if (!ip.ok || ip.tab_entries == 0) return cudaSuccess;
    std::vector<site_t> tab(ip.tab_entries, kNoSite);
    for (int l = 1; l < h->m.D; ++l) {
        const int side = 1 + 2 * (l + 1) * h->m.p;
        const IpTile t = ip_tile(kIpAcc, h->m.layer[l].cout, side * side);
        ip_site_table(side, ip.T, t.p, t.ns, h->ip_cf, tab.data() + ip.tab_off[l]);
    }
    cudaError_t e = cudaMalloc(&h->d_ip_tab, tab.size() * sizeof(site_t));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(h->d_ip_tab, tab.data(), tab.size() * sizeof(site_t), cudaMemcpyHostToDevice);
}

IpLaunch ip_launch_plan(const qmc_handle* h, int S) {
    IpLaunch L{};
    L.ok = false;
    if (!h->allow_ip) return L;
    L.ip = ip_plan(h);
    if (!L.ip.ok) return L;
    const size_t cta_bytes = (size_t)h->m.smem_param_floats * 4 + ip_cta_bytes(L.ip);
    if (h->max_smem < cta_bytes + (size_t)L.ip.per_warp_bytes) return L;
    int w = (int)((h->max_smem - cta_bytes) / (size_t)L.ip.per_warp_bytes);
    if (w > kIpMaxWarps) w = kIpMaxWarps;
    if (h->max_warps_override > 0 && w > h->max_warps_override) w = h->max_warps_override;
    // the classic kernel keeps bigger register tiles; in-place only pays when it adds warps
    EvalPlan pl = eval_plan(h->m, 1, 1, false);
    const WarpGrid gc = pick_warp_grid(h, pl.per_warp_bytes, 0, S);
    if (!h->force_ip && gc.ok && (gc.warps > 8 || w <= gc.warps)) return L;
    // fewer chains than slots: one launch, as few warps per CTA as cover S (more shared memory per warp is no use)
    const long long ctas = h->num_sms;
    if ((long long)S < ctas * w) w = (int)((S + ctas - 1) / ctas);
    L.warps = w;
    const long long need = ((long long)S + w - 1) / w;
    L.grid = (int)(need < ctas ? need : ctas);
    L.smem = cta_bytes + (size_t)L.ip.per_warp_bytes * w;
    L.ok = true;
    return L;
}

// Time slicing of a sweep of n_steps for S chains on L's warp slots (host only): chunks of >= 64 steps, at most
// ip_chunks (64) chunks per chain; one chunk when S fits the slots.  Returns the slice (task0 = 0) and the launch count.
IpSlice ip_slice_plan(const qmc_handle* h, const IpLaunch& L, int S, long long n_steps, long long* launches) {
    const long long slots = (long long)L.grid * L.warps;
    long long chunks = 1;
    if ((long long)S > slots) {
        const long long cmax = h->ip_chunks > 0 ? h->ip_chunks : 64;     // at most this many chunks per chain
        chunks = n_steps / 64;
        if (chunks > cmax) chunks = cmax;
        if (chunks < 1) chunks = 1;
    }
    IpSlice sl;
    sl.group_warps = h->ip_group > 0 ? h->ip_group : 4;
    sl.chunk_len = (n_steps + chunks - 1) / chunks;
    sl.stagger = sl.chunk_len >= 64 ? h->ip_stagger * 1024 : 0;     // not worth ~40 us on a launch of a few steps
    chunks = (n_steps + sl.chunk_len - 1) / sl.chunk_len;
    sl.n_tasks = chunks * S;
    sl.task0 = 0;
    if (launches) *launches = (sl.n_tasks + slots - 1) / slots;
    return sl;
}

void ip_slice_counts(const qmc_handle* h, const IpLaunch& L, int S, long long n_steps, long long* launches, long long* chunk_len) {
    const IpSlice sl = ip_slice_plan(h, L, S, n_steps, launches);
    if (chunk_len) *chunk_len = sl.chunk_len;
}

cudaError_t launch_sweep_ip(const qmc_handle* h, const SweepArgs& a, const IpLaunch& L, cudaStream_t st) {
    const int sy = h->ip_sync;
    // two instances are built: free-running (QMC_IP_SYNC=0, diagnosis) and phase groups (default).  Per-proposal
    // group barriers (1) and CTA-wide per-layer barriers (2) were measured (profiles/r01_summary.md) and dropped.
    cudaError_t e = cudaSuccess;
    const long long slots = (long long)L.grid * L.warps;
    IpSlice sl = ip_slice_plan(h, L, a.S, a.n_steps, nullptr);
    for (sl.task0 = 0; sl.task0 < sl.n_tasks; sl.task0 += slots) {
        const long long left = sl.n_tasks - sl.task0;
        const long long ctas = ((left < slots ? left : slots) + L.warps - 1) / L.warps;
        ++g_launches;
        e = (sy == 0 ? ip_launch_sweep_free : ip_launch_sweep_groups)((int)ctas, L.warps * 32, L.smem, st, h->m, h->d_params_padded,
                                                                      a, L.ip, sl, h->d_ip_tab);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
#endif // QMC_IP_PART == 0

// ------------------------------------------------------------------------------------------------
// k_energy_ip: TFIM local energies (ising_energy, mcmc_tf.py:59-90) with the same in-place evaluator.
// Warp task = (sample, chunk of sites): every single-flip configuration of the chunk is evaluated
// against the sample's read-only cache (K1), exp(log_pop) is accumulated in registers in site order
// (the classic k_energy's order), k_energy_finish adds the diagonal term.  Nothing is staged or
// committed.  Phase-group barriers as in k_sweep_ip: every warp runs the same number of tasks and
// sites, surplus ones shadow a valid evaluation without writing.
// ------------------------------------------------------------------------------------------------
#if QMC_IP_PART == 2
__global__ void __launch_bounds__(kIpMaxWarps * 32, 1)
k_energy_ip(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins, int N,
            const float* __restrict__ cache_all, float2* __restrict__ partial, int nchunks, IpPlan ip,
            int group_warps, const site_t* __restrict__ tab_g) {
    extern __shared__ float4 smem4[];
    float* smem_f = reinterpret_cast<float*>(smem4);
    load_params_to_smem(m, params, smem_f);
    const float* sp = smem_f;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned* mg; int* tabo; site_t* tab_s;
    char* wmem = ip_cta_setup(smem_f + m.smem_param_floats, ip, tab_g, mg, tabo, tab_s) + (size_t)warp * ip.per_warp_bytes;
    float* arena = reinterpret_cast<float*>(wmem);
    float* spt = arena + ip.arena_floats;
    unsigned* spins_s = reinterpret_cast<unsigned*>(spt + ip.spt_floats);     // one bit per spin (ip_spin)
    const int n = m.n;
    const int g = warp / group_warps, gid = 1 + g, gthreads = 32 * min(group_warps, nwarps - g * group_warps);
    const int cs = (n + nchunks - 1) / nchunks;
    const long long ntasks = (long long)N * nchunks;
    const long long slot = (long long)blockIdx.x * nwarps + warp, nslots = (long long)gridDim.x * nwarps;
    int loaded = -1;
    for (long long t0 = 0; t0 < ntasks; t0 += nslots) {         // the same trip count for every warp
        const long long task = t0 + slot;
        const bool has_task = task < ntasks;
        const long long tt = has_task ? task : ntasks - 1;
        const int s = (int)(tt / nchunks), chunk = (int)(tt - (long long)s * nchunks);
        if (s != loaded) {
            __syncwarp();
            ip_pack_spins(spins_s, spins + (size_t)s * n, n, lane);
            __syncwarp();
            loaded = s;
        }
        const float* cache = cache_all + (size_t)s * m.cache_floats;
        float are = 0.f, aim = 0.f;
        const int i0 = chunk * cs, i1 = min(n, i0 + cs);
        for (int ii = 0; ii < cs; ++ii) {
            const bool active = i0 + ii < i1;
            const int i = active ? i0 + ii : i1 - 1;
            float dre, dim, sn, cn;
            IpProf prof;
            ip_barrier<3>(gid, gthreads);
            warp_eval_flip_ip<kIpAcc, 3, false>(m, ip, sp, mg, arena, spt, spins_s, cache, nullptr, i, lane, gid,
                                                gthreads, dre, &dim, tabo, tab_s, prof);
            if (!active) continue;
            const float amp = expf(dre);
            sincosf(dim, &sn, &cn);
            are += amp * cn;               // exp(log_pop), mcmc_tf.py:88
            aim += amp * sn;
        }
        if (has_task && lane == 0) partial[(size_t)s * nchunks + chunk] = make_float2(are, aim);
    }
}

cudaError_t ip_launch_energy(int grid, int threads, size_t smem, cudaStream_t st, const DevModel& m, const float* params,
                             const int8_t* spins, int N, const float* cache, float2* partial, int nchunks, const IpPlan& ip,
                             int group_warps, const site_t* tab) {
    cudaError_t e = cudaFuncSetAttribute(k_energy_ip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_energy_ip<<<grid, threads, smem, st>>>(m, params, spins, N, cache, partial, nchunks, ip, group_warps, tab);
    return cudaGetLastError();
}
#endif // QMC_IP_PART == 2

#if QMC_IP_PART == 0
// TFIM local energies through the in-place evaluator, if the model is inside its coverage
bool energy_ip_supported(const qmc_handle* h) {
    if (!h->allow_ip) return false;
    const IpPlan ip = ip_plan(h);
    if (!ip.ok) return false;
    const size_t cta_bytes = (size_t)h->m.smem_param_floats * 4 + ip_cta_bytes(ip);
    return h->max_smem >= cta_bytes + (size_t)ip.per_warp_bytes;
}

cudaError_t launch_energy_ip(const qmc_handle* h, const int8_t* spins, int N, const float* cache, float2* partial,
                             int nchunks, cudaStream_t st) {
    const IpPlan ip = ip_plan(h);
    const size_t cta_bytes = (size_t)h->m.smem_param_floats * 4 + ip_cta_bytes(ip);
    int w = (int)((h->max_smem - cta_bytes) / (size_t)ip.per_warp_bytes);
    if (w > kIpMaxWarps) w = kIpMaxWarps;
    if (h->max_warps_override > 0 && w > h->max_warps_override) w = h->max_warps_override;
    const long long ntasks = (long long)N * nchunks;
    if (ntasks < (long long)h->num_sms * w) w = (int)((ntasks + h->num_sms - 1) / h->num_sms);
    const long long need = (ntasks + w - 1) / w;
    const int grid = (int)(need < h->num_sms ? need : h->num_sms);
    const size_t smem = cta_bytes + (size_t)ip.per_warp_bytes * w;
    ++g_launches;
    return ip_launch_energy(grid, w * 32, smem, st, h->m, h->d_params_padded, spins, N, cache, partial, nchunks, ip,
                            h->ip_group > 0 ? h->ip_group : 4, h->d_ip_tab);
}
#endif // QMC_IP_PART == 0

} // namespace qmc

#if QMC_IP_PART == 0
// phase cycles of k_sweep_ip since the last call (QMC_IP_PROFILE builds; zeros otherwise): out[0..7] cycles summed
// over warps, out[8] proposals, out[9 + w] task duration of warp w of a CTA summed over CTAs and launches
extern "C" int qmc_diag_ip_profile(unsigned long long* out /*host, 21 entries*/) {
    if (!out) return QMC_ERR_BAD_ARGUMENT;
    for (int i = 0; i < qmc::kIpProfPhases + 13; ++i) out[i] = 0;
#if QMC_IP_PROFILE
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, qmc::g_ip_prof, sizeof(qmc::g_ip_prof)) != cudaSuccess) return QMC_ERR_CUDA;
    unsigned long long zero[qmc::kIpProfPhases + 13] = {};
    if (cudaMemcpyToSymbol(qmc::g_ip_prof, zero, sizeof(zero)) != cudaSuccess) return QMC_ERR_CUDA;
    unsigned long long conc[4];
    if (cudaMemcpyFromSymbol(conc, qmc::g_ip_conc, sizeof(conc)) != cudaSuccess) return QMC_ERR_CUDA;
    if (cudaMemcpyToSymbol(qmc::g_ip_conc, zero, sizeof(conc)) != cudaSuccess) return QMC_ERR_CUDA;
    const double ct = (double)(conc[0] + conc[1] + conc[2] + conc[3]);
    if (ct > 0)
        fprintf(stderr, "k_sweep_ip conv concurrency (tap iterations that saw 1 / 2 / 3 / 4+ warps of their scheduler in a conv "
                "loop): %.3f %.3f %.3f %.3f\n", conc[0] / ct, conc[1] / ct, conc[2] / ct, conc[3] / ct);
#endif
    return QMC_OK;
}
#endif // QMC_IP_PART == 0

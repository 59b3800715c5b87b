// qmc_plane.cu - K1 with the whole lattice of a sample resident in shared memory (k_forward_plane).
// Replaces models.py:95-131 (DCRBM.factors) + helpers.py:73-91 (pad) for lattices whose activation planes fit.
//
// k_forward (qmc_forward.cu) stages 8 x 8 blocks + halo through per-warp tiles: a 20 x 20 lattice is 9 blocks of
// different sizes for 8 warps, every layer makes a round trip through the cache in L2, the staging loops divide
// per element, and the SM holds 8 warps: 0.17 of the FP32 roofline (profiles/r01_summary.md).  Here a CTA keeps
// two WRAP-PADDED planes of the sample in shared memory ((Ly + 2p) x (Lx + 2p) sites, channel-group planar float4,
// ping-pong): a layer reads its input with plain tile arithmetic (the halo is materialised by the writer of a site:
// up to three mirror stores), writes tanh outputs to the other plane and - coalesced, as the by-product the sweep,
// energy and backward kernels consume - to the cache.  The sites of the lattice are dealt to the warps in equal
// contiguous chunks and each warp runs the split-channel register tile of the in-place evaluator
// (conv_region_split: lane = site slot x channel part) over its chunk, the chunk's sites dealt to the slots
// conflict-free (plane_site_table).  Two CTAs per SM.  Same fma chain per output as every other path
// (bias, taps ascending, input channels ascending; tanhf; site_factor), so caches and factors are bit-identical
// to k_forward's (tests/test_gpu_parity.py).
#include <vector>
#include "qmc_host.h"
#include "qmc_ip.cuh"

namespace qmc {

constexpr int kPlaneMaxWarps = 8;        // x 2 CTAs per SM: 128 registers per thread

// store a site's value at its interior position of a wrap-padded plane and at its periodic images inside the halo
template <typename F>
__device__ __forceinline__ void plane_halo_store(float4* plane, int PW, int Ly, int Lx, int p, int y, int x, float4 a,
                                                 F idx_of) {
    // interior position (y + p, x + p) and its periodic images inside the halo
    const int ys[2] = {y + p, y < p ? y + p + Ly : (y >= Ly - p ? y + p - Ly : -1)};
    const int xs[2] = {x + p, x < p ? x + p + Lx : (x >= Lx - p ? x + p - Lx : -1)};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (ys[i] < 0) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (xs[j] >= 0) {
                QMC_ASSERT(ys[i] < Ly + 2 * p && xs[j] < PW, "halo store inside the padded plane");
                plane[idx_of(ys[i] * PW + xs[j])] = a;
            }
    }
}

struct PlanePlan {
    int ok;
    int warps, P;               // warps per CTA, sites per slot (split tile with two channel parts: 16 slots)
    int chunk;                  // sites per warp (the last warp may have fewer)
    int PW, PA;                 // padded pitch and padded area (float4 per channel-group plane)
    int plane_floats;           // one padded plane, all channel groups
    int tab_entries;            // uint16 entries of the site tables: [warp][j][slot]
    size_t smem;
};

__global__ void __launch_bounds__(kPlaneMaxWarps * 32, 2)
k_forward_plane(DevModel m, const float* __restrict__ params, const int8_t* __restrict__ spins, int N,
                float* __restrict__ cache_all, float2* __restrict__ factors, float2* __restrict__ logpsi,
                PlanePlan pp, const site_t* __restrict__ tab_g, ImageStrides is) {
    extern __shared__ float4 smem4[];
    float* sp = reinterpret_cast<float*>(smem4);
    params += (size_t)blockIdx.y * is.params;                  // symmetry images: own parameter block, cache, outputs
    cache_all += (size_t)blockIdx.y * is.cache;
    if (factors) factors += (size_t)blockIdx.y * N * m.n;
    if (logpsi) logpsi += (size_t)blockIdx.y * N;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, tid = threadIdx.x, nthr = blockDim.x;
    float* B0 = sp + m.smem_param_floats;
    float* B1 = B0 + pp.plane_floats;
    float* S = B1 + pp.plane_floats;                            // padded spin plane (floats), PA entries
    float* red = S + round4(pp.PA);                             // 2 x 32 partial sums (log psi)
    site_t* tab_s = reinterpret_cast<site_t*>(red + 64);
    load_params_to_smem(m, params, sp);
    for (int i = tid; i < pp.tab_entries; i += nthr) tab_s[i] = tab_g[i];
    __syncthreads();
    const site_t* tab = tab_s + warp * (pp.P + 1) * 16;

    const int p = m.p, Ly = m.Ly, Lx = m.Lx, n = m.n, D = m.D, PW = pp.PW, PA = pp.PA;
    const FastDiv dPW(PW), dLx(Lx), dn(n);
    for (int s = blockIdx.x; s < N; s += gridDim.x) {
        float* cache = cache_all + (size_t)s * m.cache_floats;
        const int8_t* sx = spins + (size_t)s * n;
        // padded spin plane
        for (int i = tid; i < PA; i += nthr) {
            const int py = dPW.div(i), px = i - py * PW;
            S[i] = (float)sx[wrap1(py - p, Ly) * Lx + wrap1(px - p, Lx)];
        }
        __syncthreads();
        // ---- layer 0 (C_in = 1): one thread per (site, 4 output channels) ------------------------------------
        {
            const LayerInfo& L = m.layer[0];
            const int ncog = L.coutp >> 2;
            float4* out4 = reinterpret_cast<float4*>(B0);
            float4* plane4 = reinterpret_cast<float4*>(cache + L.act_off);
            for (int t = tid; t < n * ncog; t += nthr) {
                const int cog = dn.div(t), site = t - cog * n;      // consecutive threads: consecutive sites
                const int y = dLx.div(site), x = site - y * Lx;
                float4 acc = *reinterpret_cast<const float4*>(sp + L.sb_off + cog * 4);
                const float* wb = sp + L.sw_off + cog * 4;
                for (int dy = 0; dy < m.k; ++dy)
                    for (int dx = 0; dx < m.k; ++dx) {
                        const float in = S[(y + dy) * PW + x + dx];
                        const float4 w = *reinterpret_cast<const float4*>(wb + (dy * m.k + dx) * L.coutp);
                        acc.x = fmaf(in, w.x, acc.x); acc.y = fmaf(in, w.y, acc.y);
                        acc.z = fmaf(in, w.z, acc.z); acc.w = fmaf(in, w.w, acc.w);
                    }
                acc.x = tanhf(acc.x); acc.y = tanhf(acc.y); acc.z = tanhf(acc.z); acc.w = tanhf(acc.w);
                plane_halo_store(out4, PW, Ly, Lx, p, y, x, acc, [&](int q) { return cog * PA + q; });
                plane4[cog * n + site] = acc;
            }
        }
        __syncthreads();
        // ---- layers 1 .. D-1: split-channel register tile over this warp's chunk of sites -------------------
        for (int l = 1; l < D; ++l) {
            const LayerInfo& L = m.layer[l];
            const bool last = (l == D - 1);
            const float* tin = (l & 1) ? B0 : B1;
            float4* out4 = reinterpret_cast<float4*>((l & 1) ? B1 : B0);
            float4* plane4 = last ? nullptr : reinterpret_cast<float4*>(cache + L.act_off);
            auto hidden = [&](int pos, int y, int x, int cog, float4 a) {
                a = ip_tanh4(a);
                plane_halo_store(out4, PW, Ly, Lx, p, y, x, a, [&](int q) { return cog * PA + q; });
                plane4[cog * n + pos] = a;
            };
            auto theta = [&](int pos, int, int, int cog, float4 a) { out4[cog * n + pos] = a; };
            NoMid mid;
#define QMC_PLANE(CI, CO, FN)                                                                                         \
    switch (pp.P) {                                                                                                   \
        case 4: conv_region_split<3, CI, CO, 2, 4>(L.sw_off, L.sb_off, sp, tin, PW, PA, Lx, lane, FN, mid, tab, pp.plane_floats >> 2); break; \
        default: conv_region_split<3, CI, CO, 2, 6>(L.sw_off, L.sb_off, sp, tin, PW, PA, Lx, lane, FN, mid, tab, pp.plane_floats >> 2); break; \
    }
            if (!last) {
                if (L.cin == 16) { QMC_PLANE(16, 16, hidden) } else { QMC_PLANE(8, 8, hidden) }
            } else {
                if (L.cin == 16 && L.cout == 16) { QMC_PLANE(16, 16, theta) }
                else if (L.cin == 16) { QMC_PLANE(16, 8, theta) }
                else { QMC_PLANE(8, 8, theta) }
            }
#undef QMC_PLANE
            __syncthreads();
        }
        // ---- head: per-site complex factor, log psi ---------------------------------------------------------
        {
            const float* th = ((D - 1) & 1) ? B1 : B0;
            float re_sum = 0.f, im_sum = 0.f;
            for (int site = tid; site < n; site += nthr) {
                float re, im;
                site_factor<true>(m, sp, th, n, site, (float)sx[site], re, im);
                cache[m.fre_off + site] = re;
                cache[m.fim_off + site] = im;
                if (factors) factors[(size_t)s * n + site] = make_float2(re, im);
                re_sum += re;
                im_sum += im;
            }
            if (logpsi) {                                   // deterministic: lanes, then warps in index order
                re_sum = warp_sum(re_sum);
                im_sum = warp_sum(im_sum);
                if (lane == 0) { red[warp] = re_sum; red[32 + warp] = im_sum; }
                __syncthreads();
                if (tid == 0) {
                    float r = 0.f, i = 0.f;
                    for (int w = 0; w < (nthr >> 5); ++w) { r += red[w]; i += red[32 + w]; }
                    logpsi[s] = make_float2(r, i);
                }
            }
        }
        __syncthreads();
    }
}

// Deal the chunk [s0, s1) of row-major lattice sites to (round j, slot): the eight slots of a half-warp get sites in
// eight different 16-byte bank groups of the padded plane ((y * PW + x) mod 8; the tap offset shifts all of them
// alike), preferring row-major order so that the cache stores of a half-warp stay nearly contiguous.
static void plane_site_table(int s0, int s1, int Lx, int PW, int P, site_t* tab) {
    const int cnt = s1 - s0, G = (cnt + P - 1) / P;
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

static PlanePlan plane_plan(const qmc_handle* h) {
    PlanePlan pp{};
    const DevModel& m = h->m;
    if (!h->allow_tiled || h->forward_blocked || m.kind != QMC_MODEL_DCRBM || m.D < 2 || m.k != 3) return pp;
    if (m.Ly > 255 || m.Lx > 255) return pp;
    int cmax = 0;
    for (int l = 0; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        if (l >= 1) {
            const bool hidden_ok = (L.cin == 16 && L.cout == 16) || (L.cin == 8 && L.cout == 8);
            const bool last_ok = hidden_ok || (L.cin == 16 && L.cout == 8);
            if (l < m.D - 1 ? !hidden_ok : !last_ok) return pp;
        }
        if (L.coutp > cmax) cmax = L.coutp;
    }
    pp.PW = m.Lx + 2 * m.p;
    pp.PA = (m.Ly + 2 * m.p) * pp.PW;
    pp.plane_floats = round4(pp.PA * cmax);
    // warps x tile height: the best use of the 16 x P site slots of a warp, ties to the taller tile
    double best = -1;
    for (int P = 6; P >= 4; P -= 2)                                 // (128 registers: no 8-site tile)
        for (int w = 1; w <= kPlaneMaxWarps; ++w) {
            const int chunk = (m.n + w - 1) / w;
            if (chunk > P * 16) continue;
            const double util = (double)m.n / (double)(w * P * 16);
            if (util > best + 1e-9) { best = util; pp.warps = w; pp.P = P; pp.chunk = chunk; }
            break;                                                  // the smallest w that fits this P
        }
    if (best < 0 || m.n * 4 > 65535) return pp;                     // more than 8 x 96 sites: k_forward
    pp.tab_entries = (pp.warps * (pp.P + 1) * 16 + 7) & ~7;
    pp.smem = ((size_t)m.smem_param_floats + 2 * (size_t)pp.plane_floats + round4(pp.PA) + 64) * 4 + (size_t)pp.tab_entries * sizeof(site_t);
    if (pp.smem > h->max_smem) return pp;
    pp.ok = 1;
    return pp;
}

bool forward_plane_supported(const qmc_handle* h) { return plane_plan(h).ok != 0; }

// device image of the plane kernel's site tables (built once, qmc_create)
cudaError_t plane_upload_tables(qmc_handle* h) {
    h->d_plane_tab = nullptr;
    const PlanePlan pp = plane_plan(h);
    if (!pp.ok) return cudaSuccess;
    std::vector<site_t> tab(pp.tab_entries, kNoSite);
    for (int w = 0; w < pp.warps; ++w) {
        const int s0 = w * pp.chunk, s1 = s0 + pp.chunk < h->m.n ? s0 + pp.chunk : h->m.n;
        plane_site_table(s0 < h->m.n ? s0 : h->m.n, s1, h->m.Lx, pp.PW, pp.P, tab.data() + (size_t)w * (pp.P + 1) * 16);
    }
    cudaError_t e = cudaMalloc(&h->d_plane_tab, tab.size() * sizeof(site_t));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(h->d_plane_tab, tab.data(), tab.size() * sizeof(site_t), cudaMemcpyHostToDevice);
}

cudaError_t launch_forward_plane(const qmc_handle* h, int nimg, const float* padded_blocks, const int8_t* spins, int N,
                                 float* cache, float* factors, float* logpsi, cudaStream_t st) {
    const PlanePlan pp = plane_plan(h);
    cudaError_t e = cudaFuncSetAttribute(k_forward_plane, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)(((size_t)227 * 1024) / (pp.smem + 1024));
    per_sm = per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm;
    const int per_img = (h->num_sms * per_sm + nimg - 1) / nimg;
    const int grid = N < per_img ? N : per_img;
    ++g_launches;
    const ImageStrides is{(size_t)h->m.smem_param_floats, (size_t)N * h->m.cache_floats};
    k_forward_plane<<<dim3(grid, nimg), pp.warps * 32, pp.smem, st>>>(h->m, padded_blocks, spins, N, cache,
                                                                     reinterpret_cast<float2*>(factors),
                                                                     reinterpret_cast<float2*>(logpsi), pp, h->d_plane_tab, is);
    return cudaGetLastError();
}

} // namespace qmc

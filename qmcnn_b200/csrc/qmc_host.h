// qmc_host.h - host-side handle, launch-geometry helpers and launcher
// declarations shared by the translation units of libqmcnn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "qmc_device.cuh"

struct qmc_handle {
    int device = 0;
    qmc_model_desc desc{};
    qmc::DevModel m{};
    float* d_params = nullptr;         // caller's flat order
    float* d_params_padded = nullptr;  // padded block layout (shared / constant memory image)
    float* d_sym_padded = nullptr;     // [nsym] padded blocks of the symmetry images (qmc_set_image_params)
    int nsym = 0;
    std::string err;
    int num_sms = 0;
    size_t max_smem = 0;   // opt-in dynamic shared memory per CTA
    bool allow_tiled = true;  // QMC_FORCE_GENERIC=1 disables the specialised conv instances
    int max_warps_override = 0;  // QMC_MAX_WARPS: tuning knob, caps the warps per CTA of the persistent kernels
    bool allow_lean = false;     // QMC_LEAN=1: lean persistent sweep kernel (more warps, smaller tiles); off by default
    bool allow_ip = true;        // QMC_SWEEP_PATH=pingpong disables the in-place persistent sweep kernel (k_sweep_ip)
    int ip_group = 4;            // QMC_IP_GROUP: warps per phase group of k_sweep_ip<3>
    int ip_sync = 3;             // QMC_IP_SYNC: 0 = k_sweep_ip's warps run free, otherwise (default) a named barrier per
                                 // layer within each phase group of ip_group warps
    bool ip_cf = true;           // conflict-free site tables for the in-place evaluator (QMC_IP_CF=0: the r01 order)
    unsigned short* d_ip_tab = nullptr;   // device image of the tables (ip_upload_tables), or nullptr
    bool force_ip = false;       // QMC_SWEEP_PATH=inplace: use k_sweep_ip whenever the model is inside its coverage
    bool allow_batched = true;  // QMC_FORCE_PERSISTENT=1 disables the layer-synchronous batched path (energy + sweep)
    bool batched_sweep = false; // QMC_SWEEP_PATH=batched: use the batched path for the sweep too (default: persistent
                                // kernel, which is faster at a few thousand chains per GPU - DESIGN.md)
    cudaStream_t side_stream[2] = {nullptr, nullptr};   // batched sweep: capturable streams, event-ordered with the caller's
    cudaEvent_t ev_in = nullptr, ev_mid = nullptr, ev_out[2] = {nullptr, nullptr};
};

namespace qmc {

// number of kernels this library has launched (bench.py's gpu_launches evidence)
extern unsigned long long g_launches;

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

// shared-memory plan of warp_eval_flip for boxes up to h0max x w0max
struct EvalPlan {
    int buf_floats[2];   // ping-pong tile buffers
    int newf_floats;     // new per-site factors over the last region (x2 when Im is needed)
    int nfstride;
    int spins_bytes;     // per-warp lattice copy
    int staging_floats;  // new hidden activations of all regions (global scratch per warp)
    size_t per_warp_bytes;
};

inline EvalPlan eval_plan(const DevModel& m, int h0max, int w0max, bool need_im) {
    EvalPlan pl{};
    const int p = m.p;
    int rh = h0max + 2 * p, rw = w0max + 2 * p;
    if (rh > m.Ly) rh = m.Ly;
    if (rw > m.Lx) rw = m.Lx;
    int bf[2] = {(rh + 2 * p) * (rw + 2 * p), 0};
    int stg = 0;
    for (int l = 0; l < m.D; ++l) {
        const LayerInfo& L = m.layer[l];
        int& dst = bf[(l + 1) & 1];
        if (l == m.D - 1) {
            dst = dst > rh * rw * L.coutp ? dst : rh * rw * L.coutp;
        } else {
            const int t = (rh + 4 * p) * (rw + 4 * p) * L.coutp;
            dst = dst > t ? dst : t;
            stg += L.coutp * rh * rw;
            rh += 2 * p; rw += 2 * p;
        }
    }
    pl.buf_floats[0] = round4(bf[0]);
    pl.buf_floats[1] = round4(bf[1]);
    pl.nfstride = round4(rh * rw);
    pl.newf_floats = pl.nfstride * (need_im ? 2 : 1);
    pl.spins_bytes = (m.n + 15) & ~15;
    pl.staging_floats = round4(stg);
    pl.per_warp_bytes = (size_t)(pl.buf_floats[0] + pl.buf_floats[1] + pl.newf_floats) * 4 + pl.spins_bytes;
    return pl;
}

// does a box of h0 x w0 flipped sites fit the incremental evaluator?
inline bool box_supported(const DevModel& m, int h0, int w0) {
    if (m.D == 1) return true;   // regions clamp to the lattice, spins are re-read by coordinate
    return h0 + 2 * m.D * m.p <= m.Ly && w0 + 2 * m.D * m.p <= m.Lx;
}

// shared-memory plan of warp_eval_flip_lean for a per-warp arena of at most budget_floats
inline LeanPlan lean_plan(const DevModel& m, int budget_floats) {
    LeanPlan best{};
    best.ok = 0;
    if (m.D < 2 || m.r > m.Ly || m.r > m.Lx) return best;
    const int p = m.p;
    for (int fg = m.D - 1; fg >= 1; --fg) {            // prefer as few gathering layers as possible
        LeanPlan lp{};
        lp.first_gather = fg;
        int bufa = 0, bufb = 0;
        for (int l = 0; l < fg; ++l) {                 // chained tiles alternate between the two buffers
            const int tside = 1 + 2 * (l + 1) * p + 2 * p;
            const int t = tside * tside * m.layer[l].cinp;
            int& dst = (l & 1) ? bufb : bufa;
            dst = dst > t ? dst : t;
        }
        lp.off_b = round4(bufa);
        int need = round4(bufa) + round4(bufb);
        bool ok = need <= budget_floats;
        for (int l = fg; l < m.D && ok; ++l) {
            const int side = 1 + 2 * (l + 1) * p, tside = side + 2 * p;
            const bool last = l == m.D - 1;
            int nb = 0;
            for (int b = 1; b <= 4; ++b) {
                const int rows = (side + b - 1) / b;
                int t = round4((rows + 2 * p) * tside * m.layer[l].cinp);
                if (last) t += round4(rows * side * m.layer[l].coutp);
                if (t <= budget_floats) { nb = b; need = need > t ? need : t; break; }
            }
            if (!nb) ok = false;
            lp.bands[l] = nb;
        }
        if (ok) { lp.arena_floats = round4(need); lp.ok = 1; return lp; }
    }
    return best;
}

struct WarpGrid { int grid, warps; size_t smem; bool ok; };

// one CTA per SM, W warps, W chosen to minimise the idle tail over `units` warp tasks
inline WarpGrid pick_warp_grid(const qmc_handle* h, size_t per_warp_bytes, size_t cta_bytes,
                               long long units, int max_warps = 16) {
    WarpGrid g{0, 0, 0, false};
    const size_t param_bytes = (size_t)h->m.smem_param_floats * 4 + cta_bytes;
    if (param_bytes + per_warp_bytes > h->max_smem) return g;
    int wmax = (int)((h->max_smem - param_bytes) / per_warp_bytes);
    if (wmax > max_warps) wmax = max_warps;
    if (h->max_warps_override > 0 && wmax > h->max_warps_override) wmax = h->max_warps_override;
    int best = wmax;
    double best_eff = -1;
    const int wmin = wmax > 2 ? (wmax + 1) / 2 : 1;
    for (int w = wmax; w >= wmin; --w) {
        const long long slots = (long long)h->num_sms * w;
        const long long waves = (units + slots - 1) / slots;
        const double eff = (double)units / (double)(waves * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = w; }
    }
    g.warps = best;
    long long ctas = (units + best - 1) / best;
    g.grid = (int)(ctas < h->num_sms ? ctas : h->num_sms);
    if (g.grid < 1) g.grid = 1;
    g.smem = param_bytes + per_warp_bytes * best;
    g.ok = true;
    return g;
}

struct SweepArgs {
    int8_t* spins; float* cache; float* staging;
    int S, num_flips;
    long long step0, n_steps;
    const int32_t* flip_pos; const float* uniforms;
    unsigned long long seed; long long chain_id0;
    long long therm_its, its_per_sample;
    int8_t* samples; long long n_sample_slots; uint8_t* accept_trace; float* logratio_trace;
    unsigned long long* n_accept;
};

// launchers (each in its own .cu); return cudaError_t of the launch
cudaError_t repack_params(const qmc_handle* h, cudaStream_t st);
cudaError_t launch_forward(const qmc_handle* h, const int8_t* spins, int N, float* cache,
                           float* factors, float* logpsi, cudaStream_t st, std::string& err);
cudaError_t launch_sweep(const qmc_handle* h, const SweepArgs& a, cudaStream_t st, std::string& err);
struct LeanLaunch { LeanPlan lp; int warps, grid; size_t smem; int newf_floats, spins_bytes, staging_floats; bool ok; };
LeanLaunch lean_launch_plan(const qmc_handle* h, int S);
cudaError_t launch_sweep_lean(const qmc_handle* h, const SweepArgs& a, const LeanLaunch& ll, cudaStream_t st);
struct IpLaunch { IpPlan ip; int warps, grid; size_t smem; bool ok; };
IpPlan ip_plan(const qmc_handle* h);
cudaError_t ip_upload_tables(qmc_handle* h);
IpLaunch ip_launch_plan(const qmc_handle* h, int S);
cudaError_t launch_sweep_ip(const qmc_handle* h, const SweepArgs& a, const IpLaunch& L, cudaStream_t st);
bool energy_ip_supported(const qmc_handle* h);
cudaError_t launch_energy_ip(const qmc_handle* h, const int8_t* spins, int N, const float* cache, float2* partial,
                             int nchunks, cudaStream_t st);
cudaError_t launch_sweep_sym(const qmc_handle* h, const SweepArgs& a, int nsym, double* drel, cudaStream_t st,
                             std::string& err);
int sweep_sym_slots(const qmc_handle* h, int S, int num_flips, int nsym, EvalPlan* plan, WarpGrid* grid);
cudaError_t repack_params_to(const qmc_handle* h, const float* flat, float* padded, cudaStream_t st);
cudaError_t launch_energy(const qmc_handle* h, int hamiltonian, float field_h, const int8_t* spins,
                          int N, float* workspace, float* e_loc, double* moments, cudaStream_t st,
                          std::string& err);
cudaError_t launch_backward(const qmc_handle* h, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad, cudaStream_t st, std::string& err);

int sweep_slots(const qmc_handle* h, int S, int num_flips, EvalPlan* plan, WarpGrid* grid);
// layer-synchronous batched path (qmc_batched.cu)
bool batched_supported(const qmc_handle* h);
size_t batched_staging_floats(const qmc_handle* h, int n_items);
size_t batched_scratch_floats(int n_items);
cudaError_t launch_sweep_batched(const qmc_handle* h, const SweepArgs& s, cudaStream_t caller, std::string& err);
cudaError_t launch_energy_batched(const qmc_handle* h, const int8_t* spins, int N, const float* cache,
                                  float* scratch, int chunk_items, float2* terms, cudaStream_t st,
                                  std::string& err);
constexpr int kEnergyChunkItems = 32768;   // (sample, site) items evaluated per batched pass
cudaError_t launch_energy_finish(const qmc_handle* h, const int8_t* spins, int N, int hamiltonian, float field_h,
                                 const float2* partial, int nchunks, float* e_loc, double* moments,
                                 cudaStream_t st);
int energy_chunks(const qmc_handle* h);
size_t backward_workspace_floats(const qmc_handle* h, int N);

} // namespace qmc

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
    // tuning / cross-check knobs, from qmc_model_desc.reserved (include/qmcnn_b200.h: QMC_FLAG_*)
    bool allow_tiled = true;     // !QMC_FLAG_GENERIC_CONV: the specialised register-tiled conv instances
    int max_warps_override = 0;  // reserved[1]: caps the warps per CTA of the persistent kernels
    bool allow_ip = true;        // !QMC_FLAG_SWEEP_CLASSIC: the in-place persistent sweep kernel (k_sweep_ip) may be chosen
    bool force_ip = false;       // QMC_FLAG_SWEEP_INPLACE: use k_sweep_ip whenever the model is inside its coverage
    int ip_group = 4;            // reserved[2]: warps per phase group of k_sweep_ip<3>
    int ip_sync = 3;             // 0 (QMC_FLAG_IP_FREE_RUNNING): k_sweep_ip's warps run free; 3: a named barrier per
                                 // layer within each phase group of ip_group warps
    int ip_stagger = 40;         // reserved[2] >> 8: start offset between the phase groups of k_sweep_ip (x 1024 cycles)
    int ip_chunks = 64;          // reserved[3]: at most this many chunks per chain of the time-sliced sweep
    bool ip_cf = true;           // !QMC_FLAG_IP_ROWMAJOR_SITES: conflict-free site tables for the in-place evaluator
    qmc::site_t* d_ip_tab = nullptr;   // device image of the tables (ip_upload_tables), or nullptr
    int energy_path = 0;         // 0 auto, 1 classic persistent (QMC_FLAG_ENERGY_CLASSIC), 2 in-place (QMC_FLAG_ENERGY_INPLACE)
    bool backward_generic = false;   // QMC_FLAG_BACKWARD_GENERIC
    bool backward_smem_only = false; // QMC_FLAG_BACKWARD_SMEM: k_backward_smem / k_backward instead of the per-layer band kernels
    qmc::site_t* d_bwd_tab = nullptr;     // site tables of the band kernels (bwd_plane_upload_tables), or nullptr
    bool forward_blocked = false;    // QMC_FLAG_FORWARD_BLOCKED: k_forward (8 x 8 blocks) instead of k_forward_plane
    qmc::site_t* d_plane_tab = nullptr;   // site tables of k_forward_plane (plane_upload_tables), or nullptr
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
    // fewer tasks than one warp per slot: spread them over the SMs (64 chains are 64 CTAs of one warp, not 5 CTAs of 14 -
    // a warp that has its scheduler to itself runs a proposal ~3x faster than one of 3.5 sharing it)
    if (units < (long long)h->num_sms * best) {
        best = (int)((units + h->num_sms - 1) / h->num_sms);
        if (best < 1) best = 1;
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

// symmetry images as an extra grid dimension (blockIdx.y): float strides of the parameter blocks and the caches
struct ImageStrides { size_t params, cache; };

// launchers (each in its own .cu); return cudaError_t of the launch
cudaError_t repack_params(const qmc_handle* h, cudaStream_t st);
cudaError_t launch_forward(const qmc_handle* h, const int8_t* spins, int N, float* cache,
                           float* factors, float* logpsi, cudaStream_t st, std::string& err);
cudaError_t launch_forward_images(const qmc_handle* h, int nimg, const float* padded_blocks, const int8_t* spins, int N,
                                  float* cache, float* factors, float* logpsi, cudaStream_t st, std::string& err);
// qmc_plane.cu: the forward with the sample's planes resident in shared memory
bool forward_plane_supported(const qmc_handle* h);
cudaError_t plane_upload_tables(qmc_handle* h);
cudaError_t launch_forward_plane(const qmc_handle* h, int nimg, const float* padded_blocks, const int8_t* spins, int N,
                                 float* cache, float* factors, float* logpsi, cudaStream_t st);
// qmc_backward_plane.cu: the gradient layer by layer over all samples, row bands resident in shared memory
bool backward_plane_supported(const qmc_handle* h);
cudaError_t bwd_plane_upload_tables(qmc_handle* h);
size_t backward_plane_workspace_floats(const qmc_handle* h, int nimg, int N);
int backward_plane_ctas(const qmc_handle* h, int nimg, int N);
cudaError_t launch_backward_plane(const qmc_handle* h, int nimg, const float* blocks, const int8_t* spins, const float2* w,
                                  int N, const float* cache, float* gbuf, float* partial, float* grad, cudaStream_t st);
cudaError_t launch_sweep(const qmc_handle* h, const SweepArgs& a, cudaStream_t st, std::string& err);
struct IpLaunch { IpPlan ip; int warps, grid; size_t smem; bool ok; };
IpPlan ip_plan(const qmc_handle* h);
cudaError_t ip_upload_tables(qmc_handle* h);
IpLaunch ip_launch_plan(const qmc_handle* h, int S);
cudaError_t launch_sweep_ip(const qmc_handle* h, const SweepArgs& a, const IpLaunch& L, cudaStream_t st);
// launches and steps per chunk of the time-sliced in-place sweep (host only; qmc_diag_sweep_plan)
void ip_slice_counts(const qmc_handle* h, const IpLaunch& L, int S, long long n_steps, long long* launches, long long* chunk_len);
bool energy_ip_supported(const qmc_handle* h);
cudaError_t launch_energy_ip(const qmc_handle* h, const int8_t* spins, int N, const float* cache, float2* partial,
                             int nchunks, cudaStream_t st);
cudaError_t launch_sweep_sym(const qmc_handle* h, const SweepArgs& a, int nsym, double* drel, cudaStream_t st,
                             std::string& err);
int sweep_sym_slots(const qmc_handle* h, int S, int num_flips, int nsym, EvalPlan* plan, WarpGrid* grid);
cudaError_t repack_params_to(const qmc_handle* h, const float* flat, float* padded, cudaStream_t st, int nimg = 1);
cudaError_t launch_energy(const qmc_handle* h, int hamiltonian, float field_h, const int8_t* spins,
                          int N, float* workspace, float* e_loc, double* moments, cudaStream_t st,
                          std::string& err);
cudaError_t launch_backward(const qmc_handle* h, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad, cudaStream_t st, std::string& err);

int sweep_slots(const qmc_handle* h, int S, int num_flips, EvalPlan* plan, WarpGrid* grid);
cudaError_t launch_energy_finish(const qmc_handle* h, const int8_t* spins, int N, int hamiltonian, float field_h,
                                 const float2* partial, int nchunks, float* e_loc, double* moments,
                                 cudaStream_t st);
int energy_chunks(const qmc_handle* h);
size_t backward_workspace_floats(const qmc_handle* h, int N);
// symmetry images (qmc_set_image_params): every image in the same launches (image = blockIdx.y)
size_t backward_images_workspace_floats(const qmc_handle* h, int nimg, int N);
cudaError_t launch_backward_images(const qmc_handle* h, int nimg, const float* blocks, const int8_t* spins,
                                   const float* weights, int N, float* workspace, float* grad, cudaStream_t st,
                                   std::string& err);
size_t energy_sym_workspace_floats(const qmc_handle* h, int nsym, int N);
cudaError_t launch_energy_sym(const qmc_handle* h, int nsym, int hamiltonian, float field_h, const int8_t* spins, int N,
                              float* workspace, float* e_loc, double* moments, cudaStream_t st, std::string& err);
cudaError_t launch_sym_logrel(const qmc_handle* h, int nsym, int N, const float* caches, double* log_rel,
                              float* logpsi_sym, cudaStream_t st);

} // namespace qmc

// qmc_api.cu - the extern "C" boundary declared in include/qmcnn_b200.h.
#include <cstring>
#include <string>
#include "qmc_host.h"

using namespace qmc;

namespace qmc { unsigned long long g_launches = 0; }

static thread_local std::string g_create_err;

static int fail(qmc_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

static int cuda_fail(qmc_handle* h, cudaError_t e, const char* where) {
    std::string msg = std::string(where) + ": " + cudaGetErrorString(e);
    if (h && !h->err.empty() && e == cudaErrorInvalidValue) msg = h->err;   // launcher's own text
    if (h) h->err = msg; else g_create_err = msg;
    return e == cudaErrorInvalidValue ? QMC_ERR_UNSUPPORTED : QMC_ERR_CUDA;
}

static bool build_model(const qmc_model_desc* d, DevModel& m, std::string& err) {
    std::memset(&m, 0, sizeof(m));
    if (d->kind != QMC_MODEL_CRBM && d->kind != QMC_MODEL_DCRBM) { err = "unknown model kind"; return false; }
    if (d->k < 1 || d->k % 2 == 0) { err = "filter side k must be odd"; return false; }
    if (d->n_layers < 1 || d->n_layers > QMC_MAX_LAYERS) { err = "n_layers out of range"; return false; }
    if (d->kind == QMC_MODEL_CRBM && d->n_layers != 1) { err = "CRBM has exactly one layer"; return false; }
    if (d->Ly < 1 || d->Lx < 1) { err = "lattice sides must be positive"; return false; }
    for (int l = 0; l < d->n_layers; ++l)
        if (d->channels[l] < 1) { err = "channel counts must be positive"; return false; }
    if (d->channels[d->n_layers - 1] % 2) { err = "last layer needs an even channel count (Re/Im halves)"; return false; }
    m.kind = d->kind; m.k = d->k; m.p = (d->k - 1) / 2; m.D = d->n_layers;
    m.Ly = d->Ly; m.Lx = d->Lx; m.n = d->Ly * d->Lx;
    m.r = m.D * (m.k - 1) + 1;
    if (m.k > m.Ly || m.k > m.Lx) { err = "filter larger than the lattice"; return false; }
    int off = 0, soff = 0, coff = 0, cin = 1;
    m.bias_vis_off = -1; m.sp_vis_off = -1;
    for (int l = 0; l < m.D; ++l) {
        LayerInfo& L = m.layer[l];
        L.cin = cin; L.cout = d->channels[l];
        L.cinp = l == 0 ? 1 : round4(cin);
        L.coutp = round4(L.cout);
        L.w_off = off; off += m.k * m.k * L.cin * L.cout;
        if (d->kind == QMC_MODEL_CRBM) { m.bias_vis_off = off; off += 2; }   // models.py:19-28 order
        L.b_off = off; off += L.cout;
        L.sw_off = soff; soff += m.k * m.k * L.cin * L.coutp;
        L.sb_off = soff; soff += L.coutp;
        if (l < m.D - 1) { L.act_off = coff; coff += L.coutp * m.n; } else L.act_off = -1;
        cin = L.cout;
    }
    if (m.bias_vis_off >= 0) { m.sp_vis_off = soff; soff += 4; }
    m.P = off;
    m.smem_param_floats = round4(soff);
    m.fre_off = coff; coff += round4(m.n);
    m.fim_off = coff; coff += round4(m.n);
    m.cache_floats = coff;
    return true;
}

// tuning / cross-check knobs (include/qmcnn_b200.h: qmc_model_desc.reserved)
static void apply_tuning(qmc_handle* h, const qmc_model_desc* desc) {
    const int flags = desc->reserved[0];
    h->allow_tiled = !(flags & QMC_FLAG_GENERIC_CONV);
    h->allow_ip = !(flags & QMC_FLAG_SWEEP_CLASSIC);
    h->force_ip = (flags & QMC_FLAG_SWEEP_INPLACE) != 0;
    h->ip_sync = (flags & QMC_FLAG_IP_FREE_RUNNING) ? 0 : 3;
    h->ip_cf = !(flags & QMC_FLAG_IP_ROWMAJOR_SITES);
    h->energy_path = (flags & QMC_FLAG_ENERGY_CLASSIC) ? 1 : (flags & QMC_FLAG_ENERGY_INPLACE) ? 2 : 0;
    h->backward_generic = (flags & QMC_FLAG_BACKWARD_GENERIC) != 0;
    h->forward_blocked = (flags & QMC_FLAG_FORWARD_BLOCKED) != 0;
    h->backward_smem_only = (flags & QMC_FLAG_BACKWARD_SMEM) != 0;
    h->max_warps_override = desc->reserved[1] > 0 ? desc->reserved[1] : 0;
    h->ip_group = (desc->reserved[2] & 0xFF) > 0 ? (desc->reserved[2] & 0xFF) : 4;
    {   // phase-group start offset of k_sweep_ip in units of 1024 cycles: 0 = default (40), 0xFFFF = none
        const int st = (desc->reserved[2] >> 8) & 0xFFFF;
        h->ip_stagger = st == 0 ? 40 : st == 0xFFFF ? 0 : st;
    }
    h->ip_chunks = desc->reserved[3] > 0 ? desc->reserved[3] : 64;
}

// fresh lattices: one Philox block = 128 spins (oracle/philox.py: initial_spins)
__global__ void k_init_spins(int8_t* __restrict__ spins, int S, int n, unsigned long long seed, long long chain_id0,
                             unsigned reset_word) {
    const int nblk = (n + 127) / 128;
    const long long total = (long long)S * nblk;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int chain = (int)(t / nblk), blk = (int)(t - (long long)chain * nblk);
        const unsigned long long g = (unsigned long long)(chain_id0 + chain);
        const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, reset_word, (uint32_t)g, (uint32_t)(g >> 32)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        int8_t* dst = spins + (size_t)chain * n + (size_t)blk * 128;
        const int cnt = min(128, n - blk * 128);
        for (int i = 0; i < cnt; ++i) dst[i] = ((w[i >> 5] >> (i & 31)) & 1u) ? 1 : -1;
    }
}

extern "C" {

int qmc_init_spins(int device, int8_t* spins, int S, int n, uint64_t seed, int64_t chain_id0, int64_t reset_index,
                   void* stream) {
    if (S < 0 || n < 1 || (S > 0 && !spins)) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "init_spins: bad argument");
    if (S == 0) return QMC_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device && cudaSetDevice(device) != cudaSuccess) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "bad device ordinal");
    const long long total = (long long)S * ((n + 127) / 128);
    const int blocks = (int)((total + 127) / 128 < 4096 ? (total + 127) / 128 : 4096);
    ++g_launches;
    k_init_spins<<<blocks, 128, 0, (cudaStream_t)stream>>>(spins, S, n, seed, chain_id0,
                                                          0x80000000u | (uint32_t)(reset_index & 0x7FFFFFFF));
    const cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    return e == cudaSuccess ? QMC_OK : cuda_fail(nullptr, e, "init_spins");
}

int qmc_diag_sweep_plan(const qmc_model_desc* desc, int S, int num_flips, int64_t n_steps, int num_sms, size_t max_smem,
                        int64_t* out) {
    if (!desc || !out || S < 1 || n_steps < 1 || num_sms < 1) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "sweep_plan: bad argument");
    if (num_flips < 1 || num_flips > QMC_MAX_FLIPS) return fail(nullptr, QMC_ERR_UNSUPPORTED, "sweep_plan: num_flips must be 1 or 2");
    qmc_handle h;                        // host-only image of a handle: no CUDA call below
    std::string err;
    if (!build_model(desc, h.m, err)) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, err);
    h.desc = *desc;
    h.num_sms = num_sms;
    h.max_smem = max_smem;
    apply_tuning(&h, desc);
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if ((size_t)h.m.smem_param_floats * 4 > max_smem) return fail(nullptr, QMC_ERR_UNSUPPORTED, "parameters do not fit in shared memory");
    IpLaunch il{};
    if (num_flips == 1 && (il = ip_launch_plan(&h, S)).ok) {
        long long launches = 0, chunk = 0;
        ip_slice_counts(&h, il, S, n_steps, &launches, &chunk);
        out[0] = QMC_PLAN_SWEEP_IP; out[1] = il.grid; out[2] = il.warps; out[3] = (int64_t)il.smem;
        out[4] = launches; out[5] = chunk; out[6] = (int64_t)il.grid * il.warps;
        return QMC_OK;
    }
    EvalPlan pl; WarpGrid g;
    const int slots = sweep_slots(&h, S, num_flips, &pl, &g);
    if (slots < 0) { out[0] = QMC_PLAN_NONE; return QMC_OK; }     // Sampler falls through to qmc_nd_sweep
    out[0] = g.warps <= 8 ? QMC_PLAN_SWEEP_W8 : g.warps <= 16 ? QMC_PLAN_SWEEP_W16 : QMC_PLAN_SWEEP_W28;
    out[1] = g.grid; out[2] = g.warps; out[3] = (int64_t)g.smem; out[4] = 1; out[5] = n_steps; out[6] = slots;
    return QMC_OK;
}

int qmc_diag_fastdiv_check(unsigned long long* mismatches) {
    if (!mismatches) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "fastdiv_check: null output");
    // host image of the constant-memory table (the same constexpr object the device copy is initialised from)
    constexpr qmc::FastDivTable table = qmc::FastDivTable();
    unsigned long long bad = 0;
    auto mulhi = [](unsigned x, unsigned m) { return (unsigned)(((unsigned long long)x * m) >> 32); };
    for (int d = 1; d < 65536; ++d) {
        const unsigned M = qmc::fastdiv_magic(d);
        if (d < qmc::kFastDivTable && table.v[d] != M) ++bad;
        // FastDiv::div(x) = d > 1 ? umulhi(x, M) : x, for every x < 65536 at and around the multiples of d
        for (int q = 0; q * d < 65536 + d; ++q)
            for (int x = q * d - 1; x <= q * d + 1; ++x) {
                if (x < 0 || x > 65535) continue;
                const int got = d > 1 ? (int)mulhi((unsigned)x, M) : x;
                if (got != x / d) ++bad;
            }
    }
    // ... and exhaustively for the divisors the table serves
    for (int d = 1; d < qmc::kFastDivTable; ++d)
        for (int x = 0; x < 65536; ++x)
            if ((d > 1 ? (int)mulhi((unsigned)x, table.v[d]) : x) != x / d) ++bad;
    *mismatches = bad;
    return QMC_OK;
}

const char* qmc_version(void) {
#if QMC_DEBUG
    return "qmcnn_b200 0.2 (sm_100a, DEBUG build: device-side bounds checks)";
#else
    return "qmcnn_b200 0.2 (sm_100a)";
#endif
}

unsigned long long qmc_launch_count(void) { return qmc::g_launches; }

int qmc_create(qmc_handle** out, int device, const qmc_model_desc* desc) {
    if (!out || !desc) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "null argument");
    *out = nullptr;
    DevModel m;
    std::string err;
    if (!build_model(desc, m, err)) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, err);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, QMC_ERR_NO_DEVICE, "no CUDA device: the qmcnn_b200 hot path has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, QMC_ERR_BAD_ARGUMENT, "bad device ordinal");
    int prev = 0;
    cudaGetDevice(&prev);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    qmc_handle* h = new qmc_handle();
    h->device = device; h->desc = *desc; h->m = m;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    h->num_sms = prop.multiProcessorCount;
    h->max_smem = prop.sharedMemPerBlockOptin;
    apply_tuning(h, desc);
    e = cudaMalloc(&h->d_params, sizeof(float) * (size_t)m.P);
    if (e == cudaSuccess) e = cudaMemset(h->d_params, 0, sizeof(float) * (size_t)m.P);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_params_padded, sizeof(float) * (size_t)m.smem_param_floats);
    if (e == cudaSuccess) e = cudaMemset(h->d_params_padded, 0, sizeof(float) * (size_t)m.smem_param_floats);
    if (e == cudaSuccess) e = ip_upload_tables(h);
    if (e == cudaSuccess) e = plane_upload_tables(h);
    if (e == cudaSuccess) e = bwd_plane_upload_tables(h);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { qmc_destroy(h); return cuda_fail(nullptr, e, "cudaMalloc(params)"); }
    if ((size_t)m.smem_param_floats * 4 > h->max_smem) {
        qmc_destroy(h);
        return fail(nullptr, QMC_ERR_UNSUPPORTED, "parameters do not fit in shared memory");
    }
    *out = h;
    return QMC_OK;
}

int qmc_destroy(qmc_handle* h) {
    if (!h) return QMC_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    cudaFree(h->d_params);
    cudaFree(h->d_params_padded);
    cudaFree(h->d_sym_padded);
    cudaFree(h->d_ip_tab);
    cudaFree(h->d_plane_tab);
    cudaFree(h->d_bwd_tab);
    cudaSetDevice(prev);
    delete h;
    return QMC_OK;
}

const char* qmc_last_error(const qmc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
size_t qmc_num_params(const qmc_handle* h) { return h ? (size_t)h->m.P : 0; }
int qmc_receptive_field(const qmc_handle* h) { return h ? h->m.r : 0; }
size_t qmc_cache_floats(const qmc_handle* h) { return h ? (size_t)h->m.cache_floats : 0; }

size_t qmc_sweep_workspace_floats(const qmc_handle* h, int S, int num_flips) {
    if (!h || S < 1) return 0;
    EvalPlan pl;
    const int slots = sweep_slots(h, S, num_flips, &pl, nullptr);
    if (slots < 0) return 0;
    size_t f = (size_t)slots * pl.staging_floats;
    if (num_flips == 1) {
        const IpLaunch il = ip_launch_plan(h, S);
        if (il.ok) {
            const size_t b = (size_t)il.grid * il.warps * il.ip.staging_floats;
            if (b > f) f = b;
        }
    }
    return f ? f : 4;
}

size_t qmc_energy_workspace_floats(const qmc_handle* h, int N) {
    if (!h || N < 1) return 0;
    return (size_t)N * h->m.cache_floats + (size_t)N * energy_chunks(h) * 2;
}

size_t qmc_backward_workspace_floats(const qmc_handle* h, int N) {
    if (!h || N < 1) return 0;
    return backward_workspace_floats(h, N);
}

#define QMC_ENTER(h)                                                       \
    if (!(h)) return QMC_ERR_BAD_ARGUMENT;                                 \
    (h)->err.clear();                                                      \
    int prev_dev_ = 0;                                                     \
    cudaGetDevice(&prev_dev_);                                             \
    if (prev_dev_ != (h)->device) cudaSetDevice((h)->device);
#define QMC_LEAVE(h) if (prev_dev_ != (h)->device) cudaSetDevice(prev_dev_);

int qmc_set_params(qmc_handle* h, const float* params, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (!params) rc = fail(h, QMC_ERR_BAD_ARGUMENT, "null params");
    else {
        cudaError_t e = cudaMemcpyAsync(h->d_params, params, sizeof(float) * (size_t)h->m.P,
                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (e == cudaSuccess) e = repack_params(h, (cudaStream_t)stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "set_params");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_get_params(qmc_handle* h, float* params, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (!params) rc = fail(h, QMC_ERR_BAD_ARGUMENT, "null params");
    else {
        cudaError_t e = cudaMemcpyAsync(params, h->d_params, sizeof(float) * (size_t)h->m.P,
                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "get_params");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_logpsi_forward(qmc_handle* h, const int8_t* spins, int N, float* cache, float* factors,
                       float* logpsi, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (N < 0 || (N > 0 && (!spins || !cache))) rc = fail(h, QMC_ERR_BAD_ARGUMENT, "forward: null spins/cache");
    else if (N > 0) {
        cudaError_t e = launch_forward(h, spins, N, cache, factors, logpsi, (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "logpsi_forward");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_metropolis_sweep(qmc_handle* h, int8_t* spins, float* cache, float* workspace, int S,
                         int num_flips, int64_t step0, int64_t n_steps, const int32_t* flip_pos,
                         const float* uniforms, uint64_t seed, int64_t chain_id0, int64_t therm_its,
                         int64_t its_per_sample, int8_t* samples, int64_t n_sample_slots,
                         uint8_t* accept_trace,
                         float* logratio_trace, unsigned long long* n_accept, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (S < 0 || n_steps < 0 || (S > 0 && (!spins || !cache || !workspace)))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep: null spins/cache/workspace");
    else if (num_flips < 1 || num_flips > QMC_MAX_FLIPS)
        rc = fail(h, QMC_ERR_UNSUPPORTED, "sweep: num_flips must be 1 or 2");
    else if ((flip_pos == nullptr) != (uniforms == nullptr))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep: flip_pos and uniforms must both be given or both be NULL");
    else if (samples && (its_per_sample < 1 || n_sample_slots < 1))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep: its_per_sample and n_sample_slots must be positive");
    else if (S > 0 && n_steps > 0) {
        SweepArgs a{spins, cache, workspace, S, num_flips, step0, n_steps, flip_pos, uniforms,
                    seed, chain_id0, therm_its, its_per_sample > 0 ? its_per_sample : 1, samples,
                    n_sample_slots, accept_trace, logratio_trace, n_accept};
        cudaError_t e;
        IpLaunch il{};
        if (num_flips == 1 && (il = ip_launch_plan(h, S)).ok)
            e = launch_sweep_ip(h, a, il, (cudaStream_t)stream);
        else
            e = launch_sweep(h, a, (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "metropolis_sweep");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_set_image_params(qmc_handle* h, int nsym, const float* params_images, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (nsym < 1 || nsym > 8 || !params_images) rc = fail(h, QMC_ERR_BAD_ARGUMENT, "set_image_params: nsym must be 1..8");
    else {
        cudaError_t e = cudaSuccess;
        if (h->nsym != nsym) {
            cudaFree(h->d_sym_padded);
            h->d_sym_padded = nullptr;
            e = cudaMalloc(&h->d_sym_padded, sizeof(float) * (size_t)nsym * h->m.smem_param_floats);
            h->nsym = e == cudaSuccess ? nsym : 0;
        }
        if (e == cudaSuccess) e = repack_params_to(h, params_images, h->d_sym_padded, (cudaStream_t)stream, nsym);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "set_image_params");
    }
    QMC_LEAVE(h);
    return rc;
}

size_t qmc_sym_sweep_workspace_floats(const qmc_handle* h, int S, int num_flips, int nsym) {
    if (!h || S < 1) return 0;
    EvalPlan pl;
    const int slots = sweep_sym_slots(h, S, num_flips, nsym, &pl, nullptr);
    if (slots < 0) return 0;
    const size_t f = (size_t)slots * nsym * pl.staging_floats;
    return f ? f : 4;
}

int qmc_metropolis_sweep_sym(qmc_handle* h, int nsym, int8_t* spins, float* caches, double* log_rel,
                             float* workspace, int S, int num_flips, int64_t step0, int64_t n_steps,
                             const int32_t* flip_pos, const float* uniforms, uint64_t seed, int64_t chain_id0,
                             int64_t therm_its, int64_t its_per_sample, int8_t* samples, int64_t n_sample_slots,
                             uint8_t* accept_trace, float* logratio_trace, unsigned long long* n_accept,
                             void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (S < 0 || n_steps < 0 || (S > 0 && (!spins || !caches || !workspace || !log_rel)))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep_sym: null argument");
    else if (nsym != h->nsym || !h->d_sym_padded)
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep_sym: call qmc_set_image_params with the same nsym first");
    else if (num_flips < 1 || num_flips > QMC_MAX_FLIPS)
        rc = fail(h, QMC_ERR_UNSUPPORTED, "sweep_sym: num_flips must be 1 or 2");
    else if ((flip_pos == nullptr) != (uniforms == nullptr))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep_sym: flip_pos and uniforms must both be given or both be NULL");
    else if (samples && (its_per_sample < 1 || n_sample_slots < 1))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "sweep_sym: its_per_sample and n_sample_slots must be positive");
    else if (S > 0 && n_steps > 0) {
        SweepArgs a{spins, caches, workspace, S, num_flips, step0, n_steps, flip_pos, uniforms,
                    seed, chain_id0, therm_its, its_per_sample > 0 ? its_per_sample : 1, samples,
                    n_sample_slots, accept_trace, logratio_trace, n_accept};
        cudaError_t e = launch_sweep_sym(h, a, nsym, log_rel, (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "metropolis_sweep_sym");
    }
    QMC_LEAVE(h);
    return rc;
}

static int sym_ready(qmc_handle* h, int nsym, const char* what) {
    if (nsym < 1 || nsym != h->nsym || !h->d_sym_padded)
        return fail(h, QMC_ERR_BAD_ARGUMENT, std::string(what) + ": call qmc_set_image_params with the same nsym first");
    return QMC_OK;
}

size_t qmc_sym_energy_workspace_floats(const qmc_handle* h, int nsym, int N) {
    if (!h || N < 1 || nsym < 1) return 0;
    return energy_sym_workspace_floats(h, nsym, N);
}

size_t qmc_sym_backward_workspace_floats(const qmc_handle* h, int nsym, int N) {
    if (!h || N < 1 || nsym < 1) return 0;
    return backward_images_workspace_floats(h, nsym, N);
}

int qmc_logpsi_forward_sym(qmc_handle* h, int nsym, const int8_t* spins, int N, float* caches, double* log_rel,
                           float* logpsi_sym, void* stream) {
    QMC_ENTER(h);
    int rc = sym_ready(h, nsym, "forward_sym");
    if (rc == QMC_OK && (N < 0 || (N > 0 && (!spins || !caches)))) rc = fail(h, QMC_ERR_BAD_ARGUMENT, "forward_sym: null spins/caches");
    if (rc == QMC_OK && N > 0) {
        cudaError_t e = launch_forward_images(h, nsym, h->d_sym_padded, spins, N, caches, nullptr, nullptr,
                                              (cudaStream_t)stream, h->err);
        if (e == cudaSuccess && (log_rel || logpsi_sym))
            e = launch_sym_logrel(h, nsym, N, caches, log_rel, logpsi_sym, (cudaStream_t)stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "logpsi_forward_sym");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_local_energy_sym(qmc_handle* h, int nsym, int hamiltonian, float field_h, const int8_t* spins, int N,
                         float* workspace, float* e_loc, double* moments, void* stream) {
    QMC_ENTER(h);
    int rc = sym_ready(h, nsym, "local_energy_sym");
    if (rc == QMC_OK && (N < 0 || (N > 0 && (!spins || !workspace || !e_loc))))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "local_energy_sym: null spins/workspace/e_loc");
    else if (rc == QMC_OK && hamiltonian != QMC_HAMILTONIAN_TFIM && hamiltonian != QMC_HAMILTONIAN_HEISENBERG)
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "local_energy_sym: unknown hamiltonian");
    if (rc == QMC_OK && N > 0) {
        cudaError_t e = launch_energy_sym(h, nsym, hamiltonian, field_h, spins, N, workspace, e_loc, moments,
                                          (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "local_energy_sym");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_logpsi_backward_sym(qmc_handle* h, int nsym, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad_images, void* stream) {
    QMC_ENTER(h);
    int rc = sym_ready(h, nsym, "backward_sym");
    if (rc == QMC_OK && (N < 0 || (N > 0 && (!spins || !weights || !workspace)) || !grad_images))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "backward_sym: null argument");
    if (rc == QMC_OK && N > 0) {
        cudaError_t e = launch_backward_images(h, nsym, h->d_sym_padded, spins, weights, N, workspace, grad_images,
                                               (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "logpsi_backward_sym");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_local_energy(qmc_handle* h, int hamiltonian, float field_h, const int8_t* spins, int N,
                     float* workspace, float* e_loc, double* moments, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (N < 0 || (N > 0 && (!spins || !workspace || !e_loc)))
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "local_energy: null spins/workspace/e_loc");
    else if (hamiltonian != QMC_HAMILTONIAN_TFIM && hamiltonian != QMC_HAMILTONIAN_HEISENBERG)
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "local_energy: unknown hamiltonian");
    else if (N > 0) {
        cudaError_t e = launch_energy(h, hamiltonian, field_h, spins, N, workspace, e_loc, moments,
                                      (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "local_energy");
    }
    QMC_LEAVE(h);
    return rc;
}

int qmc_logpsi_backward(qmc_handle* h, const int8_t* spins, const float* weights, int N,
                        float* workspace, float* grad, void* stream) {
    QMC_ENTER(h);
    int rc = QMC_OK;
    if (N < 0 || (N > 0 && (!spins || !weights || !workspace)) || !grad)
        rc = fail(h, QMC_ERR_BAD_ARGUMENT, "backward: null argument");
    else if (N > 0) {
        cudaError_t e = launch_backward(h, spins, weights, N, workspace, grad, (cudaStream_t)stream, h->err);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "logpsi_backward");
    }
    QMC_LEAVE(h);
    return rc;
}

} // extern "C"

/*
 * qmcnn_b200.h - C ABI of the B200-native variational-Monte-Carlo hot path.
 *
 * This is the drop-in boundary for the one path of dmaloneynygc/qmcnn that is
 * accelerated: the conv-wavefunction log-psi forward, the batched Metropolis
 * sweep, the local-energy estimators and the log-psi gradient.  The reference
 * has no FFI of its own (pure TensorFlow-1 Python); each entry point below
 * names the reference Python call it replaces (file:line under the reference
 * root) and is what `qmcnn_b200/_lib.py` binds with ctypes.  INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller unless it says
 *    "host".  The library allocates only handle-private memory (parameter
 *    copy), freed by qmc_destroy.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*).  No
 *    call synchronises the device, throws or aborts.
 *  - Return 0 on success, negative qmc_status on failure; text via
 *    qmc_last_error().  A handle is bound to one device and is not thread-safe
 *    (the reference drives one session with parallel_iterations=1,
 *    sampler.py:172).
 *  - Spins are int8 +-1, UN-padded, row-major [n, Ly, Lx]; periodic wrap is
 *    index arithmetic inside the kernels (replaces helpers.py:73-91 pad/unpad).
 *  - Parameters are ONE flat fp32 vector in the reference's variable creation
 *    order and HWIO filter layout:
 *      CRBM  (models.py:19-28):  filters[k,k,1,2a], bias_vis[2], bias_hid[2a]
 *      DCRBM (models.py:85-92):  filters_0[k,k,1,C1], bias_0[C1], filters_1[k,k,C1,C2], ...
 *  - Complex outputs are interleaved (re, im) fp32 pairs == numpy complex64.
 */
#ifndef QMCNN_B200_H
#define QMCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMC_MAX_LAYERS 16
#define QMC_MAX_FLIPS 2

typedef enum {
    QMC_OK = 0,
    QMC_ERR_BAD_ARGUMENT = -1,
    QMC_ERR_UNSUPPORTED = -2, /* shape outside what the kernels cover */
    QMC_ERR_CUDA = -3,
    QMC_ERR_NO_DEVICE = -4
} qmc_status;

enum { QMC_MODEL_CRBM = 0, QMC_MODEL_DCRBM = 1 };
enum { QMC_HAMILTONIAN_TFIM = 0, QMC_HAMILTONIAN_HEISENBERG = 1 };

/* host struct describing models.py CRBM(k, pad, alpha, 2) / DCRBM(k, layers, 2)
 * on a periodic Ly x Lx lattice. */
typedef struct {
    int32_t kind;                     /* QMC_MODEL_* */
    int32_t k;                        /* filter side, odd */
    int32_t n_layers;                 /* D; 1 for CRBM */
    int32_t channels[QMC_MAX_LAYERS]; /* C_1..C_D; CRBM: channels[0] = 2*alpha.
                                         channels[D-1] must be even */
    int32_t Ly, Lx;                   /* system_shape */
    int32_t reserved[4];              /* zero = defaults.  Tuning / cross-check knobs (they never change a result
                                         beyond fp32 rounding of a final sum; tests use them to run every
                                         decomposition):
                                           [0] QMC_FLAG_* bits
                                           [1] max warps per CTA of the persistent kernels (0 = as many as fit)
                                           [2] bits 0-7: warps per phase group of k_sweep_ip (0 = 4); bits 8-23: start
                                               offset between the phase groups in units of 1024 cycles (0 = 40,
                                               0xFFFF = none)
                                           [3] max chunks per chain of the time-sliced sweep (0 = 64) */
} qmc_model_desc;

enum {
    QMC_FLAG_GENERIC_CONV = 1,      /* generic conv loop instead of the register-tiled FFMA2 instances */
    QMC_FLAG_SWEEP_CLASSIC = 2,     /* sweep: never the in-place kernel (k_sweep_ip) */
    QMC_FLAG_SWEEP_INPLACE = 4,     /* sweep: k_sweep_ip whenever the model is inside its coverage */
    QMC_FLAG_IP_FREE_RUNNING = 8,   /* k_sweep_ip without phase-group barriers (diagnosis) */
    QMC_FLAG_ENERGY_CLASSIC = 16,   /* TFIM local energy: classic persistent kernel */
    QMC_FLAG_ENERGY_INPLACE = 32,   /* TFIM local energy: k_energy_ip whenever covered */
    QMC_FLAG_BACKWARD_GENERIC = 64, /* gradient: k_backward (planes in L2) instead of k_backward_smem */
    QMC_FLAG_IP_ROWMAJOR_SITES = 128, /* in-place evaluator: row-major site order instead of the conflict-free deal */
    QMC_FLAG_FORWARD_BLOCKED = 256, /* forward: k_forward (8 x 8 blocks through per-warp tiles) instead of k_forward_plane */
    QMC_FLAG_BACKWARD_SMEM = 512    /* gradient: the per-sample kernels (k_backward_smem / k_backward) instead of the
                                       per-layer band kernels (qmc_backward_plane.cu) */
};

typedef struct qmc_handle qmc_handle;

/* models.py:11-28 / 75-92 (variable creation). `device` is a CUDA ordinal. */
int qmc_create(qmc_handle** out, int device, const qmc_model_desc* desc /*host*/);
int qmc_destroy(qmc_handle* h);
const char* qmc_last_error(const qmc_handle* h); /* host string; NULL handle -> last create error */

size_t qmc_num_params(const qmc_handle* h);        /* P */
int qmc_receptive_field(const qmc_handle* h);      /* r = D(k-1)+1 */
/* floats of activation cache per chain/sample (opaque layout) */
size_t qmc_cache_floats(const qmc_handle* h);
/* floats of scratch qmc_metropolis_sweep needs for S chains */
size_t qmc_sweep_workspace_floats(const qmc_handle* h, int S, int num_flips);
/* floats of scratch qmc_local_energy / qmc_logpsi_backward need for N samples */
size_t qmc_energy_workspace_floats(const qmc_handle* h, int N);
size_t qmc_backward_workspace_floats(const qmc_handle* h, int N);

/* tf.assign of the model variables (models.py:19-28, 85-92): copy P floats. */
int qmc_set_params(qmc_handle* h, const float* params, void* stream);
int qmc_get_params(qmc_handle* h, float* params, void* stream);

/* model.factors(pad(x)) and its site sum (models.py:31-67, 95-131).
 *   cache   [N * qmc_cache_floats]  (required; filled as a by-product and
 *                                    reusable by the sweep / energy / backward)
 *   factors [N, Ly*Lx] complex64 or NULL
 *   logpsi  [N] complex64 or NULL */
int qmc_logpsi_forward(qmc_handle* h, const int8_t* spins, int N, float* cache,
                       float* factors, float* logpsi, void* stream);

/* Sampler.mcmc_step x n_steps (sampler.py:104-155) for S independent chains,
 * persistent in-kernel, incremental receptive-field update per proposal.
 *   spins   [S, Ly*Lx] in/out; cache [S * qmc_cache_floats] in/out (must hold
 *           the forward of `spins` under the CURRENT parameters: call
 *           qmc_logpsi_forward first, as mcmc_reset does, sampler.py:85-88)
 *   steps   step0 .. step0+n_steps-1 (global step index i of the while loop)
 *   flip_pos [n_steps, S, num_flips] int32 and uniforms [n_steps, S] fp32
 *           (sampler.py:95-100), or BOTH NULL -> in-kernel Philox-4x32-10
 *           keyed (seed; chain_id0 + chain, step) - see oracle/philox.py
 *   samples [n_sample_slots, S, Ly*Lx] int8 or NULL; sample j =
 *           (i-therm_its)/its_per_sample written after the update when
 *           i >= therm_its, (i-therm_its) % its_per_sample == 0
 *           (sampler.py:135-152) and j < n_sample_slots (= samples_per_sampler)
 *   accept_trace [n_steps, S] uint8 or NULL; logratio_trace [n_steps, S] fp32
 *           (Re sum(f' - f)) or NULL; n_accept [1] uint64 (+=) or NULL.
 * The call enqueues one or more kernel launches on `stream` (deep models with more
 * chains than resident warp slots are time-sliced into full-wave launches); which
 * kernel runs never changes a result bit. */
int qmc_metropolis_sweep(qmc_handle* h, int8_t* spins, float* cache, float* workspace,
                         int S, int num_flips, int64_t step0, int64_t n_steps,
                         const int32_t* flip_pos, const float* uniforms,
                         uint64_t seed, int64_t chain_id0,
                         int64_t therm_its, int64_t its_per_sample, int8_t* samples,
                         int64_t n_sample_slots, uint8_t* accept_trace, float* logratio_trace,
                         unsigned long long* n_accept, void* stream);

/* Fresh chains of Sampler.mcmc_reset (sampler.py:74-79: iid uniform +-1 lattices).  Stateless; spins
 * [S, n] int8 on `device`.  Chain c gets a function of (seed, chain_id0 + c, reset_index) only -
 * Philox-4x32-10 on counters the proposal stream never uses (oracle/philox.py: initial_spins) - so a
 * chain's start does not depend on how the chains are sharded over ranks. */
int qmc_init_spins(int device, int8_t* spins, int S, int n, uint64_t seed, int64_t chain_id0,
                   int64_t reset_index, void* stream);

/* Symmetry-averaged amplitude psi_sym(s) = (1/nsym) sum_g psi(s; W o g), g in D4
 * (symmetry.ipynb cell 0 defines the group; SURVEY.md section 8 the amplitude - the reference
 * has no amplitude code).  params_images [nsym, P]: the flat parameter vector of every image
 * (filters of every layer rotated / mirrored, biases unchanged), nsym <= 8. */
int qmc_set_image_params(qmc_handle* h, int nsym, const float* params_images, void* stream);
size_t qmc_sym_sweep_workspace_floats(const qmc_handle* h, int S, int num_flips, int nsym);
/* qmc_metropolis_sweep for |psi_sym|^2.  caches [nsym, S, qmc_cache_floats] (forward of `spins`
 * under each image), log_rel [S, nsym, 2] fp64 in/out = log psi_g - log psi_0 (Re, Im);
 * logratio_trace receives log|psi_sym(s')/psi_sym(s)|.  Other arguments as qmc_metropolis_sweep. */
int qmc_metropolis_sweep_sym(qmc_handle* h, int nsym, int8_t* spins, float* caches, double* log_rel,
                             float* workspace, int S, int num_flips, int64_t step0, int64_t n_steps,
                             const int32_t* flip_pos, const float* uniforms, uint64_t seed,
                             int64_t chain_id0, int64_t therm_its, int64_t its_per_sample,
                             int8_t* samples, int64_t n_sample_slots, uint8_t* accept_trace,
                             float* logratio_trace, unsigned long long* n_accept, void* stream);

/* The symmetry-averaged amplitude in the other kernels (after qmc_set_image_params with the same nsym).  The
 * images are an extra grid dimension of the SAME launches:
 *   forward_sym   caches [nsym, N, qmc_cache_floats] filled for every image; log_rel [N, nsym, 2] fp64 or NULL
 *                 (= log psi_g - log psi_0, what qmc_metropolis_sweep_sym carries); logpsi_sym [N] complex64 or NULL
 *                 = log((1/nsym) sum_g psi_g)                                                        (2 launches)
 *   local_energy_sym   E_loc[psi_sym] = sum_g p_g E_loc[psi_g], p_g = psi_g / sum_h psi_h           (3 launches)
 *   backward_sym  grad_images [nsym, P] (+=): image g's gradient for the weights w_n conj(p_gn); the caller adds
 *                 them into the one trainable vector through the image gather index                 (4 launches) */
size_t qmc_sym_energy_workspace_floats(const qmc_handle* h, int nsym, int N);
size_t qmc_sym_backward_workspace_floats(const qmc_handle* h, int nsym, int N);
int qmc_logpsi_forward_sym(qmc_handle* h, int nsym, const int8_t* spins, int N, float* caches, double* log_rel,
                           float* logpsi_sym, void* stream);
int qmc_local_energy_sym(qmc_handle* h, int nsym, int hamiltonian, float field_h, const int8_t* spins, int N,
                         float* workspace, float* e_loc, double* moments, void* stream);
int qmc_logpsi_backward_sym(qmc_handle* h, int nsym, const int8_t* spins, const float* weights, int N,
                            float* workspace, float* grad_images, void* stream);

/* ising_energy / heisenberg_energy (mcmc_tf.py:59-90, 93-141): local energy
 * PER SPIN of N samples; all Ly*Lx (TFIM) or 2*Ly*Lx (Heisenberg) connected
 * configurations evaluated as receptive-field deltas in one launch sequence.
 *   e_loc [N] complex64; moments [4] fp64 (+=): N, sum Re E, sum Im E,
 *   sum |E|^2, or NULL. */
int qmc_local_energy(qmc_handle* h, int hamiltonian, float field_h, const int8_t* spins,
                     int N, float* workspace, float* e_loc, double* moments, void* stream);

/* gradient of loss_op (mcmc_tf.py:35-56, 172-177) for given per-sample complex
 * weights w_n: grad[p] += sum_n Re[w_n * conj(d logpsi_n / d p)], P floats in
 * qmc_set_params order.  weights [N] complex64 = (E_n - mean E)/N reproduces
 * d loss_op / d p. */
int qmc_logpsi_backward(qmc_handle* h, const int8_t* spins, const float* weights, int N,
                        float* workspace, float* grad, void* stream);

/* ---- 1-D and 3-D lattices (models.py:56-61, 118-123: the conv1d / conv3d branches; sampler.py and
 * mcmc_tf.py are n_dims-generic) ----------------------------------------------------------------------
 * A separate, simple path: stateless entry points, one CTA per chain / sample running the reference's own
 * algorithm (a full network evaluation per proposal / connected configuration).  `params` is the flat
 * vector in the reference's variable order with filters [k]*n_dims + [C_in, C_out]; spins int8 +-1
 * un-padded [N, prod(L)], row-major over the lattice axes.  n_dims = 2 is accepted too (cross-check of the
 * tuned 2-D path). */
typedef struct {
    int32_t kind, k, n_layers;
    int32_t channels[QMC_MAX_LAYERS];
    int32_t n_dims;                   /* 1, 2 or 3 */
    int32_t L[3];                     /* system_shape, first n_dims entries */
    int32_t reserved[3];
} qmc_nd_desc;
const char* qmc_nd_last_error(void);
size_t qmc_nd_num_params(const qmc_nd_desc* d);
/* scratch floats for `units` chains / samples on `device` */
size_t qmc_nd_scratch_floats(const qmc_nd_desc* d, int device, int units);
/* model.factors / log psi: factors [N, n] complex64 or NULL, logpsi [N] complex64 or NULL */
int qmc_nd_forward(const qmc_nd_desc* d, int device, const float* params, const int8_t* spins, int N,
                   float* scratch, float* factors, float* logpsi, void* stream);
/* Sampler.mcmc_step x n_steps (sampler.py:104-155); cur_factors [S, n] complex64 in/out must hold
 * model.factors of `spins` under the current parameters (mcmc_reset, sampler.py:85-88); other arguments
 * as qmc_metropolis_sweep */
int qmc_nd_sweep(const qmc_nd_desc* d, int device, const float* params, int8_t* spins, float* cur_factors,
                 float* scratch, int S, int num_flips, int64_t step0, int64_t n_steps, const int32_t* flip_pos,
                 const float* uniforms, uint64_t seed, int64_t chain_id0, int64_t therm_its,
                 int64_t its_per_sample, int8_t* samples, int64_t n_sample_slots, uint8_t* accept_trace,
                 float* logratio_trace, unsigned long long* n_accept, void* stream);
/* ising_energy / heisenberg_energy per spin (mcmc_tf.py:59-141), e_loc [N] complex64 */
int qmc_nd_local_energy(const qmc_nd_desc* d, int device, int hamiltonian, float field_h, const float* params,
                        const int8_t* spins, int N, float* scratch, float* e_loc, void* stream);

/* gradient of loss_op (mcmc_tf.py:35-56, 172-177): grad[P] += sum_n Re[w_n conj(d log psi_n / d p)],
 * weights [N] complex64; deterministic */
size_t qmc_nd_backward_scratch_floats(const qmc_nd_desc* d, int device, int N);
int qmc_nd_logpsi_backward(const qmc_nd_desc* d, int device, const float* params, const int8_t* spins,
                           const float* weights, int N, float* scratch, float* grad, void* stream);

/* Diagnostics (synchronous, not on the hot path): measured FP32-FMA (TFLOP/s)
 * and MUFU ex2 (Gop/s) issue peaks of `device` - the roofline denominators
 * MEASURED_PEAKS.json does not carry. */
int qmc_diag_peaks(int device, double* fp32_tflops /*host*/, double* mufu_gops /*host*/);
/* same, plus the packed fma.rn.f32x2 (FFMA2) rate in TFLOP/s */
int qmc_diag_peaks2(int device, double* fp32_tflops, double* ffma2_tflops, double* mufu_gops);
/* Phase timers of k_sweep_ip since the last call - only in builds with -DQMC_IP_PROFILE=1 (scripts/build_variant.sh),
 * all zero otherwise.  out[0..7]: clock64 cycles summed over warps for draw + top barrier, spin tile + frame gathers,
 * layer 0, layer barriers, conv accumulation loops, tanh epilogues, head, accept + commit; out[8]: proposals;
 * out[9 + w]: task duration (cycles) of warp w of a CTA, summed over CTAs and launches. */
int qmc_diag_ip_profile(unsigned long long* out /*host, 21 entries*/);
/* Every float through the evaluator's tanh epilogue (small-argument fast path + tanhf) against tanhf: number of
 * bit-level mismatches (must be 0). */
int qmc_diag_tanh_check(int device, unsigned long long* mismatches /*host*/);

/* Host-only: the magic-number division every kernel's index arithmetic uses (FastDiv: x / d as one multiply-high for
 * x < 65536, d < 65536; magics of d < 512 from a constant-memory table) against integer division - every x for the
 * table's divisors, x around every multiple of d for the rest.  Number of mismatches (must be 0). */
int qmc_diag_fastdiv_check(unsigned long long* mismatches /*host*/);

/* Host-only (no CUDA call, works without a GPU): the launch qmc_metropolis_sweep would make for this model / lattice /
 * S chains / n_steps on a device with num_sms SMs and max_smem bytes of opt-in shared memory per CTA - the planner
 * behind Sampler.mcmc_op (sampler.py:158-177) as a testable function.  out[0] = kernel (QMC_PLAN_*), out[1] = CTAs,
 * out[2] = warps per CTA, out[3] = dynamic shared memory per CTA (bytes), out[4] = kernel launches for the n_steps,
 * out[5] = steps per task (the chunk length of the time-sliced in-place kernel; n_steps otherwise), out[6] = warp slots,
 * out[7] = 0.  QMC_PLAN_NONE: outside the incremental kernels' coverage (the Python Sampler then runs qmc_nd_sweep). */
#define QMC_PLAN_NONE (-1)
#define QMC_PLAN_SWEEP_W8 0   /* classic persistent kernel, <= 8 warps per CTA (255 registers) */
#define QMC_PLAN_SWEEP_W16 1  /* ... <= 16 warps (128 registers) */
#define QMC_PLAN_SWEEP_W28 2  /* ... <= 28 warps (72 registers; models with <= 8 channels per layer) */
#define QMC_PLAN_SWEEP_IP 3   /* in-place, time-sliced persistent kernel (k_sweep_ip) */
int qmc_diag_sweep_plan(const qmc_model_desc* desc, int S, int num_flips, int64_t n_steps, int num_sms, size_t max_smem,
                        int64_t* out /*host, 8 entries*/);

/* number of CUDA kernels this library has launched in this process (graph replays count
 * their kernel nodes) */
unsigned long long qmc_launch_count(void);

/* library build info, host string */
const char* qmc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QMCNN_B200_H */

/*
 * svbasl.h - C ABI of libsvbasl.so: B200 (sm_100a) kernels for the ASL
 * stochastic-variational-Bayes hot path of physimals/svb_models_asl.
 *
 * The reference exposes NO native interface for this path - it is a Python
 * plugin (entry-point group `svb.models`, /root/reference/setup.py:89-95) whose
 * arithmetic runs inside TensorFlow.  Each entry point below therefore cites
 * the reference *Python* interface whose work it takes over; the binding a
 * maintainer adds on the reference side is the ctypes stub in INTEGRATION.md
 * (mirrored by svb_models_asl_b200/_lib.py).
 *
 * Conventions
 *  - plain C, no torch / C++ types; return 0 on success, a negative
 *    SVBASL_E_* code otherwise; svbasl_last_error() gives the message
 *    (thread-local).  Nothing throws across the boundary.
 *  - the caller owns every buffer; device pointers unless the name says host.
 *    The library allocates nothing persistent (svbasl_step_host owns a small
 *    staging context created/destroyed explicitly).
 *  - all launches are asynchronous on the caller's `stream` (a cudaStream_t
 *    passed as void*; NULL = legacy default stream).
 *  - device arrays are SoA, voxel-fastest: a [K][W] array has row stride `ld`
 *    floats (ld >= n_vox), element (k, w) at base[k*ld + w].
 */
#ifndef SVBASL_H
#define SVBASL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVBASL_ABI_VERSION 2
#define SVBASL_MAX_PAR 10          /* P' = model parameters + noise */
#define SVBASL_MAX_SPATIAL 4
#define SVBASL_NN_HIDDEN 10        /* aslnn.py:238-240: 2 -> 10 -> 10 -> 1 */
#define SVBASL_NN_NWEIGHTS 151     /* W0[2][10] b0[10] W1[10][10] b1[10] W2[10] b2[1] */

/* error codes */
#define SVBASL_OK 0
#define SVBASL_E_INVALID (-1)      /* bad argument / inconsistent descriptor */
#define SVBASL_E_UNSUPPORTED (-2)  /* combination not compiled into the library */
#define SVBASL_E_CUDA (-3)         /* CUDA runtime error (message has the detail) */

/* model kinds: the three `svb.models` entry points (setup.py:91-93) */
#define SVBASL_MODEL_ASLREST 0     /* svb_models_asl/aslrest.py  AslRestModel */
#define SVBASL_MODEL_ASLREST_DISP 1/* svb_models_asl/aslrest_disp.py AslRestDisp */
#define SVBASL_MODEL_ASLNN 2       /* svb_models_asl/aslnn.py AslNNModel */

/* model flags (aslrest.py:24-67 options after resolution in __init__ :69-246) */
#define SVBASL_F_CASL 0x01         /* casl */
#define SVBASL_F_INFERATT 0x02     /* inferatt: delttiss (/deltwm/deltblood) are parameters */
#define SVBASL_F_INFERART 0x04     /* inferart: fblood (+deltblood) */
#define SVBASL_F_INCWM 0x08        /* incwm: add a WM tissue component */
#define SVBASL_F_INFERWM 0x10      /* inferwm: fwm (+deltwm, t1wm) are parameters; the WM SIGNAL needs INCWM (aslrest.py:327) */
#define SVBASL_F_INFERT1 0x20      /* infert1: t1 (+t1wm) are parameters */
#define SVBASL_F_ARTONLY 0x40      /* artonly: no tissue component */
#define SVBASL_F_DISP_INFER 0x80   /* aslrest_disp infer_disp_params: s, sp are parameters */
#define SVBASL_F_DISP_ASWRITTEN 0x100 /* reproduce gamma2-gamma2 == 0 (aslrest_disp.py:108) */
#define SVBASL_F_NN_TC 0x200       /* aslnn: the two 10x10 products per row of the fused step (the tf.matmul chain of
                                      aslnn.py:238-260 and its gradient) on the tensor cores: tcgen05.mma kind::f16 with
                                      hi/lo split operands staged in tensor memory, float32 accumulation
                                      (csrc/model_nn_tc.cuh).  Without the flag the FP32-pipe kernel runs; both give
                                      the same results to float32 rounding. */

/* parameter transforms (svb dist: Normal / LogNormal / FoldedNormal) */
#define SVBASL_XF_IDENTITY 0
#define SVBASL_XF_EXP 1
#define SVBASL_XF_ABS 2

/* prior types (get_parameter(prior_type=...), aslrest.py:237) */
#define SVBASL_PRIOR_N 0           /* fixed Normal */
#define SVBASL_PRIOR_ARD 1         /* "A": Normal with trainable per-voxel log precision */
#define SVBASL_PRIOR_MRF 2         /* "M": spatial Laplacian prior with trainable global log ak */

/* latent-loss form */
#define SVBASL_LATENT_NUMERIC 0    /* force_num_latent_loss=True (asl_example.py:41) */
#define SVBASL_LATENT_ANALYTIC 1   /* closed-form Gaussian KL */

/* Resolved model options: what AslRestModel/AslRestDisp/AslNNModel.__init__ compute
 * (aslrest.py:69-246, aslrest_disp.py:30-43, aslnn.py:61-88), as a flat POD. */
typedef struct svbasl_model {
    int32_t kind;                  /* SVBASL_MODEL_* */
    uint32_t flags;                /* SVBASL_F_* */
    float tau, t1b;                /* aslrest.py:27,50 */
    float t1, pc, fcalib, att;     /* GM tissue constants (aslrest.py:35-39,131-135) */
    float t1wm, pcwm, fcalibwm, attwm, fwm;   /* WM constants (aslrest.py:42-47) */
    float artt;                    /* arterial arrival time when not inferred (aslrest.py:86-87) */
    float leadscale;               /* aslrest.py:232 */
    float pvgm_s, pvwm_s;          /* partial volumes when uniform ... */
    const float *pvgm, *pvwm;      /* ... or per voxel [n_vox] (aslrest.py:110-114); NULL = use scalar */
    /* aslrest_disp.py:24-43 */
    float conv_dt, conv_tmax;
    int32_t conv_nt;
    float s_fixed, sp_fixed;       /* used when !DISP_INFER (aslrest_disp.py:88-90) */
    /* aslnn.py:229-241: 2->10 tanh ->10 tanh ->1, packed W0[2][10] b0[10] W1[10][10] b1[10] W2[10] b2[1]
     * in the layout of the reference's weights%i.npy / biases%i.npy files (aslnn.py:326-340) */
    const float *nn_weights;       /* HOST pointer, SVBASL_NN_NWEIGHTS floats; copied into the kernel arguments */
} svbasl_model;

/* Inference-engine description for one shard of voxels: what svb's SvbFit builds around the
 * plugin (posterior, noise, priors, loss). */
typedef struct svbasl_engine {
    int64_t n_vox;                 /* voxels this call processes: local indices [w_begin, w_begin + n_vox) */
    int64_t w_begin;               /* first local index processed (= size of the lower halo; 0 without halos) */
    int64_t ld;                    /* row stride of every [K][W] array, floats (>= w_begin + n_vox + upper halo) */
    int64_t vox_offset;            /* global index of LOCAL index 0 (RNG stream: draws depend on the global
                                      voxel only, so results do not depend on the sharding) */
    int64_t n_vox_global;          /* voxels in the whole volume */
    int32_t n_par;                 /* P' = model parameters + noise (must match the model) */
    int32_t n_samples;             /* S  (sample_size, asl_example.py:31) */
    int32_t n_batch;               /* B  time points per iteration (batch_size, asl_example.py:30) */
    int32_t t_full;                /* T  total time points (likelihood scale T/B) */
    int32_t latent;                /* SVBASL_LATENT_* */
    int32_t cov_llt;               /* 0: KL uses chol^T chol (svb); 1: chol chol^T */
    int32_t prior_type[SVBASL_MAX_PAR];
    float prior_mean[SVBASL_MAX_PAR];   /* internal-space prior moments, noise last */
    float prior_var[SVBASL_MAX_PAR];
    float ard_phi_max;             /* svb clips phi to [0, 1e6]; <= 0 disables the clip */
    float latent_weight;
    float grad_scale;              /* d(cost)/d(param) scale: 1/n_vox_global (svb minimises the MEAN cost) */
    /* posterior state [n_state][ld]: mean[P'], logvar[P'], offdiag[P'(P'-1)/2] rows (1,0),(2,0),(2,1),..,
     * then one log-phi row per ARD parameter */
    float *state;
    float *state_out;              /* where the updated state is written; NULL = in place (a voxel only ever reads
                                      its OWN state rows, also with spatial priors: neighbours are seen through
                                      spatial_samples) */
    /* data / time points.  Row r of the batch is full-array row t_row0 + r*t_row_stride (svb's strided
     * time-point mini-batches).  tpts may be NULL: then t = ti[row] + zoff[w] (aslrest.py:438-440) */
    const float *data;             /* [T][ld] */
    const float *tpts;             /* [T][ld] or NULL */
    const float *ti;               /* [T] (device) when tpts == NULL */
    const float *zoff;             /* [ld] slice offset z*slicedt, or NULL (= 0) */
    int32_t t_row0, t_row_stride;
    /* random draws: eps [P'][S][ld] read from memory when non-NULL (parity mode, the reference's
     * tf.random_normal is not reproducible); otherwise Philox2x32-10 keyed on (seed, step, global voxel), one call
     * = row j of samples 2k and 2k+1 (csrc/philox.h; svbasl_fill_eps writes the same stream out) */
    const float *eps;
    uint64_t seed;
    /* spatial prior ("M"): neighbour table [6][ld] of LOCAL voxel indices, -1 = none.  The local arrays
     * cover a contiguous global range [lower halo | owned | upper halo]; halo voxels are only read.
     * spatial_samples [n spatial params][S][ld]: theta samples of the spatially-regularised parameters of every
     * local voxel (halo included) for THIS step: written by svbasl_sample_spatial() (first iteration, or whenever
     * the state was set from outside) or by the previous fused step through spatial_samples_out. */
    const int32_t *neighbours;
    const float *spatial_samples;
    /* Where a fused step (svbasl_step / svbasl_step_spatial) writes the samples of the NEXT iteration: after the Adam
     * update each voxel draws theta(step+1) of its spatial parameters from its NEW state (the Philox stream of
     * step+1) and stores them here, so that no pre-pass kernel runs between iterations.  Must be a different buffer
     * from spatial_samples (other CTAs still read this step's samples): the caller ping-pongs two buffers.  NULL =
     * do not write (svbasl_elbo_grad, or callers that run svbasl_sample_spatial before every step).  Needs the
     * in-kernel draws (eps == NULL). */
    float *spatial_samples_out;
    const float *log_ak;           /* [n spatial params] device */
    double *ak_grad;               /* [n spatial params] device accumulators: d(sum cost)/d(log ak) */
    /* Iteration index from DEVICE memory (CUDA-graph replay: a captured launch cannot change its arguments).
     * When non-NULL it replaces the `step` argument / adam->step0 everywhere (draws, lr_t index, batch rows) and
     * `cost_sum` is then the BASE of a per-iteration history indexed by that counter.  svbasl_step_spatial (fused
     * hyper step), svbasl_hyper_step_dev / svbasl_advance_step increment it at the end of an iteration. */
    const long long *step_dev;
    int32_t cost_sum_scalar;       /* non-zero: cost_sum stays a plain accumulator even with step_dev (svbasl_step_host) */
    /* Spatial prior across GPUs, communication fused into the step kernel: a shard-boundary voxel stores its
     * next-iteration samples ALSO into the adjacent rank's halo columns over NVLink peer memory (pointers obtained
     * by CUDA IPC), so no halo exchange is launched and nothing but S floats per boundary voxel and spatial
     * parameter crosses the link.  peer_lo / peer_hi = the lower / upper neighbour's buffer that plays the
     * spatial_samples_out role in the same iteration (row stride peer_*_ld); local index w is mirrored at peer index
     * w + peer_*_shift when peer_*_first <= w < peer_*_first + peer_*_count (the owned voxels that lie in that
     * neighbour's halo; they must be the first / last owned voxels: the kernel processes them FIRST so that the
     * stores are under way while the interior computes).  Visibility is ordered by the all-reduce of ak_grad that
     * ends every iteration on every rank (svbasl_hyper: flags stored behind a system-scope fence).  NULL = none. */
    float *peer_lo, *peer_hi;
    int64_t peer_lo_ld, peer_hi_ld, peer_lo_shift, peer_hi_shift;
    int64_t peer_lo_first, peer_lo_count, peer_hi_first, peer_hi_count;
} svbasl_engine;

/* TensorFlow-form Adam (tf.train.AdamOptimizer): m,v [n_state][ld]; lr_t[step] precomputed on the
 * host: lr*sqrt(1-b2^t)/(1-b1^t), t = step+1 */
typedef struct svbasl_adam {
    float *m, *v;
    const float *lr_t;             /* device [>= step0 + n_iters] */
    float beta1, beta2, epsilon;
    int64_t step0;                 /* global index of the first iteration of this launch */
    int32_t n_iters;               /* iterations fused into this launch (>1 only without spatial priors): the state
                                    * and the batch stay in registers, m / v in shared memory between them and are
                                    * written back once; results are bit-identical to n_iters launches of one */
    int32_t n_batches;             /* time-point mini-batches per epoch: row0 = (step % n_batches) */
} svbasl_adam;

/* Tail of a spatial iteration fused into the step kernel (svbasl_step_spatial): the LAST CTA of the launch to
 * finish (device counter `done_ctas`) all-reduces d(cost)/d(log ak) over the ranks' peer-memory mailboxes, applies
 * the TF-Adam step to log ak, zeroes ak_grad and advances *step_dev - what svbasl_hyper_step_peers does as a
 * separate launch.  With world == 1 no mailbox is touched.  The wait on the other ranks is bounded (~10 s); on
 * expiry *status is set to 1 and later launches do not wait (the results are then invalid - poll status). */
#define SVBASL_MAX_PEERS 16
typedef struct svbasl_hyper {
    float *log_ak;                 /* [n_spatial], the buffer engine->log_ak points to */
    float *m, *v;                  /* [n_spatial] Adam moments of log ak */
    const float *lr_t;             /* device table, indexed by the iteration */
    long long *step_dev;           /* the counter engine->step_dev points to (required) */
    unsigned int *done_ctas;       /* device counter, zero before the first launch; reset by the kernel */
    float beta1, beta2, epsilon;
    int32_t n_spatial;
    int32_t rank, world;
    void *mailboxes[SVBASL_MAX_PEERS]; /* entry r = rank r's mailbox (svbasl_mailbox_bytes(world) bytes, zeroed) */
    int32_t *status;               /* device int32 */
} svbasl_hyper;

const char *svbasl_last_error(void);
int svbasl_abi_version(void);
/* number of model parameters P for a model descriptor (len(model.params), aslrest.py:183-246) or <0 */
int svbasl_model_n_params(const svbasl_model *model);
/* rows of the state array for (model, engine priors) or <0 */
int svbasl_n_state(const svbasl_model *model, const svbasl_engine *engine);

/* Model.evaluate / ievaluate (aslrest.py:248-340, aslrest_disp.py:48-67, aslnn.py:93-126).
 * params [P][n_rows] (row = voxel*S + sample, as the reference's [P,W,S,1]),
 * tpts [n_t_rows][B] row-major with n_t_rows == n_rows/rows_per_t (W or 1: the reference's [W,1,B] / [1,1,B]),
 * out [n_rows][B] row-major (the reference's [W,S,B]).  pv arrays in `model` are indexed by voxel = row / S. */
int svbasl_evaluate(const svbasl_model *model, const float *params, const float *tpts, float *out,
                    int64_t n_rows, int32_t n_samples, int32_t n_batch, int64_t n_t_rows, void *stream);

/* aslnn on the tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, weight tile fed by a TMA bulk copy):
 * the 10x10 hidden layer of AslNNModel.evaluate (aslnn.py:238-260) as a batched GEMM over tiles of 128 rows, with a
 * hi/lo operand split that keeps float32-level accuracy (csrc/nn_tc.cu).
 *   svbasl_nn_pack_weights: HOST helper - writes the 512-float B-operand tile for model->nn_weights.
 *   svbasl_nn_evaluate_tc:  b_tile = that tile in DEVICE memory; params [2][n_rows] (ftiss, delttiss), tpts
 *   [n_t_rows][B], out [n_rows][B]; hidden (optional, may be NULL) [n_rows*B][10] receives the second-layer
 *   pre-activations; status (device int32, may be NULL) is set non-zero if a bounded wait on the tensor core
 *   expired (results are then invalid).  Same shapes and result as svbasl_evaluate for SVBASL_MODEL_ASLNN. */
#define SVBASL_NN_BTILE_FLOATS 512
int svbasl_nn_pack_weights(const svbasl_model *model, float *host_tile);
int svbasl_nn_evaluate_tc(const svbasl_model *model, const float *b_tile, const float *params, const float *tpts,
                          float *out, float *hidden, int64_t n_rows, int32_t n_batch, int64_t n_t_rows,
                          int32_t *status, void *stream);

/* One evaluation of the per-voxel cost (negative free energy) and its gradient with respect to the
 * posterior state - the body of svb's sess.run(cost/gradients) for this plugin family.
 * cost [ld] per-voxel cost (may be NULL); grad [n_state][ld] = grad_scale * d(sum cost)/d(state);
 * cost_sum device double[1] accumulates sum of per-voxel cost (may be NULL). */
int svbasl_elbo_grad(const svbasl_model *model, const svbasl_engine *engine, int64_t step,
                     float *cost, float *grad, double *cost_sum, void *stream);

/* Fused iteration(s): ELBO + gradient + Adam update of the posterior state in place - the body of
 * svb's sess.run(optimize).  cost_sum device double[adam->n_iters] (one per fused iteration, may be NULL);
 * nan_count device int64[1] counts voxels whose update was skipped for non-finite gradients (may be NULL). */
int svbasl_step(const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam,
                double *cost_sum, long long *nan_count, void *stream);

/* svbasl_step for spatial ("M") priors with the hyper-parameter tail fused in (svbasl_hyper above): ONE launch per
 * iteration does sampling, model, likelihood, latent loss incl. the MRF term, gradients, Adam, the next iteration's
 * neighbour samples (engine->spatial_samples_out, mirrored to the adjacent ranks through engine->peer_lo/hi), the
 * all-reduce of d(cost)/d(log ak), its Adam step and the advance of the iteration counter.  hyper == NULL behaves
 * like svbasl_step (the caller then reduces ak_grad and calls svbasl_hyper_step*). */
int svbasl_step_spatial(const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam,
                        const svbasl_hyper *hyper, double *cost_sum, long long *nan_count, void *stream);

/* Pre-pass of a step with spatial priors: out [n spatial params][S][ld] = theta_{p,s} = mu_p + (L eps_s)_p for
 * every local voxel in [0, n_local) (owned + halo), from engine->state and the step's draws (engine->eps or the
 * Philox stream of `step`; with engine->step_dev the iteration is *step_dev + step).  Neighbours then read each
 * other's samples instead of rebuilding them. */
int svbasl_sample_spatial(const svbasl_engine *engine, int64_t n_local, int64_t step, float *out, void *stream);

/* The same for the NEXT iteration, right behind a step: the owned voxels [w_begin, w_begin + n_vox) only, from the
 * state the step has just written, for iteration (*engine->step_dev or 0) + step; a shard-boundary voxel's samples
 * are also stored into the adjacent ranks' halo columns (engine->peer_lo / peer_hi name THEIR `out` buffers), so the
 * halo voxels never need a pre-pass or an exchange launch.  Boundary voxels are processed by the first CTAs. */
int svbasl_sample_spatial_next(const svbasl_engine *engine, int64_t step, float *out, void *stream);

/* Adam update of the global spatial-precision hyper-parameters from ak_grad (after any allreduce). */
int svbasl_hyper_step(float *log_ak, float *m, float *v, const double *ak_grad, int32_t n, float grad_scale,
                      float lr_t, float beta1, float beta2, float epsilon, void *stream);

/* Graph-friendly tail of a spatial iteration: Adam on log ak with lr_t[*step_dev] from the device table, then
 * ak_grad is zeroed for the next iteration and *step_dev += 1. */
int svbasl_hyper_step_dev(float *log_ak, float *m, float *v, double *ak_grad, int32_t n, float grad_scale,
                          const float *lr_t, long long *step_dev, float beta1, float beta2, float epsilon, void *stream);
/* The same tail for a spatial prior sharded over the GPUs of one box, with the all-reduce of ak_grad fused in and
 * done over NVLink peer memory instead of a separate NCCL launch (replaces svb's session-wide reduction of the
 * MRFSpatialPrior gradient + svbasl_hyper_step_dev): every rank stores its partial sums and a sequence number
 * (= *step_dev + 1) into ITS slot of every rank's mailbox, waits until all `world` slots of its own mailbox carry
 * that number, adds them in rank order (bit-identical on every rank) and applies the Adam step.  Because every
 * rank's stores are ordered behind its step kernels, the wait is also the barrier that makes the neighbours'
 * mirrored boundary state (svbasl_engine.peer_lo / peer_hi) visible before the next iteration.
 * mailboxes: HOST array of `world` device pointers, entry r = rank r's mailbox (own: the local allocation, others:
 * svbasl_shared_open of their handle), each svbasl_mailbox_bytes(world) bytes, zero-initialised.
 * status: device int32, set to 1 if a peer did not arrive within ~10 s (the wait is bounded so that a lost rank
 * cannot hang the box); once set, later calls do not wait.  world <= SVBASL_MAX_PEERS. */
int64_t svbasl_mailbox_bytes(int32_t world);
int svbasl_hyper_step_peers(float *log_ak, float *m, float *v, double *ak_grad, int32_t n, float grad_scale,
                            const float *lr_t, long long *step_dev, float beta1, float beta2, float epsilon,
                            int32_t rank, int32_t world, void *const *mailboxes, int32_t *status, void *stream);
/* *step_dev += inc (tail of a captured iteration without spatial priors). */
int svbasl_advance_step(long long *step_dev, long long inc, void *stream);

/* Device buffers that an adjacent rank (another process, another GPU of the box) can store into: plain cudaMalloc
 * memory exported as a CUDA IPC handle (64 bytes).  The neighbour opens the handle WITH ITS OWN DEVICE CURRENT
 * (cudaIpcMemLazyEnablePeerAccess), which is what makes the mapping usable by kernels running on that device -
 * the pointer it gets goes into svbasl_engine.peer_lo / peer_hi.  alloc/free and open/close pair up. */
#define SVBASL_IPC_HANDLE_BYTES 64
int svbasl_shared_alloc(int64_t bytes, void **dev_ptr, unsigned char handle[SVBASL_IPC_HANDLE_BYTES]);
int svbasl_shared_free(void *dev_ptr);
int svbasl_shared_open(const unsigned char handle[SVBASL_IPC_HANDLE_BYTES], void **dev_ptr);
int svbasl_shared_close(void *dev_ptr);

/* Write the Philox stream the fused kernels consume: eps [P'][S][ld] for `step`. */
int svbasl_fill_eps(float *eps, int64_t n_vox, int64_t ld, int64_t vox_offset, int32_t n_par, int32_t n_samples,
                    uint64_t seed, int64_t step, void *stream);

/* Posterior initialisers on device (aslrest.py:461-520, svb noise init): data [T][ld] ->
 * mean_t (>=floor), max_t (>=floor), variance_t (>=1) per voxel, time of max. Any output may be NULL. */
int svbasl_init_stats(const float *data, const float *tpts, int64_t n_vox, int64_t ld, int32_t t_full,
                      float *mean_t, float *max_t, float *var_t, float *t_at_max, void *stream);

/* Model fit (mean prediction at posterior mean) for save_model_fit: out [T][ld]. */
int svbasl_model_fit(const svbasl_model *model, const svbasl_engine *engine, float *out, void *stream);

/* End-to-end iteration with HOST buffers, as svb feeds each batch through feed_dict: copies the
 * batch rows of the data (and their time points) from pinned host memory, runs svbasl_step, copies the summed
 * cost back.  Double-buffered on two internal streams; call svbasl_host_sync() before reading costs.
 * Time points: either host_tpts [B][ld] (one value per voxel and batch row), or - when host_tpts is NULL - the
 * low-rank form the model defines them by (aslrest.py:438-440): host_ti [B] (the batch's TIs, copied each step)
 * plus engine->zoff (per-voxel slice offset z*slicedt, resident on the device, may be NULL), which halves the
 * bytes that cross PCIe per step. */
typedef struct svbasl_host_ctx svbasl_host_ctx;
int svbasl_host_ctx_create(svbasl_host_ctx **ctx, int64_t ld, int32_t n_batch);
int svbasl_host_ctx_destroy(svbasl_host_ctx *ctx);
int svbasl_step_host(svbasl_host_ctx *ctx, const svbasl_model *model, const svbasl_engine *engine,
                     const svbasl_adam *adam, const svbasl_hyper *hyper /* spatial priors: fused tail, else NULL */,
                     const float *host_data /*[B][ld]*/, const float *host_tpts /*[B][ld] or NULL*/,
                     const float *host_ti /*[B], used when host_tpts == NULL*/, double *host_cost_sum /* pinned, [1] */);
int svbasl_host_sync(svbasl_host_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SVBASL_H */

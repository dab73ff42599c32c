/* fmb200.h -- C ABI of the B200-native libFM SGD hot path.
 *
 * The reference (srendle/libfm) has no plugin/FFI interface; its de-facto seam is
 * the fm_learn vtable (src/libfm/src/fm_learn.h:31-60) and the per-row calls
 * fm->predict / fm_SGD (src/libfm/src/fm_learn_sgd_element.h:57,66).  A per-row
 * seam is useless for a GPU, so this ABI replaces the BODY OF THE EPOCH LOOP
 * (fm_learn_sgd_element.h:56-67) and the evaluate / predict passes
 * (fm_learn.h:93-153, fm_learn_sgd.h:76-90) at per-epoch granularity.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message is
 *    available from fmb200_last_error() (thread-local static string).  Nothing
 *    throws across the boundary (the reference throws std::string / const char*,
 *    caught at libfm.cpp:436-440; the host learner converts codes back to that).
 *  - plain pointers and sizes only.  Host pointers unless the name says device.
 *  - one context == one GPU.  Not re-entrant per context (same as the reference,
 *    whose fm_model carries mutable scratch, fm_model.h:65); distinct contexts
 *    may be driven from distinct threads/processes.
 *  - the library is CUDA-only: there is no CPU fallback.  fmb200_create fails
 *    loudly when no sm_100 device is usable.
 */
#ifndef FMB200_H_
#define FMB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmb200_ctx fmb200_ctx;

#define FMB200_TASK_REGRESSION 0     /* fm_learn.h:47 */
#define FMB200_TASK_CLASSIFICATION 1 /* fm_learn.h:48 */

/* execution modes of fmb200_sgd_epoch */
#define FMB200_MODE_INORDER 0 /* sequential-equivalent: rows strictly in file order, fp64
                                 state; bit-compatible with fm_learn_sgd_element::learn */
#define FMB200_MODE_HOGWILD 1 /* throughput: rows in parallel, fp32 state, red.global.add
                                 write-back, damped per-tile bias step */

#define FMB200_MODE_ORDERED 2 /* sequentially consistent: every example reads all parameters as
                                 the examples before it left them (the reference's order), fp64
                                 state; conflict-free runs of rows execute in parallel and the bias
                                 chain is solved by an affine prefix scan, so sums associate
                                 differently: deterministic, within ~1e-12 of the reference (not
                                 bit-exact; the <=1e-5 RMSE gate with margin), and fast */

#define FMB200_MAX_SLOTS 8
#define FMB200_MAX_PEERS 16
#define FMB200_IPC_HANDLE_BYTES 64

/* Replaces: fm_model construction, libfm.cpp:245-256 (num_attribute, k0, k1, num_factor).
 * `device` is the CUDA ordinal. */
int fmb200_create(fmb200_ctx** out, int device, uint32_t n_attr, int num_factor, int use_w0,
                  int use_w);
void fmb200_destroy(fmb200_ctx* ctx);
const char* fmb200_last_error(void);

/* Replaces: the learner fields set at libfm.cpp:294-309,366-404:
 * task, learn_rate (scalar; fm_learn_sgd.h:67-69), reg0/regw/regv, min/max_target. */
int fmb200_set_hparams(fmb200_ctx* ctx, int task, double learn_rate, double reg0, double regw,
                       double regv, double min_target, double max_target);
int fmb200_set_mode(fmb200_ctx* ctx, int mode);

/* Replaces: the in-memory matrix Data::load builds (Data.h:180-290): upload a CSR
 * copy of a data set into `slot` (0 = train, 1 = test, ...).  The context copies;
 * the caller may free afterwards.  row_ptr has n_rows+1 entries, row_ptr[0]==0.
 * Fails if any col >= n_attr (the reference's live assert, fm_model.h:112). */
int fmb200_upload_data(fmb200_ctx* ctx, int slot, uint64_t n_rows, uint64_t nnz,
                       const uint64_t* row_ptr, const uint32_t* col, const float* val,
                       const float* target);
/* Asynchronous variant: enqueues the copy (and the device-side validation) on the
 * context's copy stream and returns.  The slot must not be in use by a running epoch.
 * The next call that touches the slot (epoch / evaluate / predict) waits for the copy
 * and reports a validation failure; this lets the upload of the NEXT batch overlap the
 * epoch on the current one (two slots, ping-pong).  Host buffers must stay valid and
 * should be page-locked (fmb200_host_alloc) for the copy to be truly asynchronous. */
int fmb200_upload_data_async(fmb200_ctx* ctx, int slot, uint64_t n_rows, uint64_t nnz,
                             const uint64_t* row_ptr, const uint32_t* col, const float* val,
                             const float* target);
/* Same, straight from the reference's AoS layout (util/fmatrix.h:34-42):
 * `rows` points at n_rows sparse_row{sparse_entry* data; uint size;} records (16 B
 * each on LP64), each entry {uint id; float value} (8 B).  When the rows lie back to back in one
 * block -- as Data::load allocates them (Data.h:238,260) -- the row array and the block are
 * copied as they are and converted on the device (offset scan + AoS->SoA split); rows scattered
 * over the heap are gathered on the host first. */
int fmb200_upload_data_aos(fmb200_ctx* ctx, int slot, uint64_t n_rows, const void* rows,
                           const float* target);
/* One-hot rows of a fixed width (every value 1.0, e.g. (user, item) pairs -- what Data::load
 * produces for `y u:1 i:1` files): only ids[n_rows * nnz_per_row] and the targets cross PCIe
 * (4*z + 4 bytes per row instead of 12*z + 12); row offsets and values are materialised on the
 * device.  The _async form behaves like fmb200_upload_data_async. */
int fmb200_upload_onehot(fmb200_ctx* ctx, int slot, uint64_t n_rows, uint32_t nnz_per_row,
                         const uint32_t* ids, const float* target);
int fmb200_upload_onehot_async(fmb200_ctx* ctx, int slot, uint64_t n_rows, uint32_t nnz_per_row,
                               const uint32_t* ids, const float* target);
int fmb200_free_data(fmb200_ctx* ctx, int slot);

/* Page-locked host memory for the arrays handed to fmb200_upload_data: uploads from it
 * run at full PCIe rate and asynchronously (pageable memory is staged by the driver).
 * The reference allocates its CSR with plain new[] (Data.h:238); a loader that wants
 * the fast path allocates here instead.  Any host memory is accepted by the upload. */
int fmb200_host_alloc(void** out, uint64_t bytes);
int fmb200_host_free(void* p);

/* Replaces: reading / writing fm_model::w0, w, v (fm_model.h:46-48).  v is the
 * reference's FACTOR-MAJOR double [num_factor][n_attr] (util/matrix.h:152-175). */
int fmb200_set_params(fmb200_ctx* ctx, double w0, const double* w, const double* v_factor_major);
int fmb200_get_params(fmb200_ctx* ctx, double* w0, double* w, double* v_factor_major);

/* Replaces: the row loop of fm_learn_sgd_element::learn (fm_learn_sgd_element.h:56-67):
 * predict + loss multiplier + fm_SGD over every row of `slot`.  Blocking; if
 * device_seconds != NULL it receives the CUDA-event time of the epoch. */
int fmb200_sgd_epoch(fmb200_ctx* ctx, int slot, double* device_seconds);
/* enqueue only (no host sync); pair with fmb200_sync */
int fmb200_sgd_epoch_async(fmb200_ctx* ctx, int slot);
int fmb200_sync(fmb200_ctx* ctx);

/* Replaces: fm_learn::evaluate_regression / evaluate_classification
 * (fm_learn.h:113-153).  Regression fills sum_sq_err and sum_abs_err of
 * clamp(p)-y; classification fills n_correct (sign agreement). */
int fmb200_evaluate(fmb200_ctx* ctx, int slot, double* sum_sq_err, double* sum_abs_err,
                    uint64_t* n_correct);

/* Replaces: fm_learn_sgd::predict (fm_learn_sgd.h:76-90).  transform=1 applies
 * the task transform (clamp / sigmoid) exactly as -out writes it; transform=0
 * returns the raw score of fm_model::predict (fm_model.h:105-127). */
int fmb200_predict(fmb200_ctx* ctx, int slot, int transform, double* out);

/* Replaces: fm_learn_mcmc::predict_data_and_write_to_eterms (fm_learn_mcmc.h:148-378, data sets
 * without relations) -- the full re-prediction of a data set the MCMC / ALS learner runs on train
 * and test once per iteration (fm_learn_mcmc_simultaneous.h:69,122).  e_out[c] receives the e-term
 * of case c, accumulated in the learner's own order (feature-major through the transposed data),
 * bit-identical to the reference; the caller subtracts the targets and keeps the Gibbs draws.
 * Uses the fp64 state (INORDER / ORDERED mode): fmb200_set_params after every draw_all(). */
int fmb200_mcmc_eterms(fmb200_ctx* ctx, int slot, double* e_out);

/* Replaces: fm_learn_sgd_element_adapt_reg (SGDA, fm_learn_sgd_element_adapt_reg.h).
 *  _begin: init() + the prologue of learn() (:60-90, :281-292): stored gradients and the per-group
 *          regularisation values start at 0, fm->w is zeroed; attr_group[n] = DataMetaInfo::attr_group
 *          (NULL = one group).
 *  _epoch: one pass of :295-311 -- a theta-step (:136-169) per training row, each followed, when
 *          lambda_steps != 0 (the reference skips them in its first epoch, :301), by a lambda-step
 *          (:201-248) on the next validation row, the cursor restarting per epoch and wrapping.
 *          Sequential semantics, fp64, bit-identical to the reference (one warp: a parity path).
 *  _get_reg: reg_w[n_groups], reg_v[n_groups][num_factor]. */
int fmb200_sgda_begin(fmb200_ctx* ctx, uint32_t n_groups, const uint32_t* attr_group);
int fmb200_sgda_epoch(fmb200_ctx* ctx, int train_slot, int val_slot, int lambda_steps, double* device_seconds);
int fmb200_sgda_get_reg(fmb200_ctx* ctx, double* reg_w, double* reg_v);

/* Multi-GPU plumbing (row sharding + one all-reduce of w0|w|V per epoch; the
 * reference has no equivalent).  The HOGWILD state is one packed fp32 device
 * buffer [w0, pad x3 | w (strided) | V[n][kp]]; the caller all-reduces it
 * (NCCL) and calls fmb200_scale_params(1/G).  Both run on fmb200_stream(). */
int fmb200_params_device(fmb200_ctx* ctx, void** device_ptr, uint64_t* n_floats);
int fmb200_scale_params(fmb200_ctx* ctx, double factor);
/* geometry of the packed fp32 state: w0 at [0], w[i] at [off_w + i*ws], V[i][f] at [off_v + i*kp + f] */
int fmb200_params_layout(fmb200_ctx* ctx, uint64_t* off_w, int* ws, uint64_t* off_v, int* kp);
int fmb200_stream(fmb200_ctx* ctx, void** cuda_stream);

/* The same exchange without NCCL, over NVLink peer memory: every rank maps the
 * peers' state (CUDA IPC between processes: export -> exchange the 64-byte handles by
 * any means -> attach; or attach_local for contexts of one process) and
 * fmb200_allreduce_mean() launches ONE kernel per rank that barriers through peer
 * flags and averages all replicas into a second local buffer (double-buffered, so
 * fmb200_params_device() changes after every call).  Attach once, before training. */
int fmb200_peer_export(fmb200_ctx* ctx, void* handle /* FMB200_IPC_HANDLE_BYTES */);
int fmb200_peer_attach_ipc(fmb200_ctx* ctx, int world, int rank, const void* handles /* world x 64 B */);
int fmb200_peer_attach_local(fmb200_ctx* ctx, int world, int rank, fmb200_ctx* const* contexts);
int fmb200_allreduce_mean(fmb200_ctx* ctx);
/* Same exchange, but instead of the plain mean the replicas' epoch steps are combined as
 * theta = theta0 + gamma_i * sum_g (theta_g - theta0), gamma_i = (1-(1-s_i)^G)/(G s_i) with s_i the
 * relative size of one shard-epoch's step on parameter i (from the per-feature counts every upload
 * builds): parameters a shard-epoch already converges (the bias, hot features) are averaged, barely
 * touched ones are summed.  8 shards then follow the single-stream trajectory instead of advancing
 * 1/8 epoch per epoch.  Call after every epoch, like fmb200_allreduce_mean.
 * State of 8 MB and more takes the SLICED form: rank r combines slice r (read from all replicas) and pushes
 * it into every rank's buffers -- 2(G-1)/G x state bytes over NVLink per rank instead of (G-1) x -- followed
 * by a peer barrier (fmb200_set_tuning variant 8 forces it, 9 forces the one-shot kernel). */
int fmb200_allreduce_meanfield(fmb200_ctx* ctx);
/* stream-ordered barrier across the attached peers (no data); used to align ranks */
int fmb200_peer_barrier(fmb200_ctx* ctx);

/* Introspection for tests / bench */
/* the device CSR of a slot, copied back (any pointer may be NULL): lets the tests check the
 * layout conversions of the upload paths bit for bit */
int fmb200_download_data(fmb200_ctx* ctx, int slot, uint64_t* n_rows, uint64_t* nnz, uint64_t* row_ptr,
                         uint32_t* col, float* val, float* target);
int fmb200_kernel_launches(fmb200_ctx* ctx, uint64_t* count); /* kernels launched so far */
int fmb200_last_epoch_config(fmb200_ctx* ctx, int* lanes_per_row, int* slots, int* rows_per_tile,
                             int* grid, int* block, int* smem_bytes, int* damp);
/* hogwild tuning knobs; 0 keeps the default.  ctas_per_sm bounds the number of
 * rows in flight (the Hogwild staleness window).  damp: 0 = automatic hot-feature
 * damping (on when the hottest feature's expected concurrency matters), 1 = force
 * on, -1 = force off (plain summed Hogwild on w/V).  variant: 0 = automatic choice
 * of the epoch kernel, 1 = sub-warp row-group kernel, 2 = one-lane-per-row kernel
 * (k <= 8, rows of <= 4 entries; ignored when not applicable), 3 = its warp-specialised
 * form (producer warp + mbarrier hand-offs; bias read three tiles ahead).
 * INORDER mode runs the wavefront schedule of the sequential epoch for k <= 8 and rows of <= 4
 * entries (conflict-free runs of examples gather and scatter in parallel, only the bias chain
 * stays serial; bit-identical to the row-at-a-time kernel, verified on the device); variant 1
 * forces the row-at-a-time kernel. */
int fmb200_set_tuning(fmb200_ctx* ctx, int ctas_per_sm, int rows_per_tile, int threads, int damp,
                      int variant);
/* The dependency index ORDERED mode builds per data set (bit-exact index work, tested against a
 * host restatement): link[e] = e - (previous entry naming the same feature), rowdep[r] = r -
 * (nearest earlier row sharing a feature; 0 = the row names a feature twice); 0xffffffff = none.
 * Builds the index if the slot does not have it yet.  Either pointer may be NULL. */
int fmb200_ordered_index(fmb200_ctx* ctx, int slot, uint32_t* link /* [nnz] */,
                         uint32_t* rowdep /* [n_rows] */);

#ifdef __cplusplus
}
#endif
#endif /* FMB200_H_ */

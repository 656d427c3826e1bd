/*
 * b200ols.h — C ABI of the B200-native batched least-squares engine (libb200ols.so).
 *
 * This is the drop-in boundary for the ONE hot path of azmyrajab/polars_ols: the six plugin
 * expressions  least_squares{,_coefficients}, recursive_least_squares{,_coefficients},
 * rolling_least_squares{,_coefficients}  (reference: src/expressions.rs:390-446, :593-701, exported by
 * `#[polars_expr]` as `_polars_plugin_<fn>(SeriesExport* inputs, size_t n, const u8* kwargs, size_t len,
 * SeriesExport* ret)`), *batched over the groups of a `.over()` / `group_by` context* so that one call
 * carries every group instead of one FFI call per group (SURVEY.md §7 hard part 3).
 *
 * Plain C: pointers, sizes, POD structs.  No C++/torch types.  Thread-safe per context; errors are
 * returned as negative codes, the message is thread-local (`b200ols_last_error`), nothing unwinds
 * across the boundary (the reference's `catch_unwind` + `_polars_plugin_get_last_error_message`).
 *
 * Data model = the Arrow C Data Interface view the reference receives inside `SeriesExport`:
 * one contiguous values buffer per column (the reference rechunks, src/expressions.rs:40,55,88) plus
 * an optional validity bitmap (bit i of byte i/8, LSB first, 1 = valid).  NaN is a value, not a null.
 * Columns may live in host memory (the engine stages them through pinned buffers) or already in
 * device memory (zero-copy; torch / cudf / cupy pointers).
 */
#ifndef B200OLS_H
#define B200OLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200OLS_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define B200OLS_API __attribute__((visibility("default")))
#else
#define B200OLS_API
#endif

/* ---- enums (values are part of the ABI) --------------------------------------------------------- */

/* element type of the values buffers; the reference casts every input to f64 before any arithmetic
 * (src/expressions.rs:33,47,80) and always returns Float64 — so do we (outputs are f64). */
enum { B200OLS_F64 = 0, B200OLS_F32 = 1 };

enum { B200OLS_HOST = 0, B200OLS_DEVICE = 1 };

/* OutputMode, polars_ols/least_squares.py:57 ("statistics" has its own entry point, b200ols_least_squares_statistics) */
enum { B200OLS_PREDICTIONS = 0, B200OLS_RESIDUALS = 1, B200OLS_COEFFICIENTS = 2 };

/* NullPolicy, src/least_squares.rs:68-91 */
enum {
    B200OLS_NULL_IGNORE = 0,
    B200OLS_NULL_ZERO = 1,
    B200OLS_NULL_DROP = 2,
    B200OLS_NULL_DROP_ZERO = 3,
    B200OLS_NULL_DROP_Y_ZERO_X = 4,
    B200OLS_NULL_DROP_WINDOW = 5
};

/* SolveMethod, src/least_squares.rs:42-65 (NONE = Option::None) */
enum {
    B200OLS_SOLVE_NONE = 0,
    B200OLS_SOLVE_QR = 1,
    B200OLS_SOLVE_SVD = 2,
    B200OLS_SOLVE_CHOL = 3,
    B200OLS_SOLVE_LU = 4,
    B200OLS_SOLVE_CD = 5,
    B200OLS_SOLVE_CD_ACTIVE_SET = 6
};

/* return codes */
enum {
    B200OLS_OK = 0,
    B200OLS_ERR_INVALID = -1,     /* bad argument (the reference's assert!/expect panics) */
    B200OLS_ERR_UNSUPPORTED = -2, /* valid in the reference, not implemented on the device (static k > 4096; rls / rolling k > 64) */
    B200OLS_ERR_CUDA = -3,        /* CUDA runtime failure (message has the cudaError string) */
    B200OLS_ERR_NO_DEVICE = -4    /* no CUDA device: there is NO CPU fallback */
};

/* ---- data descriptors --------------------------------------------------------------------------- */

typedef struct b200ols_column {
    const void *values;      /* n_rows elements of `dtype`; 16-byte aligned when in device memory */
    const uint8_t *validity; /* Arrow validity bitmap (offset 0) or NULL = no nulls */
} b200ols_column;

/* One `.over()` / `group_by` evaluation: every input Series of the plugin call, for all groups.
 * inputs[0] of the reference = target, inputs[1..] = features (src/expressions.rs:391,431).
 * Groups are polars' GroupsProxy in CSR form:
 *   group_offsets[g] .. group_offsets[g+1]   = the rows of group g in PACKED order,
 *   row_index[p]                             = original row of packed position p (GroupsIdx), or
 *   row_index == NULL                        = groups are contiguous row slices (GroupsSlice).
 * Row order inside a group is the frame's order (it matters for rls / rolling). */
typedef struct b200ols_frame {
    int64_t n_rows;
    int32_t n_features; /* data feature columns (the intercept is NOT included here) */
    int32_t dtype;      /* B200OLS_F64 | B200OLS_F32, shared by target, features and weights */
    int32_t memspace;   /* B200OLS_HOST | B200OLS_DEVICE, shared by every pointer in this struct */
    int32_t add_intercept; /* append `const` = 1.0 AFTER all features (polars_ols/least_squares.py:184-188) */
    b200ols_column target;
    const b200ols_column *features;       /* [n_features] (host array of descriptors) */
    const b200ols_column *sample_weights; /* NULL = unweighted; else WLS: sqrt(w) scaling of target and
                                             features, null w -> sqrt(w)=1e-12, predictions un-scaled by
                                             1/sqrt(w) (polars_ols/least_squares.py:190-196,234-235) */
    int64_t n_groups;             /* >= 1 */
    const int64_t *group_offsets; /* [n_groups+1], ALWAYS a host pointer (plan metadata, like this struct);
                                     NULL => one group spanning all rows */
    const int64_t *row_index;     /* [n_rows] or NULL */
    int32_t row_index_on_device;  /* 1: row_index is a DEVICE pointer although memspace == HOST (the permutation
                                     b200ols_group_plan_build left on the device: it never visits the host) */
    int32_t _reserved;
} b200ols_frame;

/* serde kwargs structs of src/expressions.rs:298-330 as POD; NaN / negative encode Option::None */
typedef struct b200ols_ols_kwargs {
    double alpha;         /* NaN = None (-> 0.0) */
    double l1_ratio;      /* NaN = None */
    int64_t max_iter;     /* < 0 = None (-> 1000) */
    double tol;           /* NaN = None (-> 1e-5) */
    int32_t positive;     /* 0/1 */
    int32_t solve_method; /* B200OLS_SOLVE_* */
    int32_t null_policy;  /* B200OLS_NULL_* (None -> IGNORE, src/expressions.rs:340-343) */
    int32_t _reserved;
    double rcond;         /* NaN = None */
} b200ols_ols_kwargs;

typedef struct b200ols_rls_kwargs {
    double half_life;                 /* NaN = None (lambda = 1) */
    double initial_state_covariance;  /* NaN = None (-> 10.0) */
    const double *initial_state_mean; /* host [n_coef] or NULL; honoured in coefficients mode only
                                         (src/expressions.rs:604-610 vs :636) */
    int32_t null_policy;
    int32_t _reserved;
    const double *initial_information; /* host [n_groups][n_coef*n_coef + n_coef] or NULL: information state
                                         (A = P^-1 row-major, then b = A theta) ENTERING each series instead of the
                                         prior (I/p0, theta0/p0) — a series continued from an earlier time shard
                                         (SURVEY.md §8e).  Not part of the reference's RLSKwargs. */
} b200ols_rls_kwargs;

typedef struct b200ols_rolling_kwargs {
    int64_t window_size;
    int64_t min_periods;  /* < 0 = None (-> min(k, window)) */
    int32_t use_woodbury; /* < 0 = None; accepted for parity, the device solves S beta = v directly */
    int32_t null_policy;
    double alpha;         /* NaN = None (-> 0) */
} b200ols_rolling_kwargs;

/* Outputs are always f64 (reference: output_type=Float64 / struct of Float64).
 *   predictions / residuals : values[n_rows]            (original row order)
 *   static coefficients     : values[n_groups * n_coef] row-major, n_coef = n_features + add_intercept,
 *                             field order = feature order then `const` (one struct row per group,
 *                             `returns_scalar`; polars broadcasts it under .over())
 *   rls / rolling coefs     : values[n_rows * n_coef]   row-major (original row order)
 * validity (optional, may be NULL): one BYTE per output element, 1 = valid, 0 = polars null.
 * For coefficient outputs null <=> NaN (src/expressions.rs:137-139 fill_nan(NULL)).
 * memspace must equal the frame's memspace. */
typedef struct b200ols_output {
    double *values;
    uint8_t *validity;
} b200ols_output;

/* ---- context ------------------------------------------------------------------------------------- */

typedef struct b200ols_ctx b200ols_ctx; /* owns: device id, stream, pinned + device scratch */

B200OLS_API int b200ols_create(int device, b200ols_ctx **out);
/* use an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream */
B200OLS_API int b200ols_create_on_stream(int device, void *cuda_stream, b200ols_ctx **out);
B200OLS_API void b200ols_destroy(b200ols_ctx *ctx);
B200OLS_API int b200ols_synchronize(b200ols_ctx *ctx);
B200OLS_API const char *b200ols_last_error(void);
B200OLS_API int b200ols_version(void);
/* number of kernels of THIS library launched through ctx since creation (bench.py's gpu_launches) */
B200OLS_API int64_t b200ols_launch_count(const b200ols_ctx *ctx);
/* page-locked host memory for frames that are uploaded every call (full-speed, async H2D/D2H) */
B200OLS_API void *b200ols_host_alloc(size_t bytes);
B200OLS_API void b200ols_host_free(void *p);
/* Device-side timing of the dominant kernel of each call (the row-streaming Gram kernel for the static
 * models, the main pass for rls / rolling): when enabled, every such launch is bracketed by CUDA events
 * on the context's stream.  b200ols_profile_drain synchronises, writes up to `max` durations (ms, oldest
 * first) and clears the list; returns the number written (or a negative error code). */
B200OLS_API int b200ols_set_profiling(b200ols_ctx *ctx, int enabled);
B200OLS_API int b200ols_profile_drain(b200ols_ctx *ctx, float *ms, int max);
/* tuning knobs (0 = default): rows per shared-memory tile and consumer warps per CTA of the
 * row-streaming Gram kernel */
B200OLS_API int b200ols_set_tuning(b200ols_ctx *ctx, int tile_rows, int warps_per_cta, int ctas_per_sm);
/* Gram kernel variant: 3 (default) = CTA-cooperative warp-specialised TMA pipeline (producer / consumer / solver warps;
 * gram_cta / gram_multi / gram_wide / gram_pred), 1 = direct 16-byte global loads + DMMA for k <= 16, the one non-TMA
 * fallback (k > 16 always takes the TMA kernel).  `unroll` = row blocks in flight per lane of variant 1 (0 = default).
 * The environment variable B200OLS_VARIANT sets the initial variant of new contexts (test hook). */
B200OLS_API int b200ols_set_variant(b200ols_ctx *ctx, int variant, int unroll);

/* ---- group-index packing: `.over()` / group_by key columns -> CSR groups, on the device ------------------
 * Replaces what polars does in front of the reference's plugin (GroupsProxy construction + the per-group gather that
 * feeds src/expressions.rs:22-103) for the batched route: ONE call turns the key column(s) of a frame
 * into the `group_offsets` / `row_index` of b200ols_frame.  Groups ascend by key tuple (signed / unsigned integer
 * order; floats ascending with -0.0 == +0.0 and all NaN one group sorted last), rows inside a group keep the frame's
 * order — i.e. exactly `numpy.unique(keys, return_inverse=True)` followed by a stable argsort of the inverse.
 * Kernels: order-preserving radix images + varying-bit detection, LSD radix sort of (image, row) pairs over the
 * varying 8-bit digits only (stable, warp match_any ranking), boundary flags + scan -> offsets.  Already-sorted keys
 * (GroupsSlice) skip the sort and yield row_index == NULL.  n_rows must be < 2^31. */
enum { B200OLS_KEY_I64 = 0, B200OLS_KEY_I32 = 1, B200OLS_KEY_U64 = 2, B200OLS_KEY_U32 = 3, B200OLS_KEY_F64 = 4, B200OLS_KEY_F32 = 5 };

typedef struct b200ols_key_column {
    const void *values; /* n_rows elements, no nulls (a null key is a key value: encode it before the call) */
    int32_t dtype;      /* B200OLS_KEY_* */
    int32_t _reserved;
} b200ols_key_column;

/* Filled by b200ols_group_plan_build.  Every pointer is owned by the context and stays valid until the next
 * b200ols_group_plan_build on the same context (or b200ols_destroy). */
typedef struct b200ols_group_plan {
    int64_t n_rows;
    int64_t n_groups;
    const int64_t *group_offsets;   /* HOST [n_groups + 1] -> b200ols_frame.group_offsets */
    const int64_t *group_first_row; /* HOST [n_groups]: original row of the first row of each group (its key values) */
    const int64_t *row_index;       /* DEVICE [n_rows] -> b200ols_frame.row_index (+ row_index_on_device = 1 for host
                                       frames), or NULL when the groups are contiguous row slices */
    double device_ms;               /* device time of the planning kernels (CUDA events), uploads excluded */
} b200ols_group_plan;

/* keys[n_keys] (1 <= n_keys <= 8) live in `memspace`; host keys are uploaded (pageable buffers through the pinned ring). */
B200OLS_API int b200ols_group_plan_build(b200ols_ctx *ctx, const b200ols_key_column *keys, int32_t n_keys, int64_t n_rows,
                                         int32_t memspace, b200ols_group_plan *out);
/* copies of the current plan's permutation (int64 [n_rows]; identity when the groups are contiguous) and of the group id
 * of every original row (int32 [n_rows]: what `.over()` uses to broadcast per-group results) into `memspace` memory */
B200OLS_API int b200ols_group_plan_row_index(b200ols_ctx *ctx, int64_t *dst, int32_t memspace);
B200OLS_API int b200ols_group_plan_group_of_row(b200ols_ctx *ctx, int32_t *dst, int32_t memspace);

/* ---- the six entry points (one per reference plugin symbol), batched over groups ------------------ */

/* replaces _polars_plugin_least_squares (src/expressions.rs:391-428): mode PREDICTIONS | RESIDUALS */
B200OLS_API int b200ols_least_squares(b200ols_ctx *ctx, const b200ols_frame *frame, const b200ols_ols_kwargs *kwargs,
                          int mode, b200ols_output *out);
/* replaces _polars_plugin_least_squares_coefficients (src/expressions.rs:431-446) */
B200OLS_API int b200ols_least_squares_coefficients(b200ols_ctx *ctx, const b200ols_frame *frame,
                                       const b200ols_ols_kwargs *kwargs, b200ols_output *out);
/* replaces _polars_plugin_recursive_least_squares (src/expressions.rs:625-646) */
B200OLS_API int b200ols_recursive_least_squares(b200ols_ctx *ctx, const b200ols_frame *frame,
                                    const b200ols_rls_kwargs *kwargs, int mode, b200ols_output *out);
/* replaces _polars_plugin_recursive_least_squares_coefficients (src/expressions.rs:594-622) */
B200OLS_API int b200ols_recursive_least_squares_coefficients(b200ols_ctx *ctx, const b200ols_frame *frame,
                                                 const b200ols_rls_kwargs *kwargs, b200ols_output *out);
/* Time-axis sharding of ONE long series: the information-form state LEAVING each series of the
 * frame, state[g] = { A (n_coef x n_coef, symmetric, row-major), b (n_coef), D }, where
 *   (A, b)_leaving = D * (A, b)_entering + (A, b)_rows   and   D = lambda ^ (#valid rows)
 * (src/least_squares.rs:505-540 in information form).  The entering state is
 * kwargs->initial_information, else the prior.  Called with an all-zero initial_information it returns the pure
 * affine map of the shard; maps of consecutive shards compose, which is the only exchange a time-sharded rls
 * needs (one all-gather of n_coef^2 + n_coef + 1 doubles per rank).  `state` is a HOST pointer
 * [n_groups][n_coef*n_coef + n_coef + 1]; the call synchronises. */
B200OLS_API int b200ols_recursive_least_squares_state(b200ols_ctx *ctx, const b200ols_frame *frame,
                                                      const b200ols_rls_kwargs *kwargs, double *state);
/* replaces _polars_plugin_rolling_least_squares (src/expressions.rs:679-701) */
B200OLS_API int b200ols_rolling_least_squares(b200ols_ctx *ctx, const b200ols_frame *frame,
                                  const b200ols_rolling_kwargs *kwargs, int mode, b200ols_output *out);
/* replaces _polars_plugin_rolling_least_squares_coefficients (src/expressions.rs:649-676) */
B200OLS_API int b200ols_rolling_least_squares_coefficients(b200ols_ctx *ctx, const b200ols_frame *frame,
                                               const b200ols_rolling_kwargs *kwargs, b200ols_output *out);

/* ---- multi-GPU: fused coefficient gather over NVLink peer memory ------------------------------------
 * Groups are independent, so ranks shard by group and the only exchange is the final gather of the
 * per-rank coefficient chunks.  Instead of a collective after the kernel, every rank's
 * solver warps store beta straight into EVERY rank's full [total_groups, n_coef] buffer (P2P stores over
 * NVLink / NVSwitch, peer pointers obtained through CUDA IPC): the gather costs no extra launch.
 *   b200ols_device_alloc / _free : cudaMalloc'd (IPC-exportable) device memory
 *   b200ols_ipc_export / _open / _close : cudaIpc{Get,Open,Close}MemHandle (64-byte handles travel between the
 *                                  ranks by any host channel, e.g. torch.distributed.all_gather_object)
 *   b200ols_set_peer_gather      : peer_coef[r] = rank r's full buffer (r = 0..n_peers-1, own rank included);
 *                                  subsequent b200ols_least_squares_coefficients calls on DEVICE frames write
 *                                  rows [group_base, group_base + n_groups) of every peer buffer
 *                                  (out->values may then be NULL).  n_peers = 0 switches it off.
 * Ordering is the caller's: a rank may read its buffer once all ranks have synchronised their streams
 * (e.g. stream sync + barrier), exactly as after a collective. */
B200OLS_API void *b200ols_device_alloc(b200ols_ctx *ctx, size_t bytes);
B200OLS_API void b200ols_device_free(b200ols_ctx *ctx, void *p);
B200OLS_API int b200ols_device_memset(b200ols_ctx *ctx, void *dev_ptr, int value, size_t bytes); /* synchronous */
B200OLS_API int b200ols_ipc_export(b200ols_ctx *ctx, const void *dev_ptr, uint8_t handle[64]);
B200OLS_API int b200ols_ipc_open(b200ols_ctx *ctx, const uint8_t handle[64], void **dev_ptr);
B200OLS_API int b200ols_ipc_close(b200ols_ctx *ctx, void *dev_ptr);
B200OLS_API int b200ols_copy_to_host(b200ols_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
B200OLS_API int b200ols_set_peer_gather(b200ols_ctx *ctx, int n_peers, void *const *peer_coef, int64_t group_base,
                                        int64_t total_groups);
/* Per-step completion of the fused gather, without a collective: peer_flags[r] = rank r's array of 8 uint64 (zeroed,
 * in IPC-exported device memory, own rank included).  b200ols_peer_step_complete(step) enqueues one tiny kernel behind
 * the step's kernels that release-stores `step` into slot [rank] of every peer's array and acquire-spins until all
 * slots of its own array reached `step`: once it has retired, every rank's rows of this step are in this rank's
 * buffer.  `step` must increase by one per call, on every rank.  b200ols_peer_timed_out: 1 when a spin gave up
 * (a peer never signalled; ~2 s), else 0; synchronises the stream. */
B200OLS_API int b200ols_set_peer_flags(b200ols_ctx *ctx, int n_peers, void *const *peer_flags, int rank);
B200OLS_API int b200ols_peer_step_complete(b200ols_ctx *ctx, uint64_t step);
/* the two halves as separate kernels, for double-buffered gathers: signal step i right behind its kernels, wait for step
 * i behind the kernels of step i + 1 (the peers' signals have arrived by then: the exchange leaves the critical path) */
B200OLS_API int b200ols_peer_step_signal(b200ols_ctx *ctx, uint64_t step);
B200OLS_API int b200ols_peer_step_wait(b200ols_ctx *ctx, uint64_t step);
/* both in ONE launch: signal `signal_step`, then wait for `wait_step` (0 = nothing to wait for yet) */
B200OLS_API int b200ols_peer_step_signal_wait(b200ols_ctx *ctx, uint64_t signal_step, uint64_t wait_step);
/* no launch at all: arm the completion for the NEXT b200ols_least_squares_coefficients call; when that call runs the fused
 * peer-gather kernel, the kernel's last solver warp signals `signal_step` and waits for `wait_step` itself (any other route:
 * the engine enqueues the flag kernel behind the call) */
B200OLS_API int b200ols_peer_arm_step(b200ols_ctx *ctx, uint64_t signal_step, uint64_t wait_step);
B200OLS_API int b200ols_peer_timed_out(b200ols_ctx *ctx);

/* "next" row: replaces _polars_plugin_predict (src/expressions.rs:706-741).
 * coefficients: n_coef Float64 child arrays of the coefficient struct Series (one value per ROW, e.g. the
 * broadcast result of `.over()` or the per-row output of rls / rolling); features: n_coef - add_intercept
 * columns of `dtype`.  predictions[r] = sum_j feature_j[r] * coefficient_j[r] (+ coefficient_last[r] when
 * add_intercept).  Null features are zero-filled unless null_policy == IGNORE (-> NaN); DROP masks rows with
 * any null input.  out->values[n_rows] (+ optional validity bytes). */
B200OLS_API int b200ols_predict(b200ols_ctx *ctx, int64_t n_rows, int32_t n_coef, int32_t dtype, int32_t memspace,
                                const b200ols_column *coefficients, const b200ols_column *features,
                                int32_t add_intercept, int32_t null_policy, b200ols_output *out);

/* "next" row: replaces _polars_plugin_least_squares_statistics (src/expressions.rs:469-509,
 * src/statistics.rs:15-156), one struct row per group (`returns_scalar`).  Coefficients follow the ordinary dispatch of
 * kwargs; r2 / mae / mse are those of the fit rows (scaled by sqrt(w) under WLS); standard errors, t and p values come
 * from (X^T X + alpha I)^-1 by Cholesky (failure -> NaN, as the reference) with df = n - p (alpha = 0) or
 * n - trace(inverse).  kwargs->alpha must not be None (the reference unwraps it).  A group with df <= 0 fails the call
 * with the reference's assertion message.  Arrays live in the frame's memspace; no nulls (NaN stays NaN).  The
 * `feature_names` list of the reference's struct is the caller's input names (host side).  Up to 4096 coefficients
 * (above 64 through the general path, as the other static modes). */
typedef struct b200ols_statistics_output {
    double *r2, *mae, *mse;                                           /* [n_groups] */
    double *coefficients, *standard_errors, *t_values, *p_values;     /* [n_groups * n_coef] row-major */
} b200ols_statistics_output;
B200OLS_API int b200ols_least_squares_statistics(b200ols_ctx *ctx, const b200ols_frame *frame,
                                                 const b200ols_ols_kwargs *kwargs, const b200ols_statistics_output *out);

/* "next" row: replaces _polars_plugin_multi_target_least_squares (src/expressions.rs:521-591;
 * solve_multi_target src/least_squares.rs:243-260).  The reference's first input is a struct Series of n_targets
 * Float fields: here `targets[n_targets]` (same dtype / memspace as the frame; frame->target is ignored).  One
 * decomposition of the group's features serves every target: OLS minimum-norm SVD solution (alpha None / 0) or ridge-SVD
 * (alpha > 0, rcond honoured).  Nulls are ZERO-filled under every policy (fill_zero = true, :549-550,569), the row mask
 * ANDs all targets (drop_y_zero_x) or all targets and features (drop, drop_zero).  mode PREDICTIONS | RESIDUALS only
 * (polars_ols/least_squares.py:314-317); kwargs must be unconstrained OLS / ridge with solve_method None | "svd".
 * out->values[n_targets * n_rows] target-major (one contiguous Float64 child array per struct field), validity bytes
 * likewise: NaN predictions are nulls (convert_array_to_struct_series, :137-139), `drop` masks invalid rows,
 * residuals are null where the target is. */
B200OLS_API int b200ols_multi_target_least_squares(b200ols_ctx *ctx, const b200ols_frame *frame, int32_t n_targets,
                                                   const b200ols_column *targets, const b200ols_ols_kwargs *kwargs,
                                                   int mode, b200ols_output *out);

/* Per-group diagnostics of the last static call on ctx (host copy): bit 0 = Cholesky failed and the
 * LU fallback ran (src/least_squares.rs:299-316), bit 1 = group had no rows after null filtering
 * (coefficients = 0, src/expressions.rs:357-359), bit 2 = re-solved by the pivoted-QR kernel
 * (ill-conditioned Gram), bit 4 = 0 < n <= k (the reference's LAPACK min-norm path), bit 5 = solved by the
 * one-sided Jacobi SVD kernel (solve_method = "svd", or n <= k).  flags must hold n_groups ints. */
B200OLS_API int b200ols_last_group_flags(b200ols_ctx *ctx, int32_t *flags, int64_t n_groups);

#ifdef __cplusplus
}
#endif
#endif /* B200OLS_H */

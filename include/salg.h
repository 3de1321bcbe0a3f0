/*
 * salg.h — C ABI of libsalg_b200.so: the B200-native replacement for the sparse-PCA hot path
 * of SingleRust/single-algebra 0.9.2 (crate `single_algebra`).
 *
 * The reference has no FFI boundary of its own: its operator API is the Rust type surface
 *   single_algebra::dimred::pca::{SparsePCA, SparsePCABuilder, MaskedSparsePCA,
 *       MaskedSparsePCABuilder, SVDMethod, PowerIterationNormalizer}   (src/dimred/pca/mod.rs:37-62)
 *   single_algebra::sparse::MatrixSum                                   (src/sparse/mod.rs:67-102)
 *   single_algebra::{Normalize, Log1P}                                  (src/lib.rs:50-51, src/utils/mod.rs:6-17)
 * Every entry point below names the reference item (file:line) whose body it replaces; the Rust
 * facade that binds them is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; all host arrays are BORROWED for the duration of the call;
 *  - every function returns an `int` status (SALG_OK == 0); `salg_last_error()` returns a
 *    thread-local message valid until the next call on that thread; nothing throws or aborts;
 *  - `_f32` / `_f64` suffixes are the two instantiations of the reference's generic `T`;
 *  - host CSR input uses nalgebra-sparse's layout: `usize` (= uint64_t) row offsets [nrows+1] and
 *    column indices [nnz], values T [nnz], column indices strictly increasing within a row;
 *  - dense outputs are row-major with the stated shape, written into caller-allocated memory;
 *  - a `salg_ctx` owns one GPU (device, streams, workspaces, optional NCCL communicator) and is
 *    not safe for concurrent calls; one process per GPU in multi-GPU runs (rows sharded).
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    SALG_ERR_CUDA.
 *  - per-GPU size limits: row offsets are 64-bit everywhere, so statistics, Normalize / Log1P, column selection, transform
 *    the fused masked fit and — with the default TMEM-operand products, whose tile records are addressed with 64 bits —
 *    unmasked f32 randomized fits take shards of any size that fits in memory; the transposed copy (f64 fits, Lanczos,
 *    salg_csr_transpose) and the round-1 dense-tile format (SALG_SPMM_IMPL=tc) index entries with 32 bits and need
 *    < 2^31 stored entries PER GPU SHARD (SALG_ERR_UNSUPPORTED otherwise: shard the rows over more GPUs).
 */
#ifndef SALG_H
#define SALG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SALG_VERSION 100   /* 0.1.0 */

enum salg_status {
    SALG_OK = 0,
    SALG_ERR_BAD_ARG = 1,
    SALG_ERR_MASK_LEN = 2,     /* "The mask vector length and the number of features (columns) have to be the same!" (pca/sparse_masked/mod.rs:258-262, 440-444) */
    SALG_ERR_NOT_FITTED = 3,   /* "Must be fitted before transform!" (pca/sparse/mod.rs:259,263) */
    SALG_ERR_CUDA = 4,
    SALG_ERR_NCCL = 5,
    SALG_ERR_NUMERIC = 6,      /* "SVD computation failed: ..." (pca/sparse/mod.rs:144) / "Randomized SVD computation failed: ..." (:180) */
    SALG_ERR_UNSUPPORTED = 7,
    SALG_ERR_OOM = 8
};

enum salg_dtype { SALG_F32 = 0, SALG_F64 = 1 };

/* single_utilities::types::Direction as used by Normalize::normalize (src/sparse/csr.rs:1032,1046) */
enum salg_direction { SALG_ROW = 0, SALG_COLUMN = 1 };

/* SVDMethod (src/dimred/pca/mod.rs:49-62); default Lanczos (:64-68) */
enum salg_svd_method { SALG_SVD_LANCZOS = 0, SALG_SVD_RANDOM = 1 };

/* single_svdlib::randomized::PowerIterationNormalizer (re-export src/dimred/pca/mod.rs:41).
 * QR and LU both orthonormalise with CholeskyQR2 on the device (same column space);
 * NONE skips the tall-side factorisation. */
enum salg_normalizer { SALG_NORM_QR = 0, SALG_NORM_LU = 1, SALG_NORM_NONE = 2 };

/* transform semantics */
enum salg_transform_mode {
    SALG_TRANSFORM_EXACT = 0,            /* (X - 1 mu^T) V^T on the kept columns                 */
    SALG_TRANSFORM_REFERENCE_COMPAT = 1  /* what the reference's loops compute, bug for bug      */
};

typedef struct salg_ctx salg_ctx;   /* one GPU + streams + workspaces (+ NCCL communicator)      */
typedef struct salg_csr salg_csr;   /* device-resident CSR row shard, value type fixed at upload  */
typedef struct salg_pca salg_pca;   /* fitted model: components_, explained_variance_, mean_      */

/* ---- errors / lifecycle --------------------------------------------------------------------- */
const char* salg_last_error(void);
int salg_version(void);
int salg_device_count(int* out);

/* Single-GPU context on `device`. */
int salg_ctx_create(int device, salg_ctx** out);
/* Row-sharded context: this process is `rank` of `nranks`, one GPU each; `nccl_unique_id` is the
 * 128-byte id from salg_nccl_unique_id() on rank 0, distributed by the host program. */
int salg_nccl_unique_id(void* out128);
int salg_ctx_create_dist(int device, int rank, int nranks, const void* nccl_unique_id, salg_ctx** out);
int salg_ctx_destroy(salg_ctx* ctx);
int salg_ctx_sync(salg_ctx* ctx);
int salg_ctx_rank(const salg_ctx* ctx, int* rank, int* nranks);

/* CUDA-event stopwatch on the context's stream (the stream every kernel of the library is launched on):
 * start synchronises the stream and records; stop records, waits and returns the elapsed device ms. */
int salg_timer_start(salg_ctx* ctx);
int salg_timer_stop(salg_ctx* ctx, double* ms);
/* Sparse x panel product implementation for f32 operators: 2 = tcgen05.mma with the sparse operand expanded into tensor
 * memory (csrc/tm.cu, default), 0 = tcgen05.mma on tiles densified in shared memory (csrc/tc.cu, round 1),
 * 1 = CUDA-core chunk kernels (always used for f64).  Environment: SALG_SPMM_IMPL=tm | tc | chunk. */
int salg_ctx_set_spmm_impl(salg_ctx* ctx, int impl);
/* Number of kernels of this library launched on the context's stream since its creation. */
int salg_launch_count(salg_ctx* ctx, int64_t* out);

/* Per-kernel-class device timings (CUDA events on the launching stream), for bench.py's roofline.
 * Classes: see salg_prof_name(). Recording is off by default; on = 1: every class, on = 2: the two sparse-product classes
 * only (what a timed benchmark region needs for the roofline, without an event pair around every small kernel). */
int salg_prof_enable(salg_ctx* ctx, int on);
int salg_prof_reset(salg_ctx* ctx);
int salg_prof_count(void);
const char* salg_prof_name(int cls);
/* total device ms, launches, algorithmic bytes accumulated for class `cls` since the last reset */
int salg_prof_get(salg_ctx* ctx, int cls, double* ms, int64_t* launches, double* bytes);

/* ---- CSR container (replaces nalgebra_sparse::CsrMatrix<T> storage) ------------ */
/* Upload with usize -> u32 narrowing of the column indices; validates offsets monotone,
 * offsets[nrows]==nnz, indices strictly increasing per row and < ncols. */
int salg_csr_upload_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                        const uint64_t* row_offsets, const uint64_t* col_indices,
                        const float* values, salg_csr** out);
int salg_csr_upload_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                        const uint64_t* row_offsets, const uint64_t* col_indices,
                        const double* values, salg_csr** out);
/* Same with 32-bit column indices and int64 offsets (AnnData/scipy layout). */
int salg_csr_upload_i32_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                            const int64_t* row_offsets, const int32_t* col_indices,
                            const float* values, salg_csr** out);
int salg_csr_upload_i32_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                            const int64_t* row_offsets, const int32_t* col_indices,
                            const double* values, salg_csr** out);
int salg_csr_free(salg_csr* csr);
int salg_csr_dims(const salg_csr* csr, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* dtype);
/* Any of the three output pointers may be NULL. */
int salg_csr_download_f32(salg_ctx* ctx, const salg_csr* csr, uint64_t* row_offsets,
                          uint64_t* col_indices, float* values);
int salg_csr_download_f64(salg_ctx* ctx, const salg_csr* csr, uint64_t* row_offsets,
                          uint64_t* col_indices, double* values);
/* Device layout as is: int64 offsets, uint32 column indices, values of the csr's type; pointers may be NULL. */
int salg_csr_download_raw(salg_ctx* ctx, const salg_csr* csr, int64_t* row_offsets, uint32_t* col_indices,
                          void* values);
/* MaskedCSRMatrix::new(x, mask) (single-svdlib lanczos::masked; call site
 * pca/sparse_masked/mod.rs:313) as a materialised compaction: keeps columns with mask[c] != 0,
 * renumbered by rank among kept columns; bit-exact with `A[:, mask]`. */
int salg_csr_select_columns(salg_ctx* ctx, const salg_csr* csr, const uint8_t* mask,
                            int64_t mask_len, salg_csr** out);
/* Device CSR <-> CSC: a NEW handle holding the CSR of A^T — i.e. the CSC arrays of A (col_offsets,
 * row_indices ascending within a column, values) — built by a stable sort of the entries by column.  ncols(out) = nrows(in).
 * Applied to a handle from salg_csc_upload_* it returns the CSR of A.  The input keeps the transposed copy cached for its
 * own gather-form A^T products.  Per-GPU limit: < 2^31 stored entries (32-bit sort positions). */
int salg_csr_transpose(salg_ctx* ctx, const salg_csr* csr, salg_csr** out);
/* Device copy of the value array (same type, nnz + 16 entries) and its restore: lets a caller run the in-place
 * Normalize / Log1P chain repeatedly from the same raw counts without another upload.  Free the clone with salg_dev_free. */
int salg_csr_values_clone(salg_ctx* ctx, const salg_csr* csr, void** clone);
int salg_csr_values_restore(salg_ctx* ctx, salg_csr* csr, const void* clone);
int salg_dev_free(salg_ctx* ctx, void* p);
/* Device-side synthetic count-matrix generator (bench input; not a reference function). Rows
 * [row0, row0+nrows) of the matrix defined by the tables (see single-algebra_b200/synth.py). */
int salg_csr_synth(salg_ctx* ctx, int dtype, uint64_t seed, int64_t row0, int64_t nrows,
                   int64_t ncols, int32_t n_clusters, const uint8_t* base_level /*[n_clusters*ncols]*/,
                   const int32_t* sf_offset /*[16]*/, const uint32_t* cdf /*[256*40]*/,
                   salg_csr** out);

/* ---- MatrixSum (src/sparse/mod.rs:67-102) ---------------------------------------------------- */
/* sum_col (src/sparse/csr.rs:259-312) and sum_col_squared (:558-608) in ONE pass; `sumsq` may be
 * NULL. Accumulates in f64 on the device, returns in T. Under a dist ctx the result is the
 * all-reduced global column sum. */
int salg_sum_col_f32(salg_ctx* ctx, const salg_csr* csr, float* sum, float* sumsq);
int salg_sum_col_f64(salg_ctx* ctx, const salg_csr* csr, double* sum, double* sumsq);
/* sum_row (src/sparse/csr.rs:314-392) */
int salg_sum_row_f32(salg_ctx* ctx, const salg_csr* csr, float* out);
int salg_sum_row_f64(salg_ctx* ctx, const salg_csr* csr, double* out);
/* MatrixNonZero::nonzero_col (src/sparse/csr.rs:23-77) and MatrixVariance::var_col (:632-678) — SURVEY §8f-1; they fall
 * out of the same pass. `var` uses the reference's (sumsq/n - mean^2) * n/(n-1). Any pointer may be NULL. */
int salg_col_stats_f64(salg_ctx* ctx, const salg_csr* csr, double* sum, double* sumsq,
                       double* nnz_col, double* var_col);

/* MatrixNonZero::nonzero_row (src/sparse/csr.rs:79-122): stored entries per row = row_offsets[r+1] - row_offsets[r],
 * out[nrows]; MatrixNonZero::nonzero_col (:23-77): stored entries per column, out[ncols] (all-reduced under a dist ctx).
 * The reference is generic over unsigned integer T; the ABI returns u64 and the facade narrows. */
int salg_nonzero_row(salg_ctx* ctx, const salg_csr* csr, uint64_t* out);
int salg_nonzero_col(salg_ctx* ctx, const salg_csr* csr, uint64_t* out);

/* ---- Normalize / Log1P (src/utils/mod.rs:6-17) ----------------------------------------------- */
/* Normalize::normalize for CsrMatrix (src/sparse/csr.rs:1013-1068), in place on the device values:
 * scale[i] = sums[i] > 0 ? target/sums[i] : 0; v <- T(U(v)*scale[idx]) only where scale > 0.
 * `n_sums` must be >= nrows (ROW) / ncols (COLUMN). U = T here; *_u64 takes U = f64. */
int salg_normalize_f32(salg_ctx* ctx, salg_csr* csr, const float* sums, int64_t n_sums,
                       float target, int direction);
int salg_normalize_f64(salg_ctx* ctx, salg_csr* csr, const double* sums, int64_t n_sums,
                       double target, int direction);
int salg_normalize_f32_u64(salg_ctx* ctx, salg_csr* csr, const double* sums, int64_t n_sums,
                           double target, int direction);
/* ---- CSC twins -------------------------------------------------------------------------------------
 * nalgebra_sparse::CscMatrix<T> = (col_offsets, row_indices, values).  The device handle stores it as the CSR of A^T
 * (salg_csr_dims reports ncols stored rows); salg_csr_download_* / salg_csr_free / salg_log1p work on it unchanged
 * (Log1P for CscMatrix, src/sparse/csc.rs:737-746, touches values only). */
int salg_csc_upload_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* col_offsets,
                        const uint64_t* row_indices, const float* values, salg_csr** out);
int salg_csc_upload_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* col_offsets,
                        const uint64_t* row_indices, const double* values, salg_csr** out);
/* MatrixSum::sum_col / sum_col_squared for CscMatrix (src/sparse/csc.rs:157-197, 323-335): one pass, either output may be
 * NULL; length ncols. */
int salg_csc_sum_col_f32(salg_ctx* ctx, const salg_csr* csc, float* sum, float* sumsq);
int salg_csc_sum_col_f64(salg_ctx* ctx, const salg_csr* csc, double* sum, double* sumsq);
/* MatrixSum::sum_row for CscMatrix (src/sparse/csc.rs:199-220): scatter over row_indices; length nrows. */
int salg_csc_sum_row_f32(salg_ctx* ctx, const salg_csr* csc, float* out);
int salg_csc_sum_row_f64(salg_ctx* ctx, const salg_csr* csc, double* out);
/* Normalize::normalize for CscMatrix (src/sparse/csc.rs:680-735): `direction` is in terms of A (ROW indexes sums by
 * row_indices, COLUMN by column), same scale rule as the CSR version. */
int salg_csc_normalize_f32(salg_ctx* ctx, salg_csr* csc, const float* sums, int64_t n_sums, float target, int direction);
int salg_csc_normalize_f64(salg_ctx* ctx, salg_csr* csc, const double* sums, int64_t n_sums, double target, int direction);
int salg_csc_normalize_f32_u64(salg_ctx* ctx, salg_csr* csc, const double* sums, int64_t n_sums, double target,
                               int direction);

/* Log1P::log1p_normalize (src/sparse/csr.rs:1070-1079): v <- ln(fl(1 + v)). */
int salg_log1p(salg_ctx* ctx, salg_csr* csr);
/* Fused sum_row -> normalize(ROW, target) -> log1p -> sum_col/sum_col_squared in one pass over
 * the values. Output pointers may be NULL. */
int salg_preprocess_f32(salg_ctx* ctx, salg_csr* csr, float target, float* col_sum, float* col_sumsq);
int salg_preprocess_f64(salg_ctx* ctx, salg_csr* csr, double target, double* col_sum, double* col_sumsq);

/* ---- PCA (src/dimred/pca/sparse/mod.rs, src/dimred/pca/sparse_masked/mod.rs) ------------------ */
typedef struct salg_pca_params {
    int32_t n_components;        /* builder default 50          (pca/sparse/mod.rs:391)        */
    int32_t svd_method;          /* salg_svd_method; default LANCZOS (pca/mod.rs:64-68)        */
    int32_t n_oversamples;       /* SVDMethod::Random field     (pca/mod.rs:55)                */
    int32_t n_power_iterations;  /*                             (pca/mod.rs:57)                */
    int32_t normalizer;          /* salg_normalizer             (pca/mod.rs:59)                */
    int32_t center;              /* default 1                   (pca/sparse/mod.rs:398)        */
    int32_t verbose;             /* default 0                   (pca/sparse/mod.rs:399)        */
    uint32_t random_seed;        /* default 42                  (pca/sparse/mod.rs:397)        */
    double alpha;                /* stored, unused by the reference (pca/sparse/mod.rs:38)     */
    double tolerance;            /* stored, unused by the reference (pca/sparse/mod.rs:39)     */
    int32_t lanczos_max_steps;   /* 0 = library default                                        */
    int32_t keep_scores;         /* 1: also project the fit rows (fit_transform) and keep the scores */
    int32_t reserved[6];
} salg_pca_params;

int salg_pca_params_default(salg_pca_params* p);

/* SparsePCA::fit (pca/sparse/mod.rs:102-242) when mask == NULL, MaskedSparsePCA::fit
 * (pca/sparse_masked/mod.rs:255-419) otherwise (mask_len must equal ncols, else SALG_ERR_MASK_LEN).
 * `omega` (nullable): host Gaussian test matrix, n_eff x (rank + n_oversamples) row-major, the
 * "same host-generated omega" of the parity contract; NULL => drawn on the device from random_seed
 * (statistical parity only — the reference's ChaCha12 stream is not reproducible outside Rust). */
int salg_pca_fit_f32(salg_ctx* ctx, const salg_csr* x, const salg_pca_params* params,
                     const uint8_t* mask, int64_t mask_len,
                     const float* omega, int64_t omega_rows, int64_t omega_cols, salg_pca** out);
int salg_pca_fit_f64(salg_ctx* ctx, const salg_csr* x, const salg_pca_params* params,
                     const uint8_t* mask, int64_t mask_len,
                     const double* omega, int64_t omega_rows, int64_t omega_cols, salg_pca** out);
/* The same fit straight from a HOST matrix in the AnnData / scipy layout (int64 offsets, int32 column indices, f32
 * values).  A masked randomized fit STREAMS the matrix: row chunks cross PCIe once, double-buffered, and go
 * through validation + the statistics / mask-compaction pass while the next chunk uploads; only the kept entries and the
 * tile format stay on the device (a 16.8 GB matrix needs ~3 GB).  Pinned host memory gives the overlap; pageable memory
 * works but serialises.  Other configurations (no mask, Lanczos, values that are not raw counts in rows too dense for the
 * kept-entry slots) upload the matrix, fit and free it inside the call.  Replaces the upload + fit pair of
 * MaskedSparsePCA::fit (pca/sparse_masked/mod.rs:255-419) for host-resident inputs. */
int salg_pca_fit_host_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* row_offsets,
                          const int32_t* col_indices, const float* values, const salg_pca_params* params,
                          const uint8_t* mask, int64_t mask_len, const float* omega, int64_t omega_rows,
                          int64_t omega_cols, salg_pca** out);
int salg_pca_free(salg_pca* pca);
/* d = number of components returned, n_eff = kept columns, ncols = full column count */
int salg_pca_dims(const salg_pca* pca, int64_t* d, int64_t* n_eff, int64_t* ncols, int* dtype);
/* components_ : d x n_eff row-major (pca/sparse/mod.rs:208) */
int salg_pca_components_f32(const salg_pca* pca, float* out);
int salg_pca_components_f64(const salg_pca* pca, double* out);
int salg_pca_singular_values_f64(const salg_pca* pca, double* out /* d */);
/* explained_variance_[i] = s[i]^2/(n-1) (pca/sparse/mod.rs:210-216) */
int salg_pca_explained_variance_f64(const salg_pca* pca, double* out /* d */);
/* mean_ : full ncols length (pca/sparse_masked/mod.rs:280-291) */
int salg_pca_mean_f64(const salg_pca* pca, double* out /* ncols */);
int salg_pca_total_var(const salg_pca* pca, double* out);
/* Samples (rows over ALL ranks) the model was fitted on: the n of explained_variance = s^2 / (n - 1)
 * (pca/sparse/mod.rs:210-216) and of the noise-variance estimate the reference prints (:225-238). */
int salg_pca_n_samples(const salg_pca* pca, int64_t* out);
/* bit 1: a sketch column was numerically dependent on the ones before it (rank-deficient panel: rank(A) below
 *        n_components + n_oversamples, or an ill-conditioned f32 panel): its Cholesky pivot was floored and, when it was
 *        still dependent in the last CholeskyQR pass, the column was dropped — the triplets past the numerical rank then
 *        come back as zero singular values with zero component rows (the reference's Householder QR completes the basis
 *        with arbitrary orthonormal vectors instead);
 * bit 2: Jacobi sweep limit reached;
 * bit 4: Lanczos returned before every requested triplet met the acceptance bound.
 * A randomized fit whose singular values exceed the operator's Frobenius norm (sum s_i^2 > ||A_c||_F^2, known from the
 * statistics pass) returns SALG_ERR_NUMERIC "Randomized SVD computation failed: ..." instead of a model (rank-deficient
 * sketch with n_power_iterations = 0). */
int salg_pca_numeric_flags(const salg_pca* pca, int* out);
/* SparsePCA::transform (pca/sparse/mod.rs:255-285) / MaskedSparsePCA::transform
 * (pca/sparse_masked/mod.rs:438-546): scores nrows(x) x d row-major into `scores` (host). */
int salg_pca_transform_f32(salg_ctx* ctx, const salg_pca* pca, const salg_csr* x, int mode, float* scores);
int salg_pca_transform_f64(salg_ctx* ctx, const salg_pca* pca, const salg_csr* x, int mode, double* scores);
/* fit_transform (pca/sparse/mod.rs:355-358): fit, then project the same rows. With
 * params->keep_scores the fit itself finishes with the EXACT projection of the fit rows (one more
 * SpMM over the still-resident compacted operator) and these calls only copy it out. */
int salg_pca_fit_scores_f32(salg_ctx* ctx, const salg_pca* pca, float* scores);
int salg_pca_fit_scores_f64(salg_ctx* ctx, const salg_pca* pca, double* scores);
/* Device-resident variant used for the kernel-only timing: runs transform and leaves the result
 * on the device (no D2H). */
int salg_pca_transform_device(salg_ctx* ctx, const salg_pca* pca, const salg_csr* x, int mode);

/* ---- operator-level entry points (parity tests / microbenchmarks of the individual kernels) --- */
/* out = A * dense - 1 * (mu^T dense)   (transposed == 0; dense is ncols x k, out nrows x k), or
 * out = A^T * dense - mu * (1^T dense) (transposed == 1; dense is nrows x k, out ncols x k).
 * `mu` (ncols) may be NULL for the uncentred product. Host row-major arrays, k <= 64. */
int salg_op_spmm_f32(salg_ctx* ctx, const salg_csr* csr, int transposed, const float* dense,
                     int64_t k, const float* mu, float* out);
int salg_op_spmm_f64(salg_ctx* ctx, const salg_csr* csr, int transposed, const double* dense,
                     int64_t k, const double* mu, double* out);
/* Thin QR by CholeskyQR2 of a host m x k panel (k <= 64): q (m x k) and r (k x k) row-major. */
int salg_op_cholqr2_f32(salg_ctx* ctx, const float* panel, int64_t m, int64_t k, float* q, double* r);
int salg_op_cholqr2_f64(salg_ctx* ctx, const double* panel, int64_t m, int64_t k, double* q, double* r);
/* Singular values (descending) and right/left factors of a host k x k matrix by the one-CTA
 * Jacobi kernel: a = u * diag(s) * vt. */
int salg_op_small_svd(salg_ctx* ctx, const double* a, int64_t k, double* u, double* s, double* vt);
/* Test / probe hook for the fused tall-panel pass of the f32 tensor-core path (tc.cu: tc_gram_prep_kernel): Gram matrix
 * (k x k row-major) and column sums of a panel in one pass, f32 products exact in fp16 two-term form, f32 accumulation
 * drained to f64 every 1024 rows.  It replaces the Gram of nalgebra's qr()/lu() replacement (CholeskyQR) on the tall
 * side of single-svdlib's power iteration (call site pca/sparse/mod.rs:170-180).  panel: host m x k row-major, or NULL
 * for a device-generated normal panel of device_rows rows (timing); avg_ms: average kernel time over iters launches. */
int salg_op_tall_gram_f32(salg_ctx* ctx, const float* panel, int64_t m, int64_t k, int64_t device_rows, double* gram,
                          double* colsum, int iters, double* avg_ms);

/* Repeats one centred SpMM (transposed or not) `iters` times on a device-resident random panel and
 * returns the average device ms — a microbenchmark of the product kernels (tools/scripts_tc_*.py). */
int salg_op_spmm_bench(salg_ctx* ctx, const salg_csr* csr, int transposed, int64_t k, int iters,
                       double* avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* SALG_H */

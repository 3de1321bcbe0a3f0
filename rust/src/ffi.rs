//! Raw bindings: one `extern "C"` item per entry point of include/salg.h (same order as the header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct salg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct salg_csr { _p: [u8; 0] }
#[repr(C)] pub struct salg_pca { _p: [u8; 0] }

pub const SALG_OK: c_int = 0;
pub const SALG_ROW: c_int = 0;
pub const SALG_COLUMN: c_int = 1;
pub const SALG_SVD_LANCZOS: i32 = 0;
pub const SALG_SVD_RANDOM: i32 = 1;
pub const SALG_TRANSFORM_EXACT: c_int = 0;
pub const SALG_TRANSFORM_REFERENCE_COMPAT: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct salg_pca_params {
    pub n_components: i32, pub svd_method: i32, pub n_oversamples: i32, pub n_power_iterations: i32,
    pub normalizer: i32, pub center: i32, pub verbose: i32, pub random_seed: u32,
    pub alpha: f64, pub tolerance: f64, pub lanczos_max_steps: i32, pub keep_scores: i32, pub reserved: [i32; 6],
}

extern "C" {
    // errors / lifecycle
    pub fn salg_last_error() -> *const c_char;
    pub fn salg_version() -> c_int;
    pub fn salg_device_count(out: *mut c_int) -> c_int;
    pub fn salg_ctx_create(device: c_int, out: *mut *mut salg_ctx) -> c_int;
    pub fn salg_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn salg_ctx_create_dist(device: c_int, rank: c_int, nranks: c_int, id: *const c_void, out: *mut *mut salg_ctx) -> c_int;
    pub fn salg_ctx_destroy(ctx: *mut salg_ctx) -> c_int;
    pub fn salg_ctx_sync(ctx: *mut salg_ctx) -> c_int;
    pub fn salg_ctx_rank(ctx: *const salg_ctx, rank: *mut c_int, nranks: *mut c_int) -> c_int;
    pub fn salg_timer_start(ctx: *mut salg_ctx) -> c_int;
    pub fn salg_timer_stop(ctx: *mut salg_ctx, ms: *mut f64) -> c_int;
    pub fn salg_ctx_set_spmm_impl(ctx: *mut salg_ctx, implementation: c_int) -> c_int;
    pub fn salg_launch_count(ctx: *mut salg_ctx, out: *mut i64) -> c_int;
    // CSR / CSC containers (nalgebra-sparse hands out &[usize]; on 64-bit targets that is the u64 layout of the ABI)
    pub fn salg_csr_upload_f32(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const f32, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_upload_f64(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const f64, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_upload_i32_f32(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, off: *const i64, idx: *const i32, val: *const f32, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_upload_i32_f64(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, off: *const i64, idx: *const i32, val: *const f64, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csc_upload_f32(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, col_off: *const u64, row_idx: *const u64, val: *const f32, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csc_upload_f64(ctx: *mut salg_ctx, nrows: i64, ncols: i64, nnz: i64, col_off: *const u64, row_idx: *const u64, val: *const f64, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_free(csr: *mut salg_csr) -> c_int;
    pub fn salg_csr_dims(csr: *const salg_csr, nrows: *mut i64, ncols: *mut i64, nnz: *mut i64, dtype: *mut c_int) -> c_int;
    pub fn salg_csr_download_f32(ctx: *mut salg_ctx, csr: *const salg_csr, off: *mut u64, idx: *mut u64, val: *mut f32) -> c_int;
    pub fn salg_csr_download_f64(ctx: *mut salg_ctx, csr: *const salg_csr, off: *mut u64, idx: *mut u64, val: *mut f64) -> c_int;
    pub fn salg_csr_download_raw(ctx: *mut salg_ctx, csr: *const salg_csr, off: *mut i64, idx: *mut u32, val: *mut c_void) -> c_int;
    pub fn salg_csr_select_columns(ctx: *mut salg_ctx, csr: *const salg_csr, mask: *const u8, mask_len: i64, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_transpose(ctx: *mut salg_ctx, csr: *const salg_csr, out: *mut *mut salg_csr) -> c_int;
    pub fn salg_csr_values_clone(ctx: *mut salg_ctx, csr: *const salg_csr, clone: *mut *mut c_void) -> c_int;
    pub fn salg_csr_values_restore(ctx: *mut salg_ctx, csr: *mut salg_csr, clone: *const c_void) -> c_int;
    pub fn salg_dev_free(ctx: *mut salg_ctx, p: *mut c_void) -> c_int;
    // MatrixSum / MatrixNonZero / MatrixVariance
    pub fn salg_sum_col_f32(ctx: *mut salg_ctx, csr: *const salg_csr, sum: *mut f32, sumsq: *mut f32) -> c_int;
    pub fn salg_sum_col_f64(ctx: *mut salg_ctx, csr: *const salg_csr, sum: *mut f64, sumsq: *mut f64) -> c_int;
    pub fn salg_sum_row_f32(ctx: *mut salg_ctx, csr: *const salg_csr, out: *mut f32) -> c_int;
    pub fn salg_sum_row_f64(ctx: *mut salg_ctx, csr: *const salg_csr, out: *mut f64) -> c_int;
    pub fn salg_col_stats_f64(ctx: *mut salg_ctx, csr: *const salg_csr, sum: *mut f64, sumsq: *mut f64, nnz_col: *mut f64, var_col: *mut f64) -> c_int;
    pub fn salg_nonzero_row(ctx: *mut salg_ctx, csr: *const salg_csr, out: *mut u64) -> c_int;
    pub fn salg_nonzero_col(ctx: *mut salg_ctx, csr: *const salg_csr, out: *mut u64) -> c_int;
    pub fn salg_csc_sum_col_f32(ctx: *mut salg_ctx, csc: *const salg_csr, sum: *mut f32, sumsq: *mut f32) -> c_int;
    pub fn salg_csc_sum_col_f64(ctx: *mut salg_ctx, csc: *const salg_csr, sum: *mut f64, sumsq: *mut f64) -> c_int;
    pub fn salg_csc_sum_row_f32(ctx: *mut salg_ctx, csc: *const salg_csr, out: *mut f32) -> c_int;
    pub fn salg_csc_sum_row_f64(ctx: *mut salg_ctx, csc: *const salg_csr, out: *mut f64) -> c_int;
    // Normalize / Log1P
    pub fn salg_normalize_f32(ctx: *mut salg_ctx, csr: *mut salg_csr, sums: *const f32, n: i64, target: f32, direction: c_int) -> c_int;
    pub fn salg_normalize_f64(ctx: *mut salg_ctx, csr: *mut salg_csr, sums: *const f64, n: i64, target: f64, direction: c_int) -> c_int;
    pub fn salg_normalize_f32_u64(ctx: *mut salg_ctx, csr: *mut salg_csr, sums: *const f64, n: i64, target: f64, direction: c_int) -> c_int;
    pub fn salg_csc_normalize_f32(ctx: *mut salg_ctx, csc: *mut salg_csr, sums: *const f32, n: i64, target: f32, direction: c_int) -> c_int;
    pub fn salg_csc_normalize_f64(ctx: *mut salg_ctx, csc: *mut salg_csr, sums: *const f64, n: i64, target: f64, direction: c_int) -> c_int;
    pub fn salg_csc_normalize_f32_u64(ctx: *mut salg_ctx, csc: *mut salg_csr, sums: *const f64, n: i64, target: f64, direction: c_int) -> c_int;
    pub fn salg_log1p(ctx: *mut salg_ctx, csr: *mut salg_csr) -> c_int;
    pub fn salg_preprocess_f32(ctx: *mut salg_ctx, csr: *mut salg_csr, target: f32, col_sum: *mut f32, col_sumsq: *mut f32) -> c_int;
    pub fn salg_preprocess_f64(ctx: *mut salg_ctx, csr: *mut salg_csr, target: f64, col_sum: *mut f64, col_sumsq: *mut f64) -> c_int;
    // PCA
    pub fn salg_pca_params_default(p: *mut salg_pca_params) -> c_int;
    pub fn salg_pca_fit_f32(ctx: *mut salg_ctx, x: *const salg_csr, p: *const salg_pca_params, mask: *const u8, mask_len: i64, omega: *const f32, omega_rows: i64, omega_cols: i64, out: *mut *mut salg_pca) -> c_int;
    pub fn salg_pca_fit_f64(ctx: *mut salg_ctx, x: *const salg_csr, p: *const salg_pca_params, mask: *const u8, mask_len: i64, omega: *const f64, omega_rows: i64, omega_cols: i64, out: *mut *mut salg_pca) -> c_int;
    pub fn salg_pca_free(p: *mut salg_pca) -> c_int;
    pub fn salg_pca_dims(p: *const salg_pca, d: *mut i64, n_eff: *mut i64, ncols: *mut i64, dtype: *mut c_int) -> c_int;
    pub fn salg_pca_components_f32(p: *const salg_pca, out: *mut f32) -> c_int;
    pub fn salg_pca_components_f64(p: *const salg_pca, out: *mut f64) -> c_int;
    pub fn salg_pca_singular_values_f64(p: *const salg_pca, out: *mut f64) -> c_int;
    pub fn salg_pca_explained_variance_f64(p: *const salg_pca, out: *mut f64) -> c_int;
    pub fn salg_pca_mean_f64(p: *const salg_pca, out: *mut f64) -> c_int;
    pub fn salg_pca_total_var(p: *const salg_pca, out: *mut f64) -> c_int;
    pub fn salg_pca_n_samples(p: *const salg_pca, out: *mut i64) -> c_int;
    pub fn salg_pca_numeric_flags(p: *const salg_pca, out: *mut c_int) -> c_int;
    pub fn salg_pca_transform_f32(ctx: *mut salg_ctx, p: *const salg_pca, x: *const salg_csr, mode: c_int, scores: *mut f32) -> c_int;
    pub fn salg_pca_transform_f64(ctx: *mut salg_ctx, p: *const salg_pca, x: *const salg_csr, mode: c_int, scores: *mut f64) -> c_int;
    pub fn salg_pca_fit_scores_f32(ctx: *mut salg_ctx, p: *const salg_pca, scores: *mut f32) -> c_int;
    pub fn salg_pca_fit_scores_f64(ctx: *mut salg_ctx, p: *const salg_pca, scores: *mut f64) -> c_int;
}

/// Non-zero status -> `anyhow::Error` carrying the library's message: the reference's own strings where it has one
/// (mask length, "Must be fitted before transform!", "SVD computation failed: ...").
pub fn check(status: c_int) -> anyhow::Result<()> {
    if status == SALG_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(salg_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!(msg))
}

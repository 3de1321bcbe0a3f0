//! Context singleton, RAII device handles and the sealed `SalgFloat` trait that picks the `_f32` / `_f64` symbol for the
//! reference's generic `T`.
use crate::ffi::*;
use nalgebra_sparse::{CscMatrix, CsrMatrix};
use std::os::raw::c_int;
use std::sync::OnceLock;

struct CtxPtr(*mut salg_ctx);
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {}
static CTX: OnceLock<CtxPtr> = OnceLock::new();

/// Process-wide single-GPU context on device `SALG_DEVICE` (default 0).  A `salg_ctx` is not safe for concurrent calls:
/// the facade's `&mut self` / `&self` methods serialise through the caller, as the reference's do through Rayon's pool.
pub fn ctx() -> *mut salg_ctx {
    CTX.get_or_init(|| {
        let dev: c_int = std::env::var("SALG_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p = std::ptr::null_mut();
        check(unsafe { salg_ctx_create(dev, &mut p) }).expect("salg_ctx_create (no CPU fallback: a CUDA device is required)");
        CtxPtr(p)
    }).0
}

mod sealed { pub trait Sealed {} impl Sealed for f32 {} impl Sealed for f64 {} }

/// One function per suffix pair of the ABI.
pub trait SalgFloat: sealed::Sealed + Copy + num_traits::Float + 'static {
    unsafe fn csr_upload(nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const Self, out: *mut *mut salg_csr) -> c_int;
    unsafe fn csc_upload(nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const Self, out: *mut *mut salg_csr) -> c_int;
    unsafe fn download_values(csr: *const salg_csr, val: *mut Self) -> c_int;
    unsafe fn sum_col(csr: *const salg_csr, sum: *mut Self, sumsq: *mut Self) -> c_int;
    unsafe fn sum_row(csr: *const salg_csr, out: *mut Self) -> c_int;
    unsafe fn csc_sum_col(csc: *const salg_csr, sum: *mut Self, sumsq: *mut Self) -> c_int;
    unsafe fn csc_sum_row(csc: *const salg_csr, out: *mut Self) -> c_int;
    unsafe fn normalize(csr: *mut salg_csr, sums: *const Self, n: i64, target: Self, direction: c_int) -> c_int;
    unsafe fn csc_normalize(csc: *mut salg_csr, sums: *const Self, n: i64, target: Self, direction: c_int) -> c_int;
    unsafe fn preprocess(csr: *mut salg_csr, target: Self, sum: *mut Self, sumsq: *mut Self) -> c_int;
    unsafe fn pca_fit(x: *const salg_csr, p: *const salg_pca_params, mask: *const u8, mask_len: i64, omega: *const Self, orows: i64, ocols: i64, out: *mut *mut salg_pca) -> c_int;
    unsafe fn pca_components(p: *const salg_pca, out: *mut Self) -> c_int;
    unsafe fn pca_transform(p: *const salg_pca, x: *const salg_csr, mode: c_int, scores: *mut Self) -> c_int;
    unsafe fn pca_fit_scores(p: *const salg_pca, scores: *mut Self) -> c_int;
}

macro_rules! impl_salg_float {
    ($t:ty, $up:ident, $cup:ident, $dl:ident, $sc:ident, $sr:ident, $csc:ident, $csr_:ident, $nm:ident, $cnm:ident, $pp:ident,
     $fit:ident, $comp:ident, $tr:ident, $fs:ident) => {
        impl SalgFloat for $t {
            unsafe fn csr_upload(nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const Self, out: *mut *mut salg_csr) -> c_int { $up(ctx(), nrows, ncols, nnz, off, idx, val, out) }
            unsafe fn csc_upload(nrows: i64, ncols: i64, nnz: i64, off: *const u64, idx: *const u64, val: *const Self, out: *mut *mut salg_csr) -> c_int { $cup(ctx(), nrows, ncols, nnz, off, idx, val, out) }
            unsafe fn download_values(csr: *const salg_csr, val: *mut Self) -> c_int { $dl(ctx(), csr, std::ptr::null_mut(), std::ptr::null_mut(), val) }
            unsafe fn sum_col(csr: *const salg_csr, sum: *mut Self, sumsq: *mut Self) -> c_int { $sc(ctx(), csr, sum, sumsq) }
            unsafe fn sum_row(csr: *const salg_csr, out: *mut Self) -> c_int { $sr(ctx(), csr, out) }
            unsafe fn csc_sum_col(csc: *const salg_csr, sum: *mut Self, sumsq: *mut Self) -> c_int { $csc(ctx(), csc, sum, sumsq) }
            unsafe fn csc_sum_row(csc: *const salg_csr, out: *mut Self) -> c_int { $csr_(ctx(), csc, out) }
            unsafe fn normalize(csr: *mut salg_csr, sums: *const Self, n: i64, target: Self, direction: c_int) -> c_int { $nm(ctx(), csr, sums, n, target, direction) }
            unsafe fn csc_normalize(csc: *mut salg_csr, sums: *const Self, n: i64, target: Self, direction: c_int) -> c_int { $cnm(ctx(), csc, sums, n, target, direction) }
            unsafe fn preprocess(csr: *mut salg_csr, target: Self, sum: *mut Self, sumsq: *mut Self) -> c_int { $pp(ctx(), csr, target, sum, sumsq) }
            unsafe fn pca_fit(x: *const salg_csr, p: *const salg_pca_params, mask: *const u8, mask_len: i64, omega: *const Self, orows: i64, ocols: i64, out: *mut *mut salg_pca) -> c_int { $fit(ctx(), x, p, mask, mask_len, omega, orows, ocols, out) }
            unsafe fn pca_components(p: *const salg_pca, out: *mut Self) -> c_int { $comp(p, out) }
            unsafe fn pca_transform(p: *const salg_pca, x: *const salg_csr, mode: c_int, scores: *mut Self) -> c_int { $tr(ctx(), p, x, mode, scores) }
            unsafe fn pca_fit_scores(p: *const salg_pca, scores: *mut Self) -> c_int { $fs(ctx(), p, scores) }
        }
    };
}
impl_salg_float!(f32, salg_csr_upload_f32, salg_csc_upload_f32, salg_csr_download_f32, salg_sum_col_f32, salg_sum_row_f32,
                 salg_csc_sum_col_f32, salg_csc_sum_row_f32, salg_normalize_f32, salg_csc_normalize_f32, salg_preprocess_f32,
                 salg_pca_fit_f32, salg_pca_components_f32, salg_pca_transform_f32, salg_pca_fit_scores_f32);
impl_salg_float!(f64, salg_csr_upload_f64, salg_csc_upload_f64, salg_csr_download_f64, salg_sum_col_f64, salg_sum_row_f64,
                 salg_csc_sum_col_f64, salg_csc_sum_row_f64, salg_normalize_f64, salg_csc_normalize_f64, salg_preprocess_f64,
                 salg_pca_fit_f64, salg_pca_components_f64, salg_pca_transform_f64, salg_pca_fit_scores_f64);

/// Device-resident copy of a `CsrMatrix` / `CscMatrix` (the CSC one is stored as the CSR of A^T); freed on drop.
pub struct DeviceCsr { raw: *mut salg_csr }
impl DeviceCsr {
    pub fn upload<T: SalgFloat>(x: &CsrMatrix<T>) -> anyhow::Result<Self> {
        let mut raw = std::ptr::null_mut();
        // usize == u64 on every target the library supports
        check(unsafe { T::csr_upload(x.nrows() as i64, x.ncols() as i64, x.nnz() as i64, x.row_offsets().as_ptr() as *const u64,
                                      x.col_indices().as_ptr() as *const u64, x.values().as_ptr(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn upload_csc<T: SalgFloat>(x: &CscMatrix<T>) -> anyhow::Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { T::csc_upload(x.nrows() as i64, x.ncols() as i64, x.nnz() as i64, x.col_offsets().as_ptr() as *const u64,
                                      x.row_indices().as_ptr() as *const u64, x.values().as_ptr(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn raw(&self) -> *mut salg_csr { self.raw }
    pub fn download_values_into<T: SalgFloat>(&self, values: &mut [T]) -> anyhow::Result<()> {
        check(unsafe { T::download_values(self.raw, values.as_mut_ptr()) })
    }
}
impl Drop for DeviceCsr { fn drop(&mut self) { unsafe { salg_csr_free(self.raw); } } }

/// Fitted model handle; freed on drop.
pub struct Model { pub(crate) raw: *mut salg_pca }
impl Model {
    pub fn dims(&self) -> anyhow::Result<(usize, usize, usize)> {
        let (mut d, mut n_eff, mut ncols, mut dt) = (0i64, 0i64, 0i64, 0 as c_int);
        check(unsafe { salg_pca_dims(self.raw, &mut d, &mut n_eff, &mut ncols, &mut dt) })?;
        Ok((d as usize, n_eff as usize, ncols as usize))
    }
    pub fn components<T: SalgFloat>(&self) -> anyhow::Result<Vec<T>> {
        let (d, n_eff, _) = self.dims()?;
        let mut out = vec![T::zero(); d * n_eff];
        check(unsafe { T::pca_components(self.raw, out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn explained_variance(&self) -> anyhow::Result<Vec<f64>> {
        let (d, _, _) = self.dims()?;
        let mut out = vec![0f64; d];
        check(unsafe { salg_pca_explained_variance_f64(self.raw, out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn mean(&self) -> anyhow::Result<Vec<f64>> {
        let (_, _, ncols) = self.dims()?;
        let mut out = vec![0f64; ncols];
        check(unsafe { salg_pca_mean_f64(self.raw, out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn total_var(&self) -> anyhow::Result<f64> {
        let mut v = 0f64;
        check(unsafe { salg_pca_total_var(self.raw, &mut v) })?;
        Ok(v)
    }
}
impl Drop for Model { fn drop(&mut self) { unsafe { salg_pca_free(self.raw); } } }

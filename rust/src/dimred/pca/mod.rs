//! `single_algebra::dimred::pca` (src/dimred/pca/mod.rs): module root, `SVDMethod`, re-exports.
mod sparse;
mod sparse_masked;

pub use sparse::{SparsePCA, SparsePCABuilder};
pub use sparse_masked::{MaskedSparsePCA, MaskedSparsePCABuilder};

pub use crate::device::SalgFloat as SvdFloat;   // the reference re-exports single_svdlib::SvdFloat here (:42)

/// `single_svdlib::randomized::PowerIterationNormalizer` (re-export src/dimred/pca/mod.rs:41).  QR and LU both run
/// CholeskyQR2 on the device (same column space); `None` skips the tall-side factorisation.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum PowerIterationNormalizer { QR = 0, LU = 1, None = 2 }

/// SVD computation method for PCA (src/dimred/pca/mod.rs:49-62); default `Lanczos` (:64-68).
#[derive(Debug, Clone, Copy, PartialEq)]
pub enum SVDMethod {
    Lanczos,
    Random { n_oversamples: usize, n_power_iterations: usize, normalizer: PowerIterationNormalizer },
}
impl Default for SVDMethod { fn default() -> Self { Self::Lanczos } }

/// What `transform` computes.  `Exact` = the intended projection (X - 1 mu^T) V^T on the kept
/// columns; `ReferenceCompat` = what the reference's loops compute, bug for bug.  Not in the reference API.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum TransformMode { Exact = 0, ReferenceCompat = 1 }

pub(crate) fn fill_params(p: &mut crate::ffi::salg_pca_params, n_components: usize, alpha: f64, tolerance: f64, seed: u32,
                          center: bool, verbose: bool, m: &SVDMethod, keep_scores: bool) {
    p.n_components = n_components as i32;
    p.alpha = alpha;
    p.tolerance = tolerance;
    p.random_seed = seed;
    p.center = center as i32;
    p.verbose = verbose as i32;
    p.keep_scores = keep_scores as i32;
    match *m {
        SVDMethod::Lanczos => p.svd_method = crate::ffi::SALG_SVD_LANCZOS,
        SVDMethod::Random { n_oversamples, n_power_iterations, normalizer } => {
            p.svd_method = crate::ffi::SALG_SVD_RANDOM;
            p.n_oversamples = n_oversamples as i32;
            p.n_power_iterations = n_power_iterations as i32;
            p.normalizer = normalizer as i32;
        }
    }
}

//! `SparsePCA<T>` / `SparsePCABuilder<T>` (src/dimred/pca/sparse/mod.rs:33-484) with FFI bodies.
use super::{fill_params, SVDMethod, TransformMode};
use crate::device::{ctx, DeviceCsr, Model, SalgFloat};
use crate::ffi::*;
use anyhow::anyhow;
use nalgebra_sparse::CsrMatrix;
use ndarray::{Array1, Array2};
use num_traits::NumCast;

pub struct SparsePCA<T: SalgFloat> {
    n_components: usize,
    alpha: T,
    tolerance: T,
    random_seed: u32,
    components_: Option<Array2<T>>,
    explained_variance_: Option<Array1<T>>,
    mean_: Option<Array1<T>>,
    center: bool,
    verbose: bool,
    svdmethod: SVDMethod,
    model: Option<Model>,
    /// host Gaussian test matrix n_features x (n_components + n_oversamples), row-major — the "same host-generated
    /// Omega" of the parity contract; None => drawn on the device from `random_seed`
    pub omega: Option<Array2<T>>,
    pub transform_mode: TransformMode,
    total_var_: Option<f64>,
}

fn to_arr1<T: SalgFloat>(v: &[f64]) -> Array1<T> { Array1::from_iter(v.iter().map(|&x| <T as NumCast>::from(x).unwrap())) }

impl<T: SalgFloat> SparsePCA<T> {
    /// pca/sparse/mod.rs:63-84 (the parameter really is spelled `tollerance`)
    pub fn new(n_components: usize, alpha: T, tollerance: Option<T>, random_seed: Option<u32>, center: bool, verbose: bool,
               svdmethod: SVDMethod) -> Self {
        Self { n_components, alpha, tolerance: tollerance.unwrap_or(<T as NumCast>::from(1e-6).unwrap()),
               random_seed: random_seed.unwrap_or(42), components_: None, explained_variance_: None, mean_: None, center,
               verbose, svdmethod, model: None, omega: None, transform_mode: TransformMode::Exact, total_var_: None }
    }

    fn fit_impl(&mut self, x: &CsrMatrix<T>, keep_scores: bool) -> anyhow::Result<()> {
        let dev = DeviceCsr::upload(x)?;
        let mut p: salg_pca_params = unsafe { std::mem::zeroed() };
        check(unsafe { salg_pca_params_default(&mut p) })?;
        fill_params(&mut p, self.n_components, self.alpha.to_f64().unwrap(), self.tolerance.to_f64().unwrap(), self.random_seed,
                    self.center, self.verbose, &self.svdmethod, keep_scores);
        let (om, orows, ocols) = match &self.omega {
            Some(o) => (o.as_slice().ok_or_else(|| anyhow!("omega must be contiguous row-major"))?.as_ptr(), o.nrows() as i64, o.ncols() as i64),
            None => (std::ptr::null(), 0, 0),
        };
        let mut raw = std::ptr::null_mut();
        // the library's message already reads "SVD computation failed: ..." (:144) / "Randomized SVD computation failed: ..." (:180)
        check(unsafe { T::pca_fit(dev.raw(), &p, std::ptr::null(), 0, om, orows, ocols, &mut raw) })?;
        let model = Model { raw };
        let (d, n_eff, _) = model.dims()?;
        self.components_ = Some(Array2::from_shape_vec((d, n_eff), model.components::<T>()?)?);      // :208
        self.explained_variance_ = Some(to_arr1(&model.explained_variance()?));                      // :210-216
        self.mean_ = Some(to_arr1(&model.mean()?));                                                  // :106-117
        self.total_var_ = Some(model.total_var()?);
        self.model = Some(model);
        Ok(())
    }

    /// `fit(&mut self, x)` — pca/sparse/mod.rs:102-242
    pub fn fit(&mut self, x: &CsrMatrix<T>) -> anyhow::Result<&mut Self> {
        self.fit_impl(x, false)?;
        Ok(self)
    }

    /// `transform(&self, x)` — pca/sparse/mod.rs:255-285
    pub fn transform(&self, x: &CsrMatrix<T>) -> anyhow::Result<Array2<T>> {
        let model = self.model.as_ref().ok_or_else(|| anyhow!("Must be fitted before transform!"))?;   // :259
        let (d, _, _) = model.dims()?;
        let dev = DeviceCsr::upload(x)?;
        let mut out = vec![T::zero(); x.nrows() * d];
        check(unsafe { T::pca_transform(model.raw, dev.raw(), self.transform_mode as i32, out.as_mut_ptr()) })?;
        Ok(Array2::from_shape_vec((x.nrows(), d), out)?)
    }

    /// `fit_transform(&mut self, x)` — pca/sparse/mod.rs:355-358: fit + transform; under `Exact` the projection is computed
    /// inside the fit call while the operator is resident on the device (one upload instead of two)
    pub fn fit_transform(&mut self, x: &CsrMatrix<T>) -> anyhow::Result<Array2<T>> {
        if self.transform_mode != TransformMode::Exact {
            self.fit_impl(x, false)?;
            return self.transform(x);
        }
        self.fit_impl(x, true)?;
        let model = self.model.as_ref().unwrap();
        let (d, _, _) = model.dims()?;
        let mut out = vec![T::zero(); x.nrows() * d];
        check(unsafe { T::pca_fit_scores(model.raw, out.as_mut_ptr()) })?;
        Ok(Array2::from_shape_vec((x.nrows(), d), out)?)
    }

    /// pca/sparse/mod.rs:295-302
    pub fn feature_importances(&self) -> anyhow::Result<Array2<T>> {
        let c = self.components_.as_ref().ok_or_else(|| anyhow!("Model must be fitted first!"))?;
        Ok(c.mapv(|x| x * x))
    }
    /// pca/sparse/mod.rs:312-322 — normalised by the sum over the COMPUTED components
    pub fn explained_variance_ratio(&self) -> anyhow::Result<Array1<T>> {
        let ev = self.explained_variance_.as_ref().ok_or_else(|| anyhow!("Model must be fitted first!"))?;
        let total = ev.iter().fold(T::zero(), |a, &b| a + b);
        Ok(ev.mapv(|v| v / total))
    }
    /// pca/sparse/mod.rs:333-343
    pub fn cumulative_explained_variance_ratio(&self) -> anyhow::Result<Array1<T>> {
        let r = self.explained_variance_ratio()?;
        let mut sum = T::zero();
        Ok(r.mapv(|v| { sum = sum + v; sum }))
    }
    pub fn components(&self) -> Option<&Array2<T>> { self.components_.as_ref() }
    pub fn explained_variance(&self) -> Option<&Array1<T>> { self.explained_variance_.as_ref() }
    pub fn mean(&self) -> Option<&Array1<T>> { self.mean_.as_ref() }
    /// total variance of the centred columns — what the reference only prints under `verbose` (:225-238)
    pub fn total_variance(&self) -> Option<f64> { self.total_var_ }
}

/// `SparsePCABuilder<T>` — pca/sparse/mod.rs:375-484, defaults :388-403
pub struct SparsePCABuilder<T: SalgFloat> {
    n_components: usize, alpha: T, tolerance: T, random_seed: Option<u32>, center: bool, verbose: bool, svdmethod: SVDMethod,
}
impl<T: SalgFloat> Default for SparsePCABuilder<T> {
    fn default() -> Self {
        Self { n_components: 50, alpha: <T as NumCast>::from(1.0).unwrap(), tolerance: <T as NumCast>::from(1e-6).unwrap(),
               random_seed: Some(42), center: true, verbose: false, svdmethod: SVDMethod::default() }
    }
}
impl<T: SalgFloat> SparsePCABuilder<T> {
    pub fn new() -> Self { Self::default() }
    pub fn n_components(mut self, n: usize) -> Self { self.n_components = n; self }
    pub fn alpha(mut self, a: T) -> Self { self.alpha = a; self }
    pub fn tolerance(mut self, t: T) -> Self { self.tolerance = t; self }
    pub fn random_seed(mut self, s: u32) -> Self { self.random_seed = Some(s); self }
    pub fn center(mut self, c: bool) -> Self { self.center = c; self }
    pub fn verbose(mut self, v: bool) -> Self { self.verbose = v; self }
    pub fn svd_method(mut self, m: SVDMethod) -> Self { self.svdmethod = m; self }
    pub fn build(self) -> SparsePCA<T> {
        SparsePCA::new(self.n_components, self.alpha, Some(self.tolerance), self.random_seed, self.center, self.verbose,
                       self.svdmethod)
    }
}

//! `MaskedSparsePCA<T>` / `MaskedSparsePCABuilder<T>` (src/dimred/pca/sparse_masked/mod.rs:37-619) with FFI bodies.
//! `Vec<bool>` crosses the ABI as it is: a Rust `bool` is one byte holding 0 or 1.
use super::{fill_params, SVDMethod, TransformMode};
use crate::device::{DeviceCsr, Model, SalgFloat};
use crate::ffi::*;
use anyhow::anyhow;
use nalgebra_sparse::CsrMatrix;
use ndarray::{Array1, Array2};
use num_traits::NumCast;

pub struct MaskedSparsePCA<T: SalgFloat> {
    n_components: usize,
    alpha: T,
    tolerance: T,
    random_seed: u32,
    components_: Option<Array2<T>>,          // d x n_masked (:368)
    explained_variance_: Option<Array1<T>>,
    mean_: Option<Array1<T>>,                // FULL column count (:280-291)
    mask: Vec<bool>,
    center: bool,
    verbose: bool,
    svd_method: SVDMethod,
    model: Option<Model>,
    pub omega: Option<Array2<T>>,
    pub transform_mode: TransformMode,
    total_var_: Option<f64>,
}

fn to_arr1<T: SalgFloat>(v: &[f64]) -> Array1<T> { Array1::from_iter(v.iter().map(|&x| <T as NumCast>::from(x).unwrap())) }
const MASK_MSG: &str = "The mask vector length and the number of features (columns) have to be the same!";

impl<T: SalgFloat> MaskedSparsePCA<T> {
    /// pca/sparse_masked/mod.rs:214-237
    #[allow(clippy::too_many_arguments)]
    pub fn new(n_components: usize, alpha: T, tollerance: Option<T>, random_seed: Option<u32>, mask: Vec<bool>, center: bool,
               verbose: bool, svd_method: SVDMethod) -> Self {
        Self { n_components, alpha, tolerance: tollerance.unwrap_or(<T as NumCast>::from(1e-6).unwrap()),
               random_seed: random_seed.unwrap_or(42), components_: None, explained_variance_: None, mean_: None, mask, center,
               verbose, svd_method, model: None, omega: None, transform_mode: TransformMode::Exact, total_var_: None }
    }

    fn fit_impl(&mut self, x: &CsrMatrix<T>, keep_scores: bool) -> anyhow::Result<()> {
        if self.mask.len() != x.ncols() { return Err(anyhow!(MASK_MSG)); }                          // :258-262
        let dev = DeviceCsr::upload(x)?;
        let mut p: salg_pca_params = unsafe { std::mem::zeroed() };
        check(unsafe { salg_pca_params_default(&mut p) })?;
        fill_params(&mut p, self.n_components, self.alpha.to_f64().unwrap(), self.tolerance.to_f64().unwrap(), self.random_seed,
                    self.center, self.verbose, &self.svd_method, keep_scores);
        let (om, orows, ocols) = match &self.omega {
            Some(o) => (o.as_slice().ok_or_else(|| anyhow!("omega must be contiguous row-major"))?.as_ptr(), o.nrows() as i64, o.ncols() as i64),
            None => (std::ptr::null(), 0, 0),
        };
        let mut raw = std::ptr::null_mut();
        check(unsafe { T::pca_fit(dev.raw(), &p, self.mask.as_ptr() as *const u8, self.mask.len() as i64, om, orows, ocols, &mut raw) })?;
        let model = Model { raw };
        let (d, n_eff, _) = model.dims()?;
        self.components_ = Some(Array2::from_shape_vec((d, n_eff), model.components::<T>()?)?);
        self.explained_variance_ = Some(to_arr1(&model.explained_variance()?));                      // :379-382
        self.mean_ = Some(to_arr1(&model.mean()?));
        self.total_var_ = Some(model.total_var()?);
        self.model = Some(model);
        Ok(())
    }

    /// `fit(&mut self, x)` — pca/sparse_masked/mod.rs:255-419
    pub fn fit(&mut self, x: &CsrMatrix<T>) -> anyhow::Result<&mut Self> {
        self.fit_impl(x, false)?;
        Ok(self)
    }

    /// `transform(&self, x)` — pca/sparse_masked/mod.rs:438-546 (the mask length is checked before the fitted state, :440-444)
    pub fn transform(&self, x: &CsrMatrix<T>) -> anyhow::Result<Array2<T>> {
        if self.mask.len() != x.ncols() { return Err(anyhow!(MASK_MSG)); }
        let model = self.model.as_ref().ok_or_else(|| anyhow!("Must be fitted before transform!"))?;
        let (d, _, _) = model.dims()?;
        let dev = DeviceCsr::upload(x)?;
        let mut out = vec![T::zero(); x.nrows() * d];
        check(unsafe { T::pca_transform(model.raw, dev.raw(), self.transform_mode as i32, out.as_mut_ptr()) })?;
        Ok(Array2::from_shape_vec((x.nrows(), d), out)?)
    }

    /// `fit_transform(&mut self, x)` — pca/sparse_masked/mod.rs:616-619
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

    pub fn feature_importances(&self) -> anyhow::Result<Array2<T>> {
        let c = self.components_.as_ref().ok_or_else(|| anyhow!("Model must be fitted first!"))?;
        Ok(c.mapv(|x| x * x))
    }
    pub fn explained_variance_ratio(&self) -> anyhow::Result<Array1<T>> {
        let ev = self.explained_variance_.as_ref().ok_or_else(|| anyhow!("Model must be fitted first!"))?;
        let total = ev.iter().fold(T::zero(), |a, &b| a + b);
        Ok(ev.mapv(|v| v / total))
    }
    pub fn cumulative_explained_variance_ratio(&self) -> anyhow::Result<Array1<T>> {
        let r = self.explained_variance_ratio()?;
        let mut sum = T::zero();
        Ok(r.mapv(|v| { sum = sum + v; sum }))
    }
    pub fn components(&self) -> Option<&Array2<T>> { self.components_.as_ref() }
    pub fn explained_variance(&self) -> Option<&Array1<T>> { self.explained_variance_.as_ref() }
    pub fn mean(&self) -> Option<&Array1<T>> { self.mean_.as_ref() }
    pub fn total_variance(&self) -> Option<f64> { self.total_var_ }
}

/// `MaskedSparsePCABuilder<T>` — pca/sparse_masked/mod.rs:37-160, defaults :51-67
pub struct MaskedSparsePCABuilder<T: SalgFloat> {
    n_components: usize, alpha: T, tolerance: T, random_seed: Option<u32>, center: bool, verbose: bool, mask: Vec<bool>,
    svdmethod: SVDMethod,
}
impl<T: SalgFloat> Default for MaskedSparsePCABuilder<T> {
    fn default() -> Self {
        Self { n_components: 50, alpha: <T as NumCast>::from(1.0).unwrap(), tolerance: <T as NumCast>::from(1e-6).unwrap(),
               random_seed: Some(42), center: true, verbose: false, mask: Vec::new(), svdmethod: SVDMethod::default() }
    }
}
impl<T: SalgFloat> MaskedSparsePCABuilder<T> {
    pub fn new() -> Self { Self::default() }
    pub fn n_components(mut self, n: usize) -> Self { self.n_components = n; self }
    pub fn alpha(mut self, a: T) -> Self { self.alpha = a; self }
    pub fn tolerance(mut self, t: T) -> Self { self.tolerance = t; self }
    pub fn random_seed(mut self, s: u32) -> Self { self.random_seed = Some(s); self }
    pub fn center(mut self, c: bool) -> Self { self.center = c; self }
    pub fn verbose(mut self, v: bool) -> Self { self.verbose = v; self }
    pub fn mask(mut self, m: Vec<bool>) -> Self { self.mask = m; self }
    pub fn svd_method(mut self, m: SVDMethod) -> Self { self.svdmethod = m; self }
    pub fn build(self) -> MaskedSparsePCA<T> {
        MaskedSparsePCA::new(self.n_components, self.alpha, Some(self.tolerance), self.random_seed, self.mask, self.center,
                             self.verbose, self.svdmethod)
    }
}

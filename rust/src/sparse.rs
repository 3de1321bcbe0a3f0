//! `single_algebra::sparse` — the traits on the hot path (src/sparse/mod.rs:35-102) with FFI bodies.
//! `MatrixSum::{sum_col, sum_row, sum_col_squared}` and `MatrixNonZero::{nonzero_col, nonzero_row}` for `CsrMatrix<M>`
//! (src/sparse/csr.rs:23-122, 259-392, 558-608) and `CscMatrix<M>` (src/sparse/csc.rs:157-220, 323-335).  The `*_chunk`,
//! `*_masked` and `sum_row_squared` methods are not called by the PCA / preprocessing path and keep
//! the reference's host implementations in a real integration; they are omitted from this facade.
use crate::device::{ctx, DeviceCsr, SalgFloat};
use crate::ffi::*;
use nalgebra_sparse::{CscMatrix, CsrMatrix};
use num_traits::{Float, NumCast, PrimInt, Unsigned, Zero};
use std::ops::AddAssign;

pub trait MatrixNonZero {
    fn nonzero_col<T>(&self) -> anyhow::Result<Vec<T>> where T: PrimInt + Unsigned + Zero + AddAssign + Send + Sync;
    fn nonzero_row<T>(&self) -> anyhow::Result<Vec<T>> where T: PrimInt + Unsigned + Zero + AddAssign + Send + Sync;
}

pub trait MatrixSum {
    type Item: NumCast;
    fn sum_col<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync;
    fn sum_row<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync;
    fn sum_col_squared<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync;
}

fn cast_vec<A: NumCast + Copy, B: NumCast>(v: &[A]) -> anyhow::Result<Vec<B>> {
    v.iter().map(|&x| B::from(x).ok_or_else(|| anyhow::anyhow!("value exceeds target type capacity"))).collect()
}

impl<M: SalgFloat> MatrixNonZero for CsrMatrix<M> {
    fn nonzero_col<T>(&self) -> anyhow::Result<Vec<T>> where T: PrimInt + Unsigned + Zero + AddAssign + Send + Sync {
        if self.nnz() == 0 || self.ncols() == 0 { return Ok(vec![T::zero(); self.ncols()]); }     // csr.rs:33-35
        let dev = DeviceCsr::upload(self)?;
        let mut out = vec![0u64; self.ncols()];
        check(unsafe { salg_nonzero_col(ctx(), dev.raw(), out.as_mut_ptr()) })?;
        cast_vec(&out)
    }
    fn nonzero_row<T>(&self) -> anyhow::Result<Vec<T>> where T: PrimInt + Unsigned + Zero + AddAssign + Send + Sync {
        if self.nrows() == 0 { return Ok(Vec::new()); }                                             // csr.rs:87-89
        let dev = DeviceCsr::upload(self)?;
        let mut out = vec![0u64; self.nrows()];
        check(unsafe { salg_nonzero_row(ctx(), dev.raw(), out.as_mut_ptr()) })?;
        cast_vec(&out)       // "Count {} exceeds target type capacity" (csr.rs:104-106)
    }
}

impl<M: SalgFloat> MatrixSum for CsrMatrix<M> {
    type Item = M;
    fn sum_col<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        if self.nnz() == 0 || self.ncols() == 0 { return Ok(vec![T::zero(); self.ncols()]); }     // csr.rs:269-271
        let dev = DeviceCsr::upload(self)?;
        let mut out = vec![M::zero(); self.ncols()];
        check(unsafe { M::sum_col(dev.raw(), out.as_mut_ptr(), std::ptr::null_mut()) })?;
        cast_vec(&out)
    }
    fn sum_row<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        if self.nrows() == 0 { return Ok(Vec::new()); }                                             // csr.rs:323-325
        let dev = DeviceCsr::upload(self)?;
        let mut out = vec![M::zero(); self.nrows()];
        check(unsafe { M::sum_row(dev.raw(), out.as_mut_ptr()) })?;
        cast_vec(&out)
    }
    fn sum_col_squared<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        if self.nnz() == 0 || self.ncols() == 0 { return Ok(vec![T::zero(); self.ncols()]); }
        let dev = DeviceCsr::upload(self)?;
        let (mut s, mut q) = (vec![M::zero(); self.ncols()], vec![M::zero(); self.ncols()]);
        check(unsafe { M::sum_col(dev.raw(), s.as_mut_ptr(), q.as_mut_ptr()) })?;
        cast_vec(&q)
    }
}

impl<M: SalgFloat> MatrixSum for CscMatrix<M> {
    type Item = M;
    fn sum_col<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        let dev = DeviceCsr::upload_csc(self)?;
        let mut out = vec![M::zero(); self.ncols()];
        check(unsafe { M::csc_sum_col(dev.raw(), out.as_mut_ptr(), std::ptr::null_mut()) })?;
        cast_vec(&out)
    }
    fn sum_row<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        let dev = DeviceCsr::upload_csc(self)?;
        let mut out = vec![M::zero(); self.nrows()];
        check(unsafe { M::csc_sum_row(dev.raw(), out.as_mut_ptr()) })?;
        cast_vec(&out)
    }
    fn sum_col_squared<T>(&self) -> anyhow::Result<Vec<T>> where T: Float + NumCast + AddAssign + std::iter::Sum + Send + Sync {
        let dev = DeviceCsr::upload_csc(self)?;
        let mut q = vec![M::zero(); self.ncols()];
        check(unsafe { M::csc_sum_col(dev.raw(), std::ptr::null_mut(), q.as_mut_ptr()) })?;
        cast_vec(&q)
    }
}

//! `single_algebra::{Normalize, Log1P}` (src/utils/mod.rs:6-17) with FFI bodies for `CsrMatrix<T>`
//! (src/sparse/csr.rs:1013-1079) and `CscMatrix<T>` (src/sparse/csc.rs:680-746).  `&mut self` semantics: the host values
//! are refreshed from the device after the in-place kernel.
use crate::device::{ctx, DeviceCsr, SalgFloat};
use crate::ffi::*;
use nalgebra_sparse::{CscMatrix, CsrMatrix};
use num_traits::NumCast;
use single_utilities::traits::FloatOpsTS;
use single_utilities::types::Direction;

pub trait Normalize<T: FloatOpsTS> {
    fn normalize<U: FloatOpsTS>(&mut self, sums: &[U], target: U, direction: &Direction) -> anyhow::Result<()>;
}
pub trait Log1P<T: FloatOpsTS> {
    fn log1p_normalize(&mut self) -> anyhow::Result<()>;
}

fn dir(d: &Direction) -> std::os::raw::c_int { match d { Direction::ROW => SALG_ROW, Direction::COLUMN => SALG_COLUMN } }

/// U -> the value type when it is the same width, else f64 (`salg_normalize_f32_u64`: arithmetic in U, cast back to T,
/// as csr.rs:1042,1060 do)
fn sums_as<V: NumCast + Copy, U: NumCast + Copy>(sums: &[U]) -> Vec<V> { sums.iter().map(|&s| V::from(s).unwrap()).collect() }

impl<T: SalgFloat + FloatOpsTS> Normalize<T> for CsrMatrix<T> {
    fn normalize<U: FloatOpsTS>(&mut self, sums: &[U], target: U, direction: &Direction) -> anyhow::Result<()> {
        let dev = DeviceCsr::upload(self)?;
        if std::mem::size_of::<T>() == 4 && std::mem::size_of::<U>() == 8 {
            let s: Vec<f64> = sums_as(sums);
            check(unsafe { salg_normalize_f32_u64(ctx(), dev.raw(), s.as_ptr(), s.len() as i64, <f64 as NumCast>::from(target).unwrap(), dir(direction)) })?;
        } else {
            let s: Vec<T> = sums_as(sums);
            check(unsafe { T::normalize(dev.raw(), s.as_ptr(), s.len() as i64, <T as NumCast>::from(target).unwrap(), dir(direction)) })?;
        }
        dev.download_values_into(self.values_mut())
    }
}
impl<T: SalgFloat + FloatOpsTS> Log1P<T> for CsrMatrix<T> {
    fn log1p_normalize(&mut self) -> anyhow::Result<()> {
        let dev = DeviceCsr::upload(self)?;
        check(unsafe { salg_log1p(ctx(), dev.raw()) })?;                  // ln(fl(1 + v)), csr.rs:1074-1075
        dev.download_values_into(self.values_mut())
    }
}
impl<T: SalgFloat + FloatOpsTS> Normalize<T> for CscMatrix<T> {
    fn normalize<U: FloatOpsTS>(&mut self, sums: &[U], target: U, direction: &Direction) -> anyhow::Result<()> {
        let dev = DeviceCsr::upload_csc(self)?;
        if std::mem::size_of::<T>() == 4 && std::mem::size_of::<U>() == 8 {
            let s: Vec<f64> = sums_as(sums);
            check(unsafe { salg_csc_normalize_f32_u64(ctx(), dev.raw(), s.as_ptr(), s.len() as i64, <f64 as NumCast>::from(target).unwrap(), dir(direction)) })?;
        } else {
            let s: Vec<T> = sums_as(sums);
            check(unsafe { T::csc_normalize(dev.raw(), s.as_ptr(), s.len() as i64, <T as NumCast>::from(target).unwrap(), dir(direction)) })?;
        }
        dev.download_values_into(self.values_mut())
    }
}
impl<T: SalgFloat + FloatOpsTS> Log1P<T> for CscMatrix<T> {
    fn log1p_normalize(&mut self) -> anyhow::Result<()> {
        let dev = DeviceCsr::upload_csc(self)?;
        check(unsafe { salg_log1p(ctx(), dev.raw()) })?;                  // csc.rs:741-742
        dev.download_values_into(self.values_mut())
    }
}

/// Not in the reference: the whole `sum_row -> normalize(ROW, target) -> log1p -> (sum_col, sum_col_squared)` chain of
/// SURVEY §3.4 in one call on one upload (`salg_preprocess_*`); returns the column sums and sums of squares.
pub fn preprocess<T: SalgFloat + FloatOpsTS>(x: &mut CsrMatrix<T>, target: T) -> anyhow::Result<(Vec<T>, Vec<T>)> {
    let dev = DeviceCsr::upload(x)?;
    let (mut s, mut q) = (vec![T::zero(); x.ncols()], vec![T::zero(); x.ncols()]);
    check(unsafe { T::preprocess(dev.raw(), target, s.as_mut_ptr(), q.as_mut_ptr()) })?;
    dev.download_values_into(x.values_mut())?;
    Ok((s, q))
}

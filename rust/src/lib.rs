//! `single_algebra_b200` — the sparse-PCA hot path of SingleRust/single-algebra 0.9.2 behind the reference's own type
//! surface, every body an FFI call into `libsalg_b200.so` (hand-written sm_100a CUDA; C ABI in include/salg.h).
//!
//! Module paths mirror the reference (src/lib.rs:43-51): `sparse::{MatrixSum, MatrixNonZero}`, `{Normalize, Log1P}`,
//! `dimred::pca::{SparsePCA, SparsePCABuilder, MaskedSparsePCA, MaskedSparsePCABuilder, SVDMethod,
//! PowerIterationNormalizer}`.  UNVERIFIED SOURCE (no Rust toolchain in the build image) — see rust/Cargo.toml.
pub mod device;
pub mod dimred;
pub mod ffi;
pub mod sparse;
pub mod utils;

pub use utils::{Log1P, Normalize};

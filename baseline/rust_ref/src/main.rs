//! Times `MaskedSparsePCA::fit_transform` / `SparsePCA::fit_transform` of the unmodified reference crate on a CSR read from
//! raw little-endian files: `DIR/shape.txt` ("nrows ncols nnz n_mask"), `DIR/indptr.i64`, `DIR/indices.i32`, `DIR/data.f32`,
//! `DIR/mask.u8` (optional).  UNVERIFIED (never compiled in the build image).  The reference draws its own Omega from the
//! seed (ChaCha12 + ziggurat, not reproducible outside Rust), so this harness gives TIMING, not entry-wise parity.
use anyhow::Result;
use nalgebra_sparse::CsrMatrix;
use single_algebra::dimred::pca::{MaskedSparsePCABuilder, PowerIterationNormalizer, SVDMethod, SparsePCABuilder};
use std::{fs, time::Instant};

fn read<T: Copy>(path: &str) -> Result<Vec<T>> {
    let b = fs::read(path)?;
    let n = b.len() / std::mem::size_of::<T>();
    let mut v = Vec::<T>::with_capacity(n);
    unsafe { std::ptr::copy_nonoverlapping(b.as_ptr(), v.as_mut_ptr() as *mut u8, n * std::mem::size_of::<T>()); v.set_len(n); }
    Ok(v)
}

fn main() -> Result<()> {
    let dir = std::env::args().nth(1).expect("usage: salg_rust_ref DIR [steps]");
    let steps: usize = std::env::args().nth(2).and_then(|s| s.parse().ok()).unwrap_or(3);
    let shape: Vec<usize> = fs::read_to_string(format!("{dir}/shape.txt"))?.split_whitespace().map(|s| s.parse().unwrap()).collect();
    let (nrows, ncols, n_mask) = (shape[0], shape[1], shape[3]);
    let indptr: Vec<usize> = read::<i64>(&format!("{dir}/indptr.i64"))?.into_iter().map(|x| x as usize).collect();
    let indices: Vec<usize> = read::<i32>(&format!("{dir}/indices.i32"))?.into_iter().map(|x| x as usize).collect();
    let data: Vec<f32> = read::<f32>(&format!("{dir}/data.f32"))?;
    let x = CsrMatrix::try_from_csr_data(nrows, ncols, indptr, indices, data).map_err(|e| anyhow::anyhow!("{e}"))?;
    let method = SVDMethod::Random { n_oversamples: 10, n_power_iterations: 7, normalizer: PowerIterationNormalizer::QR };
    println!("threads {}", rayon::current_num_threads());
    for s in 0..steps {
        let t = Instant::now();
        if n_mask > 0 {
            let mask: Vec<bool> = read::<u8>(&format!("{dir}/mask.u8"))?.into_iter().map(|b| b != 0).collect();
            let mut pca = MaskedSparsePCABuilder::<f32>::new().n_components(50).mask(mask).svd_method(method).build();
            let scores = pca.fit_transform(&x)?;
            println!("step {s}: {:.3} s, scores {:?}", t.elapsed().as_secs_f64(), scores.dim());
        } else {
            // NOTE: SparsePCA::transform is O(rows * k * nnz_total * log) in the reference: time `fit` only
            let mut pca = SparsePCABuilder::<f32>::new().n_components(50).svd_method(method).build();
            pca.fit(&x)?;
            println!("step {s}: fit {:.3} s", t.elapsed().as_secs_f64());
        }
    }
    Ok(())
}

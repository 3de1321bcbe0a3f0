"""Run under torchrun with N >= 2 GPUs: row-sharded fit over NCCL must equal the oracle on the whole
matrix (same Omega) — randomized (masked and unmasked, f32/f64), Lanczos, column statistics, operators."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    import scipy.sparse as sp
    import single_algebra_b200 as s
    from conftest import planted_counts
    from oracle import oracle as O
    ctx = s.dist.init_context_from_env()
    s.set_default_context(ctx)
    for dtype, stol in ((np.float64, 1e-5), (np.float32, 1e-4)):
        A = planted_counts(5000, 900, seed=41, dtype=dtype)
        A = sp.vstack([A, sp.csr_matrix((7, 900), dtype=dtype)]).tocsr()
        r0, r1 = s.dist.partition_rows_by_nnz(A.indptr, world)[rank]
        off, idx, val = s.dist.shard_csr(A.indptr, A.indices, A.data, r0, r1)
        x = s.CsrMatrix(r1 - r0, 900, off, idx.astype(np.uint64), val, ctx)
        d = x.to_device()
        # column statistics are global
        sm, sq = d.sum_col_and_squared()
        assert np.allclose(sm, O.sum_col(A.indptr, A.indices, A.data, 900), rtol=1e-5)
        assert np.allclose(sq, O.sum_col_squared(A.indptr, A.indices, A.data, 900), rtol=1e-5)
        # transposed product is all-reduced, with the centring applied once
        mu = (sm / A.shape[0]).astype(dtype)
        Yfull = np.random.default_rng(1).standard_normal((A.shape[0], 60)).astype(dtype)
        Z = s.op_spmm(d, Yfull[r0:r1], mu=mu, transposed=True)
        Zr = A.T.astype(np.float64) @ Yfull - mu.astype(np.float64)[:, None] * Yfull.sum(axis=0)[None, :]
        assert np.abs(Z - Zr).max() < (1e-10 if dtype == np.float64 else 3e-5) * np.abs(Zr).max()
        for mask in (None, s.synth.make_mask(900, 250, seed=7)):
            n_eff = 900 if mask is None else 250
            om = s.synth.make_omega(n_eff, 40, seed=42, dtype=dtype)
            b = s.SparsePCABuilder() if mask is None else s.MaskedSparsePCABuilder().mask(mask.tolist())
            p = b.n_components(30).svd_method(s.SVDMethod.Random(10, 7, s.PowerIterationNormalizer.QR)).build()
            sc = p.fit_transform(x, omega=om)
            ref = O.sparse_pca_fit(A.astype(np.float64), 30, omega=om.astype(np.float64), mask=mask)
            serr, ang = O.rel_err(p.singular_values_, ref.singular_values), O.largest_principal_angle(p.components_, ref.components)
            if rank == 0:
                print(f"[dist check] world {world} {np.dtype(dtype).name} {'unmasked' if mask is None else 'masked'} randomized fit: "
                      f"sigma rel err {serr:.2e} (< {stol:.0e}), largest principal angle {ang:.2e} rad (< 1e-3)", flush=True)
            assert serr < stol, (dtype, mask is None)
            assert ang < 1e-3
            assert np.allclose(p.mean_, ref.mean, rtol=1e-4, atol=1e-12)
            ex = O.transform(A, p.components_, p.mean_, mask=mask, mode=O.EXACT)[r0:r1]
            assert np.abs(sc - ex).max() < (1e-9 if dtype == np.float64 else 3e-4) * np.abs(ex).max()
        # Lanczos, rows sharded
        p = s.SparsePCABuilder().n_components(15).build()
        p.fit(x)
        u, sv, vt = O.truncated_svd_truth(A.astype(np.float64), 15)
        if rank == 0:
            print(f"[dist check] world {world} {np.dtype(dtype).name} Lanczos: sigma rel err {O.rel_err(p.singular_values_, sv):.2e}, "
                  f"angle {O.largest_principal_angle(p.components_, vt):.2e} rad", flush=True)
        assert O.rel_err(p.singular_values_, sv) < stol
        assert O.largest_principal_angle(p.components_, vt) < 1e-3
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_CHECK_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 (gloo, CPU) check of the row-sharded schedule the library runs over NCCL:
each rank holds a row block; column statistics, the Gram matrices of the tall panels and every
n_eff-sized panel are all-reduced; the result must equal the single-process oracle.  Also covers the
plumbing that hands rank 0's NCCL unique id to the other ranks."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
    dist.all_reduce(t)
    return t.numpy()


def _cholqr2_sharded(Y):
    for _ in range(2):
        G = _allreduce(Y.T @ Y)
        R = np.linalg.cholesky(G).T
        Y = np.linalg.solve(R.T, Y.T).T
    return Y


def _sharded_rsvd(A_loc, n_total, k, p, q, omega):
    """The device schedule in numpy: local A X, all-reduced A^T Y with the rank-1 centring applied once."""
    world = dist.get_world_size()
    mu = _allreduce(np.asarray(A_loc.sum(axis=0)).ravel()) / n_total
    At = A_loc.T.tocsr()

    def AtY(Y):
        cs = _allreduce(Y.sum(axis=0))                       # global 1^T Y
        part = At @ Y - mu[:, None] * cs[None, :]            # every rank subtracts the full correction ...
        Z = _allreduce(part)
        return Z + (world - 1) * mu[:, None] * cs[None, :]   # ... so add it back world-1 times

    Y = A_loc @ omega - (mu @ omega)[None, :]
    for _ in range(q):
        Y = _cholqr2_sharded(Y)
        Z = AtY(Y)
        Z, _ = np.linalg.qr(Z)
        Y = A_loc @ Z - (mu @ Z)[None, :]
    Q = _cholqr2_sharded(Y)
    Bt = AtY(Q)
    _, s, vt = np.linalg.svd(Bt.T, full_matrices=False)
    return s[:k], vt[:k]


def _sharded_rsvd_implicit(A_loc, n_total, k, p, q, omega):
    """The f32 tensor-core path's schedule (pca.cu, `fused`): the tall panel is never normalised explicitly — one pass
    gives its Gram matrix and column sums, every rank centres its partial A^T Y with ITS OWN column sums, Gram and
    partial panel share one all-reduce, and R^{-1} multiplies the small side: A_c^T (Y R^{-1}) = (A_c^T Y) R^{-1}."""
    mu = _allreduce(np.asarray(A_loc.sum(axis=0)).ravel()) / n_total
    At = A_loc.T.tocsr()
    ncols = A_loc.shape[1]

    def half_step(Y):
        G = Y.T @ Y
        cs = Y.sum(axis=0)
        part = At @ Y - mu[:, None] * cs[None, :]            # local centring term only
        buf = _allreduce(np.concatenate([G.ravel(), part.ravel()]))      # ONE collective
        l = Y.shape[1]
        G, Zp = buf[:l * l].reshape(l, l), buf[l * l:].reshape(ncols, l)
        R = np.linalg.cholesky(G).T
        return np.linalg.solve(R.T, Zp.T).T, R              # Z' R^{-1}

    Y = A_loc @ omega - (mu @ omega)[None, :]
    for _ in range(q):
        Z, _ = half_step(Y)
        Z, _ = np.linalg.qr(Z)
        Y = A_loc @ Z - (mu @ Z)[None, :]
    # final CholeskyQR2: Q1 = Y R1^{-1} explicitly, second Gram of Q1, R2^{-1} on the small side
    _, R1 = half_step(Y)
    Q1 = np.linalg.solve(R1.T, Y.T).T
    Bt, _ = half_step(Q1)
    _, s, vt = np.linalg.svd(Bt.T, full_matrices=False)
    return s[:k], vt[:k]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import single_algebra_b200 as s
    from conftest import planted_counts
    from oracle import oracle as O
    try:
        uid = s.dist.broadcast_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128))
        A = planted_counts(900, 120, seed=3)
        parts = s.dist.partition_rows_by_nnz(A.indptr, world)
        r0, r1 = parts[rank]
        off, idx, val = s.dist.shard_csr(A.indptr, A.indices, A.data, r0, r1)
        import scipy.sparse as sp
        A_loc = sp.csr_matrix((val, idx, off), shape=(r1 - r0, 120))
        om = np.random.default_rng(0).standard_normal((120, 25))
        sv, vt = _sharded_rsvd(A_loc, 900, 15, 10, 5, om)
        u, s_ref, vt_ref = O.randomized_svd(A, 15, 10, 5, om, mean_center=True)
        assert O.rel_err(sv, s_ref) < 1e-9
        assert O.largest_principal_angle(vt, vt_ref) < 1e-6
        sv2, vt2 = _sharded_rsvd_implicit(A_loc, 900, 15, 10, 5, om)
        assert O.rel_err(sv2, s_ref) < 1e-9
        assert O.largest_principal_angle(vt2, vt_ref) < 1e-6
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_row_sharded_schedule_matches_oracle_world2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): BASELINE configs 4 and 5 on one B200.

  cfg4: SparsePCA f64, SVDMethod::Lanczos, 250k x 20k CSR @7 %, k = 50 (SpMV + full reorthogonalisation)
  cfg5: one GPU's share of the 4M x 33k f32 matrix (500k x 33k @7 %): sum_row, normalize(ROW), log1p,
        fused preprocess, sum_col / sum_col_squared — achieved GB/s against the algorithmic bytes of SURVEY §8d.
Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import single_algebra_b200 as s  # noqa: E402

PEAK = 6451.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(ctx, fn, reps=3):
    fn()
    ctx.prof_reset()
    ctx.prof_enable(True)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop() / reps
    ctx.prof_enable(False)
    return ms, {k: (v[0] / reps, v[1] // reps, v[2] / reps) for k, v in ctx.prof().items()}


def cfg5(ctx):
    nrows, ncols = 500_000, 33_000
    spec = s.synth.make_spec(4_000_000, ncols, density=0.07, seed=42)
    d = s.synth_device(spec, 0, nrows, dtype=np.float32, ctx=ctx)
    nnz, S, I, O = d.nnz, 4, 4, 8
    out = []

    def rec(name, ms, byt):
        out.append({"config": "cfg5 shard 500k x 33k f32", "op": name, "ms": ms, "algorithmic_bytes": byt,
                    "achieved_gbs": byt / ms / 1e6, "frac_of_measured_hbm": byt / ms / 1e6 / PEAK})

    ms, pr = timed(ctx, lambda: d.sum_row())
    rec("sum_row (kernel)", pr["stats"][0], nnz * S + (nrows + 1) * O + nrows * S)
    out[-1]["ms_incl_d2h_of_2MB"] = ms
    ms, pr = timed(ctx, lambda: d.sum_col_and_squared())
    rec("sum_col + sum_col_squared, one pass (kernel)", pr["stats"][0], nnz * (S + I) + 2 * ncols * S)
    rs = d.sum_row()
    ms, pr = timed(ctx, lambda: d.normalize(rs, 1e4, s.Direction.ROW), reps=1)
    rec("normalize ROW (kernel)", pr["elementwise"][0], 2 * nnz * S + (nrows + 1) * O + nrows * S)
    ms, pr = timed(ctx, lambda: d.log1p_normalize(), reps=1)
    rec("log1p (kernel)", pr["elementwise"][0], 2 * nnz * S)
    d.free()
    d = s.synth_device(spec, 0, nrows, dtype=np.float32, ctx=ctx)
    ctx.prof_reset(); ctx.prof_enable(True)
    d.preprocess(1e4)
    ctx.prof_enable(False)
    pr = ctx.prof()
    rec("fused preprocess: row pass (sum_row+normalize+log1p)", pr["elementwise"][0], 2 * nnz * S + (nrows + 1) * O)
    rec("fused preprocess: column statistics pass", pr["stats"][0], nnz * (S + I) + 2 * ncols * S)
    # parity spot check of the fused chain on the first 64 rows against the oracle on the host-regenerated rows
    from oracle import oracle as O_
    ip, ix, dv = s.synth.generate_rows(spec, 0, 64, dtype=np.float32)
    v = O_.normalize(ip, ix, dv, O_.sum_row(ip, ix, dv, 64), np.float32(1e4), O_.ROW)
    v = O_.log1p_normalize(v)
    got = d.download_values()[:len(v)]
    out.append({"config": "cfg5 shard", "op": "parity of fused preprocess on 64 host-regenerated rows",
                "max_rel_err": float(np.max(np.abs(got - v) / np.maximum(np.abs(v), 1e-30)))})
    d.free()
    return out


def cfg4(ctx):
    nrows, ncols = 250_000, 20_000
    spec = s.synth.make_spec(nrows, ncols, density=0.07, seed=42)
    d = s.synth_device(spec, dtype=np.float64, ctx=ctx)
    pca = s.SparsePCABuilder().n_components(50).svd_method(s.SVDMethod.Lanczos).build()
    ctx.sync()
    t0 = time.perf_counter()
    pca.fit(d)                 # cold: the transposed copy's buffers and the Krylov bases grow the memory pool
    dt_cold = time.perf_counter() - t0
    ctx.prof_reset(); ctx.prof_enable(True)
    ctx.sync()
    t0 = time.perf_counter()
    pca.fit(d)
    dt = time.perf_counter() - t0
    ctx.prof_enable(False)
    pr = ctx.prof()
    sp = pr.get("spmv")
    res = {"config": "cfg4 SparsePCA f64 Lanczos 250k x 20k @7%, k=50", "fit_seconds": dt, "first_fit_seconds_cold_pool": dt_cold,
           "nnz": d.nnz,
           "numeric_flags": pca.numeric_flags()}
    if sp:
        res.update({"spmv_launches": sp[1], "spmv_ms_avg": sp[0] / sp[1], "spmv_achieved_gbs": sp[2] / sp[0] / 1e6,
                    "spmv_frac_of_measured_hbm": sp[2] / sp[0] / 1e6 / PEAK, "lanczos_steps": sp[1] // 2})
    # size-independent checks: orthonormal components, sigma_i = ||A v_i|| (uncentred operator), descending
    V = pca.components_
    res["orthonormality_err"] = float(np.abs(V @ V.T - np.eye(V.shape[0])).max())
    AV = s.op_spmm(d, np.ascontiguousarray(V.T))
    res["sigma_vs_norm_Av_rel_err"] = float(np.max(np.abs(np.linalg.norm(AV, axis=0) - pca.singular_values_) / pca.singular_values_))
    # residual of the eigen-relation A^T A v = sigma^2 v
    W = s.op_spmm(d, AV, transposed=True)
    res["max_residual_AtAv_rel"] = float(np.max(np.linalg.norm(W - V.T * pca.singular_values_ ** 2, axis=0) / pca.singular_values_ ** 2))
    res["sigma_top5"] = [float(x) for x in pca.singular_values_[:5]]
    d.free()
    return [res]


if __name__ == "__main__":
    ctx = s.default_context()
    which = sys.argv[1:] or ["cfg5", "cfg4"]
    for w in which:
        for r in (cfg5(ctx) if w == "cfg5" else cfg4(ctx)):
            print(json.dumps(r), flush=True)

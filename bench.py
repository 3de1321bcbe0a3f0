#!/usr/bin/env python
"""bench.py — randomized sparse-PCA fit_transform on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/cpu_ref.cpp) on the host cores
  torchrun ... bench.py --gpus N ...                       # rows sharded over N GPUs

A step is one MaskedSparsePCA / SparsePCA fit_transform (column statistics, mask compaction, tile format,
16 centred sparse x panel products, CholeskyQR normalisers, Jacobi SVD, svd_flip, projection) over one synthetic count
matrix.  `value` = cells/s with the CSR resident in HBM; `e2e` = the same through the public API from
pinned HOST buffers (upload + validation inside the timed region, scores copied back).
Workloads (BASELINE.json configs):
  cfg3 (default) MaskedSparsePCA f32 1M x 30k @7 %, 2000-gene mask — the config the metric's "1/2/4/8 B200" and the
                 north-star target are quoted on; strong scaling (rows split over the ranks, balanced by entries)
  cfg2           SparsePCA f32 100k x 20k @7 %; strong scaling
  cfg5           SparsePCA f32 4M x 33k @7 % row-sharded over 8 GPUs + normalize(ROW, 1e4) + log1p + sum_col /
                 sum_col_squared preprocessing; weak scaling: 500k rows per GPU (N = 8 is the named 4M x 33k matrix)
Inputs are far larger than L2 (126 MB), so no explicit L2 flush is needed between steps.
The reference arm runs the SAME config (full matrix, host-generated, all host cores) through oracle/cpu_ref.cpp — the
multi-threaded C++ restatement of the reference's Rust path (the crate cannot be built here: no Rust toolchain, and its SVD
engine single-svdlib is not on disk).  It loads no file of the CUDA library.
"""
import argparse
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(name="MaskedSparsePCA f32 1Mx30k CSR @7% nnz, 2000-gene mask, Random{p=10,q=7,QR}, k=50, center",
                 nrows=1_000_000, ncols=30_000, density=0.07, n_mask=2_000, k=50, p=10, q=7, scaling="strong"),
    "cfg2": dict(name="SparsePCA f32 100kx20k CSR @7% nnz, Random{p=10,q=7,QR}, k=50, center",
                 nrows=100_000, ncols=20_000, density=0.07, n_mask=0, k=50, p=10, q=7, scaling="strong"),
    "cfg5": dict(name="SparsePCA f32 4Mx33k CSR @7% nnz row-sharded over 8 GPUs (500k-row shard per GPU), "
                      "preprocess normalize(ROW,1e4)+log1p+sum_col/sum_col_squared, then Random{p=10,q=7,QR}, k=50, center",
                 nrows=500_000, ncols=33_000, density=0.07, n_mask=0, k=50, p=10, q=7, scaling="weak", preprocess=1e4,
                 spec_rows=4_000_000, cpu_rows=100_000),
    "tiny": dict(name="MaskedSparsePCA f32 20kx3k CSR @7% nnz, 500-gene mask (self-test)",
                 nrows=20_000, ncols=3_000, density=0.07, n_mask=500, k=50, p=10, q=7, scaling="strong"),
}


def total_rows(wl, world):
    return wl["nrows"] * world if wl["scaling"] == "weak" else wl["nrows"]


def config_of(wl, world):
    """The workload description — identical in both arms (the arm-specific facts go to `details`)."""
    return {"workload": wl["name"], "rows": total_rows(wl, world), "cols": wl["ncols"], "density": wl["density"],
            "mask_genes": wl["n_mask"], "n_components": wl["k"], "n_oversamples": wl["p"], "n_power_iterations": wl["q"],
            "l2": "inputs larger than L2; no flush needed",
            "omega": "host-generated PCG64(42) standard normal, same on both arms"}


def load_synth_module():
    """single-algebra_b200/synth.py by path: pure numpy, so the reference arm maps no file of the CUDA library."""
    spec = importlib.util.spec_from_file_location("_salg_synth", os.path.join(ROOT, "single-algebra_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_salg_synth"] = mod      # dataclasses resolves the module through sys.modules
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 200 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok or os.environ.get("SALG_BENCH_NO_CLOCKS"):     # (experiment switch: is NVML perturbing the timing?)
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def pinned_empty(n, dtype):
    """numpy view of a pinned host buffer (torch is plumbing: pinned memory)."""
    import torch
    t = torch.empty(int(n), dtype={np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64,
                                   np.uint8: torch.uint8}[dtype], pin_memory=True)
    return t, t.numpy()


def host_mem_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def principal_angle(V1, V2):
    """largest principal angle (rad) between the row spaces of two d x n matrices (sin-based, accurate when small)"""
    q1, _ = np.linalg.qr(np.asarray(V1, dtype=np.float64).T)
    q2, _ = np.linalg.qr(np.asarray(V2, dtype=np.float64).T)
    r = q2 - q1 @ (q1.T @ q2)
    return float(np.arcsin(min(1.0, np.linalg.svd(r, compute_uv=False).max())))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ------------------------------------------------------------------------------------------------------
def cpu_fit_transform(R, wl, ptr, idx, val, nrows, mask, om, preprocess_src=None):
    """One step of the workload on the host cores through oracle/cpu_ref.cpp; returns (result, seconds)."""
    t0 = time.perf_counter()
    if wl.get("preprocess"):
        np.copyto(val, preprocess_src)            # the step starts from the raw counts, like the GPU step
        R.preprocess_f32(ptr, idx, val, nrows, wl["ncols"], wl["preprocess"])
    r = R.pca_fit(ptr, idx, val, nrows, wl["ncols"], wl["k"], om, mask=mask, n_oversamples=wl["p"],
                  n_power_iterations=wl["q"], center=True, want_scores=True)
    return r, time.perf_counter() - t0


def run_reference(args, wl):
    """The reference arm: the reference's CPU algorithm (oracle/cpu_ref.cpp — multi-threaded C++ restatement; the Rust
    crate cannot be built in this image) on ALL host cores, on the same config as the CUDA arm.  The input comes from the
    host generator in the same file, so no file of the CUDA library is mapped."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_ref as R
    synth = load_synth_module()
    world = args.gpus
    threads = R.set_threads()        # explicit: torchrun exports OMP_NUM_THREADS=1
    rows_cfg = total_rows(wl, world)
    rows = min(rows_cfg, wl.get("cpu_rows", rows_cfg))
    need_gb = rows * wl["ncols"] * wl["density"] * 8 * (2.2 if wl.get("preprocess") else 1.2) / 1e9 + 4
    while host_mem_gb() < need_gb and rows > 20_000:       # bounded by host memory: say so in `sample`
        rows //= 2
        need_gb = rows * wl["ncols"] * wl["density"] * 8 * 1.2 / 1e9 + 4
    spec = synth.make_spec(wl.get("spec_rows", wl["nrows"]), wl["ncols"], density=wl["density"], seed=42)
    t0 = time.perf_counter()
    ptr, idx, val = R.synth_rows(spec, 0, rows)
    t_gen = time.perf_counter() - t0
    mask = synth.make_mask(wl["ncols"], wl["n_mask"], seed=7) if wl["n_mask"] else None
    n_eff = wl["n_mask"] or wl["ncols"]
    om = synth.make_omega(n_eff, wl["k"] + wl["p"], seed=42, dtype=np.float32)
    raw = val.copy() if wl.get("preprocess") else None

    def step():
        return cpu_fit_transform(R, wl, ptr, idx, val, rows, mask, om, raw)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last, _dt = step()
    dt = (time.perf_counter() - t0) / args.steps
    value = rows / dt
    sample = (f"the full {rows_cfg} x {wl['ncols']} matrix" if rows == rows_cfg
              else f"first {rows} of {rows_cfg} rows (all {wl['ncols']} columns)")
    line = {
        "impl": "reference", "metric": "randomized_pca_fit_transform_throughput", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "seconds_per_fit_transform": dt, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(wl, world),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port",
                         "sample": f"{sample}, host-generated in {t_gen:.1f} s; oracle/cpu_ref.cpp (C++17 + OpenMP restatement "
                                   "of the reference path: 3 + 1 statistics passes, 16 row-parallel products testing the mask "
                                   "per stored entry, 15 Householder QRs parallel over rows, QR + Jacobi SVD of B, transform)",
                         "phase_seconds_last_step": last.timings if last is not None else None},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "details": {"rows_timed": rows, "host_threads": threads,
                    "sigma_head": [float(x) for x in last.singular_values[:3]] if last is not None else None},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import single_algebra_b200 as s

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa_cpus = 0
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl")
        if not os.environ.get("SALG_NO_NUMA_BIND"):
            numa_cpus = s.dist.bind_to_gpu_numa_node(local)     # before any pinned allocation (first touch)
    if s.device_count() == 0:
        raise RuntimeError("bench.py needs a CUDA device: libsalg_b200 has no CPU fallback")
    ctx = s.dist.init_context_from_env()
    s.set_default_context(ctx)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    # ---- inputs: this rank's row block, generated on the device ------------------------------------
    weak = wl["scaling"] == "weak"
    n_total = total_rows(wl, world)
    spec = s.synth.make_spec(wl.get("spec_rows", wl["nrows"]), wl["ncols"], density=wl["density"], seed=42)
    if weak:
        r0, r1 = rank * wl["nrows"], (rank + 1) * wl["nrows"]
        shard_rule = "fixed 500k-row shard per GPU"
    elif world > 1:
        # contiguous row blocks balanced by (expected) stored entries, SURVEY §8e
        exp = s.synth.expected_row_nnz(spec, 0, n_total)
        r0, r1 = s.dist.partition_rows_by_nnz(np.concatenate([[0.0], np.cumsum(exp)]), world)[rank]
        shard_rule = "contiguous row blocks balanced by expected stored entries (partition_rows_by_nnz)"
    else:
        r0, r1 = 0, n_total
        shard_rule = "one shard"
    dev = s.synth_device(spec, r0, r1 - r0, dtype=np.float32, ctx=ctx)
    mask = s.synth.make_mask(wl["ncols"], wl["n_mask"], seed=7) if wl["n_mask"] else None
    n_eff = wl["n_mask"] or wl["ncols"]
    om = s.synth.make_omega(n_eff, wl["k"] + wl["p"], seed=42, dtype=np.float32)
    raw_vals = dev.clone_values() if wl.get("preprocess") else None      # device copy of the raw counts

    def make_pca():
        b = s.MaskedSparsePCABuilder().mask(mask.tolist()) if mask is not None else s.SparsePCABuilder()
        return b.n_components(wl["k"]).svd_method(
            s.SVDMethod.Random(wl["p"], wl["q"], s.PowerIterationNormalizer.QR)).build()

    pca = make_pca()

    def step_resident(fetch=False):
        if raw_vals is not None:
            dev.restore_values(raw_vals)              # the step starts from raw counts (device-to-device, inside the timing)
            dev.preprocess_device(wl["preprocess"])   # normalize ROW + log1p + column sums, results stay on the device
        pca._fit(dev, om, keep_scores=True, fetch=fetch)

    # ---- device-resident timing -----------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    ctx.prof_reset()
    # inside the timed region only the product / preprocessing classes are event-timed (the roofline's live
    # measurement); the full per-class table comes from two extra, untimed, fully profiled fits right after
    ctx.prof_enable(True, products_only=not os.environ.get("SALG_BENCH_PROF_ALL"))
    launches0 = ctx.launch_count()
    barrier()
    ctx.sync()
    sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        _t0 = time.perf_counter()
        step_resident()
        if os.environ.get("SALG_BENCH_VERBOSE") and rank == 0:
            ctx.sync()
            print(f"[resident] step {1e3 * (time.perf_counter() - _t0):.1f} ms", file=sys.stderr, flush=True)
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.finish()
    ctx.prof_enable(False)
    launches = ctx.launch_count() - launches0
    prof = ctx.prof()
    ms_step = max_over_ranks(ms) / args.steps
    ctx.prof_reset()
    ctx.prof_enable(True)
    n_prof_steps = 2
    for _ in range(n_prof_steps):
        step_resident()
    ctx.sync()
    ctx.prof_enable(False)
    prof_all = ctx.prof()
    value = n_total / (ms_step * 1e-3)

    # ---- multi-rank parity: the N-rank fit against a single-GPU fit of the whole matrix (rank 0, untimed) ----------
    parity_n1 = None
    if world > 1 and not weak and not args.no_parity:
        step_resident(fetch=True)
        if rank == 0:
            ctx1 = s.Context(local)
            full = s.synth_device(spec, 0, n_total, dtype=np.float32, ctx=ctx1)
            p1 = make_pca()
            p1._fit(full, om, keep_scores=False, fetch=True)
            parity_n1 = {"sigma_rel_err": rel_err(pca.singular_values_, p1.singular_values_),
                         "largest_principal_angle_rad": principal_angle(pca.components_, p1.components_),
                         "mean_max_abs_diff": float(np.abs(pca.mean_.astype(np.float64) - p1.mean_).max()),
                         "against": "single-GPU fit of the whole matrix on rank 0, same Omega (untimed)"}
            p1._free_model()
            full.free()
            ctx1.close()
        barrier()

    # ---- roofline of the dominant kernel (live CUDA events around every launch of the class) ------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # per-class table: the event-timed classes from the timed region, the others from the untimed profiled fits scaled to
    # the same number of steps (so ms_total / steps is per fit for every class)
    scale = args.steps / n_prof_steps
    merged = {k: (v[0] * scale, int(round(v[1] * scale)), v[2] * scale) for k, v in prof_all.items()}
    merged.update(prof)
    classes = {k: {"ms_total": v[0], "launches": v[1], "gbs": (v[2] / v[0] / 1e6) if v[0] > 0 and v[2] > 0 else None,
                   "frac_of_hbm_peak": (v[2] / v[0] / 1e6 / peak) if v[0] > 0 and v[2] > 0 else None}
               for k, v in merged.items()}
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two product kernels from ONE
    # `ncu --set full` capture each, on this workload at 1 GPU (profiles/, see TRAFFIC_SRC)
    TRAFFIC_SRC = "profiles/r02c_ncu_full_top_kernels.csv"
    ncu_traffic = {("cfg3", "spmm"): 1.094885e9 + 0.236818e9, ("cfg3", "spmm_t"): 1.238378e9 + 0.006522e9}
    dom = max((k for k in ("spmm", "spmm_t") if k in prof), key=lambda k: prof[k][0], default=None)
    roofline = None
    if dom:
        tms, n, b = prof[dom]
        achieved = b / tms / 1e6     # GB/s
        roofline = {"bound": "hbm", "kernel": ("tm_product_kernel<true> (A^T Y, tcgen05.mma with the sparse operand expanded into TMEM)" if dom == "spmm_t"
                               else "tm_product_kernel<false> (A X, tcgen05.mma with the sparse operand expanded into TMEM)"),
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic.get((args.workload, dom)) if world == 1 else None,
                    "traffic_source": TRAFFIC_SRC if world == 1 and (args.workload, dom) in ncu_traffic else None,
                    "peak_source": peak_src, "avg_launch_ms": tms / n, "launches": n,
                    "algorithmic_bytes_per_launch": b / n, "share_of_step": tms / (ms * 1.0)}

    # ---- end to end through the public API from pinned host buffers --------------------------------------
    e2e = None
    host = None
    need_gb = dev.nnz * 8 / 1e9 * 1.3 + 2
    if not args.no_e2e and host_mem_gb() > need_gb:
        nnz = dev.nnz
        t_off, off = pinned_empty(dev.nrows + 1, np.int64)
        t_idx, idx = pinned_empty(nnz, np.int32)
        t_val, val = pinned_empty(nnz, np.float32)
        if raw_vals is not None:
            dev.restore_values(raw_vals)
        dev.download_raw(off, idx.view(np.uint32), val)
        host = (off, idx, val)
        nloc = dev.nrows
        t_sc, sc = pinned_empty(nloc * wl["k"], np.float32)
        if raw_vals is not None:
            dev.free_values_clone(raw_vals)
            raw_vals = None
        dev.free()   # the e2e leg owns device memory from here on

        verbose = bool(os.environ.get("SALG_BENCH_VERBOSE"))

        def step_e2e():
            t0 = time.perf_counter()
            x = s.CsrMatrix(nloc, wl["ncols"], off.view(np.uint64), idx, val, ctx)
            p2 = make_pca()
            if wl.get("preprocess"):
                x.to_device()
                x._dev.preprocess_device(wl["preprocess"])
            out = p2.fit_transform(x, omega=om, out=sc.reshape(nloc, -1))   # scores land in pinned host memory
            t2 = time.perf_counter()
            x.drop_device()
            p2._free_model()
            t3 = time.perf_counter()
            if verbose and rank == 0:
                print(f"[e2e] upload + fit_transform + fetch {1e3*(t2-t0):.1f} ms, free {1e3*(t3-t2):.1f} ms",
                      file=sys.stderr, flush=True)
            return out

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ctx.prof_reset()
        ctx.prof_enable(True)
        barrier()
        ctx.sync()
        ctx.timer_start()
        k_e2e = max(1, min(args.steps, 3))
        for _ in range(k_e2e):
            step_e2e()
        ms2 = ctx.timer_stop()
        barrier()
        ctx.prof_enable(False)
        e2e_classes = {k: round(v[0] / k_e2e, 3) for k, v in ctx.prof().items()}
        ms2_step = max_over_ranks(ms2) / k_e2e
        h2d = sum_over_ranks(nnz * 8 + (nloc + 1) * 8 + om.nbytes + (wl["ncols"] if mask is not None else 0))
        d2h = sum_over_ranks(nloc * wl["k"] * 4 + 8 * wl["ncols"] + 16 * wl["k"])
        e2e = {"value": n_total / (ms2_step * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2_step, "steps": k_e2e,
               "class_ms_per_step": e2e_classes}
    elif not args.no_e2e:
        e2e = {"value": None, "unit": "cells/s", "skipped": f"host memory {host_mem_gb():.0f} GB < {need_gb:.0f} GB needed"}

    # ---- CPU baseline: ONE step of the same workload on the host cores (rank 0, N = 1 only) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = _cpu_baseline(s, spec, wl, mask, om, host, pca if host is None else None, make_pca, ctx)

    if rank == 0:
        line = {
            "metric": "randomized_pca_fit_transform_throughput", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "seconds_per_fit_transform": ms_step * 1e-3, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, world),
            "details": {"parallelism": f"rows sharded over {world} GPU(s): {shard_rule}", "nnz_rank0": int(prof_nnz(dev, host)),
                        "arithmetic": "f32 storage and accumulation; tensor-core products take fp16 operands: the dense "
                                      "panel as two fp16 terms (22 significant bits), the operator as one term when every "
                                      "stored value is exact in fp16 (raw counts) else two; fp16 x fp16 products are exact "
                                      "in f32",
                        "tile_format": "masked fits rebuild the tile format every fit (inside value and e2e); unmasked "
                                       "operators cache it on the handle: after warm-up `value` excludes that one-time "
                                       "build, e2e (fresh upload per step) includes it",
                        "small_side": "one launch per half step for Cholesky(Y^T Y) / Gram(Z) / Cholesky -> M and one for Z M + "
                                      "centring term + fp16 pre-split (90 launches per fit, round 1: 140); the A X time of an "
                                      "inner iteration no longer contains the pre-split of its panel",
                        "host_cpus_bound_to_gpu_numa_node_rank0": numa_cpus,
                        "step": ("restore raw counts (D2D) + fused preprocess + fit_transform" if wl.get("preprocess")
                                 else "fit_transform")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernel_classes": classes, "cpu_baseline": cpu, "parity_vs_n1": parity_n1,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def prof_nnz(dev, host):
    return len(host[2]) if host is not None else dev.nnz


def _cpu_baseline(s, spec, wl, mask, om, host, pca_dev, make_pca, ctx):
    """One fit_transform of the same matrix on the host cores (oracle/cpu_ref.cpp, all cores) and, with it, the parity of
    the CUDA path on the FULL operator: singular values and largest principal angle against the CPU result."""
    from oracle import cpu_ref as R
    threads = R.set_threads()
    rows_cfg = wl["nrows"]
    if host is not None:
        off, idx, val = host
        ptr, cidx, cval, rows = off, idx.view(np.uint32), val, rows_cfg
        if wl.get("cpu_rows") and wl["cpu_rows"] < rows_cfg:        # bounded sample for the big shard
            rows = wl["cpu_rows"]
            ptr = off[:rows + 1]
            cidx, cval = cidx[:int(ptr[rows])], cval[:int(ptr[rows])]
    else:
        rows = min(rows_cfg, 40_000)
        ptr, cidx, cval = R.synth_rows(spec, 0, rows)
    raw = np.array(cval[:ptr[rows]], copy=True) if wl.get("preprocess") else None
    work = np.array(cval[:ptr[rows]], copy=True) if wl.get("preprocess") else cval
    r, dt = cpu_fit_transform(R, wl, ptr, cidx, work, rows, mask, om, raw)
    out = {"value": rows / dt, "unit": "cells/s", "cores": threads, "kind": "port",
           "sample": (f"the full {rows} x {wl['ncols']} matrix" if rows == rows_cfg else f"first {rows} of {rows_cfg} rows")
                     + f", one fit_transform: {dt:.2f} s (oracle/cpu_ref.cpp, C++17 + OpenMP restatement of the reference path)",
           "phase_seconds": r.timings}
    # parity of the CUDA path on the same rows, same Omega (the checker role of the oracle)
    x = s.CsrMatrix(rows, wl["ncols"], ptr.view(np.uint64), cidx.view(np.int32), raw if raw is not None else cval, ctx)
    p = make_pca()
    if wl.get("preprocess"):
        x.to_device()
        x._dev.preprocess_device(wl["preprocess"])
    p.fit(x, omega=om)
    out["parity_full_operator" if rows == rows_cfg else "parity_on_sample"] = {
        "sigma_rel_err_vs_cpu_f32": rel_err(p.singular_values_, r.singular_values),
        "largest_principal_angle_rad": principal_angle(p.components_, r.components),
        "rows": rows}
    x.drop_device()
    p._free_model()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SALG_BENCH_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()

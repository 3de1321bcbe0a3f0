#!/usr/bin/env python
"""bench.py — randomized sparse-PCA fit_transform on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the CPU path (oracle port) on the host cores
  torchrun ... bench.py --gpus N ...                       # rows sharded over N GPUs (strong scaling)

A step is one MaskedSparsePCA / SparsePCA fit_transform (column statistics, mask compaction, transposed
copy, 16 centred SpMM passes, 17 CholeskyQR2, Jacobi SVD, svd_flip, projection) over one synthetic count
matrix.  `value` = cells/s with the CSR resident in HBM; `e2e` = the same through the public API from
pinned HOST buffers (upload + validation inside the timed region, scores copied back).
Workloads (BASELINE.json configs): cfg3 = MaskedSparsePCA f32 1M x 30k @7 %, 2000-gene mask (default; the
config the metric's "1/2/4/8 B200" and the north-star target are quoted on), cfg2 = SparsePCA f32
100k x 20k @7 %.  Inputs are far larger than L2 (126 MB), so no explicit L2 flush is needed between steps.
"""
import argparse
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
                 nrows=1_000_000, ncols=30_000, density=0.07, n_mask=2_000, k=50, p=10, q=7),
    "cfg2": dict(name="SparsePCA f32 100kx20k CSR @7% nnz, Random{p=10,q=7,QR}, k=50, center",
                 nrows=100_000, ncols=20_000, density=0.07, n_mask=0, k=50, p=10, q=7),
    "tiny": dict(name="MaskedSparsePCA f32 20kx3k CSR @7% nnz, 500-gene mask (self-test)",
                 nrows=20_000, ncols=3_000, density=0.07, n_mask=500, k=50, p=10, q=7),
}


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


# ------------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """The reference arm: the CPU path (oracle port of the reference algorithm; the Rust crate cannot be
    built in this image) on the host cores, on a bounded row sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy.sparse as sp
    from oracle import oracle as O
    import single_algebra_b200 as s
    try:
        from threadpoolctl import threadpool_info
        thr = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        thr = os.cpu_count() or 1
    sample_rows = min(wl["nrows"], args.cpu_sample_rows)
    spec = s.synth.make_spec(wl["nrows"], wl["ncols"], density=wl["density"], seed=42)
    A = _sample_matrix(s, spec, sample_rows)
    mask = s.synth.make_mask(wl["ncols"], wl["n_mask"], seed=7) if wl["n_mask"] else None
    n_eff = wl["n_mask"] or wl["ncols"]
    om = s.synth.make_omega(n_eff, wl["k"] + wl["p"], seed=42, dtype=np.float64)

    def step():
        r = O.sparse_pca_fit(A, wl["k"], omega=om, mask=mask, n_oversamples=wl["p"], n_power_iterations=wl["q"],
                             dtype=np.float32)
        return O.transform(A, r.components, r.mean, mask=mask, mode=O.EXACT)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_rows / dt
    line = {
        "impl": "reference", "metric": "randomized_pca_fit_transform_throughput", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "seconds_per_fit_transform": dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample_rows": sample_rows},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": thr, "kind": "port",
                         "sample": f"first {sample_rows} rows of the workload matrix (all {wl['ncols']} columns); "
                                   "numpy/scipy oracle: LAPACK QR/SVD multi-threaded, scipy CSR products single-threaded"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _sample_matrix(s, spec, sample_rows):
    """First `sample_rows` rows as scipy CSR: from the device generator when a GPU is present (input
    creation only, bit-identical to the host generator), else from the host generator."""
    import scipy.sparse as sp
    if s.device_count() > 0:
        d = s.synth_device(spec, 0, sample_rows, dtype=np.float32)
        off, idx, val = d.download()
        d.free()
        return sp.csr_matrix((val, idx.astype(np.int32), off.astype(np.int64)), shape=(sample_rows, spec.ncols))
    ip, ix, dv = s.synth.generate_rows(spec, 0, sample_rows, dtype=np.float32)
    return sp.csr_matrix((dv, ix, ip), shape=(sample_rows, spec.ncols))


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
    spec = s.synth.make_spec(wl["nrows"], wl["ncols"], density=wl["density"], seed=42)
    r0, r1 = s.dist.partition_rows_even(wl["nrows"], world)[rank]
    dev = s.synth_device(spec, r0, r1 - r0, dtype=np.float32, ctx=ctx)
    mask = s.synth.make_mask(wl["ncols"], wl["n_mask"], seed=7) if wl["n_mask"] else None
    n_eff = wl["n_mask"] or wl["ncols"]
    om = s.synth.make_omega(n_eff, wl["k"] + wl["p"], seed=42, dtype=np.float32)

    def make_pca():
        b = s.MaskedSparsePCABuilder().mask(mask.tolist()) if mask is not None else s.SparsePCABuilder()
        return b.n_components(wl["k"]).svd_method(
            s.SVDMethod.Random(wl["p"], wl["q"], s.PowerIterationNormalizer.QR)).build()

    pca = make_pca()

    def step_resident():
        pca._fit(dev, om, keep_scores=True, fetch=False)

    # ---- device-resident timing -----------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    ctx.prof_reset()
    # inside the timed region only the two product classes are event-timed (the roofline's live measurement); the full
    # per-class table comes from two extra, untimed, fully profiled fits right after
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
    value = wl["nrows"] / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA events around every launch of the class) ------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # per-class table: the product classes from the timed region, the others from the untimed profiled fits scaled to
    # the same number of steps (so ms_total / steps is per fit for every class)
    scale = args.steps / n_prof_steps
    merged = {k: (v[0] * scale, int(round(v[1] * scale)), v[2] * scale) for k, v in prof_all.items()}
    merged.update(prof)
    classes = {k: {"ms_total": v[0], "launches": v[1], "gbs": (v[2] / v[0] / 1e6) if v[0] > 0 and v[2] > 0 else None}
               for k, v in merged.items()}
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two product kernels from ONE
    # `ncu --set full` capture each, on this workload at 1 GPU: profiles/r01_v3_ncu_full_tc_ax_aty.csv
    ncu_traffic = {("cfg3", "spmm"): 1.122587e9 + 0.238028e9, ("cfg3", "spmm_t"): 1.468230e9 + 0.004183e9}
    dom = max((k for k in ("spmm", "spmm_t") if k in prof), key=lambda k: prof[k][0], default=None)
    roofline = None
    if dom:
        tms, n, b = prof[dom]
        achieved = b / tms / 1e6     # GB/s
        roofline = {"bound": "hbm", "kernel": ("tc_aty_kernel (A^T Y, tcgen05 tile-densified)" if dom == "spmm_t" else "tc_ax_kernel (A X, tcgen05 tile-densified)"),
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic.get((args.workload, dom)) if world == 1 else None,
                    "traffic_source": "profiles/r01_v3_ncu_full_tc_ax_aty.csv" if world == 1 and (args.workload, dom) in ncu_traffic else None,
                    "peak_source": peak_src, "avg_launch_ms": tms / n, "launches": n,
                    "algorithmic_bytes_per_launch": b / n, "share_of_step": tms / (ms * 1.0)}

    # ---- end to end through the public API from pinned host buffers --------------------------------------
    e2e = None
    need_gb = dev.nnz * 8 / 1e9 * 1.3 + 2
    if not args.no_e2e and host_mem_gb() > need_gb:
        nnz = dev.nnz
        t_off, off = pinned_empty(dev.nrows + 1, np.int64)
        t_idx, idx = pinned_empty(nnz, np.int32)
        t_val, val = pinned_empty(nnz, np.float32)
        _download_i32(s, ctx, dev, off, idx, val)
        nloc = dev.nrows
        t_sc, sc = pinned_empty(nloc * wl["k"], np.float32)
        dev.free()   # the e2e leg owns device memory from here on

        verbose = bool(os.environ.get("SALG_BENCH_VERBOSE"))

        def step_e2e():
            t0 = time.perf_counter()
            x = s.CsrMatrix(nloc, wl["ncols"], off.view(np.uint64), idx, val, ctx)
            x.to_device()
            t1 = time.perf_counter()
            p2 = make_pca()
            out = p2.fit_transform(x, omega=om, out=sc.reshape(nloc, -1))   # scores land in pinned host memory
            t2 = time.perf_counter()
            x.drop_device()
            p2._free_model()
            t3 = time.perf_counter()
            if verbose and rank == 0:
                print(f"[e2e] upload {1e3*(t1-t0):.1f} ms, fit_transform+fetch {1e3*(t2-t1):.1f} ms, copy+free {1e3*(t3-t2):.1f} ms",
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
        e2e_classes = {k: round(v[0] / max(1, min(args.steps, 3)), 3) for k, v in ctx.prof().items()}
        ms2_step = max_over_ranks(ms2) / k_e2e
        h2d = sum_over_ranks(nnz * 8 + (nloc + 1) * 8 + om.nbytes + (wl["ncols"] if mask is not None else 0))
        d2h = sum_over_ranks(nloc * wl["k"] * 4 + 8 * wl["ncols"] + 16 * wl["k"])
        e2e = {"value": wl["nrows"] / (ms2_step * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2_step, "steps": k_e2e,
               "class_ms_per_step": e2e_classes}
    elif not args.no_e2e:
        e2e = {"value": None, "unit": "cells/s", "skipped": f"host memory {host_mem_gb():.0f} GB < {need_gb:.0f} GB needed"}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = _cpu_baseline(s, spec, wl, mask, args.cpu_sample_rows)

    if rank == 0:
        line = {
            "metric": "randomized_pca_fit_transform_throughput", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "seconds_per_fit_transform": ms_step * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "rows": wl["nrows"], "cols": wl["ncols"], "nnz_per_gpu_rank0": dev.nnz,
                       "parallelism": f"rows sharded over {world} GPU(s)", "l2": "inputs larger than L2; no flush needed",
                       "omega": "host-generated PCG64(42) standard normal",
                       "host_cpus_bound_to_gpu_numa_node_rank0": numa_cpus},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernel_classes": classes, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _download_i32(s, ctx, dev, off, idx, val):
    """Device CSR -> pinned host arrays in the scipy/AnnData layout (int64 offsets, int32 indices)."""
    dev.download_raw(off, idx.view(np.uint32), val)


def _cpu_baseline(s, spec, wl, mask, sample_rows):
    import scipy.sparse as sp
    from oracle import oracle as O
    try:
        from threadpoolctl import threadpool_info
        thr = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        thr = os.cpu_count() or 1
    sample_rows = min(wl["nrows"], sample_rows)
    A = _sample_matrix(s, spec, sample_rows)
    n_eff = wl["n_mask"] or wl["ncols"]
    om = s.synth.make_omega(n_eff, wl["k"] + wl["p"], seed=42, dtype=np.float64)
    t0 = time.perf_counter()
    r = O.sparse_pca_fit(A, wl["k"], omega=om, mask=mask, n_oversamples=wl["p"], n_power_iterations=wl["q"],
                         dtype=np.float32)
    O.transform(A, r.components, r.mean, mask=mask, mode=O.EXACT)
    dt = time.perf_counter() - t0
    # parity of the product on the same sample, same Omega (the checker role of the oracle)
    x = s.CsrMatrix.from_scipy(A)
    b = s.MaskedSparsePCABuilder().mask(mask.tolist()) if mask is not None else s.SparsePCABuilder()
    p = b.n_components(wl["k"]).svd_method(s.SVDMethod.Random(wl["p"], wl["q"], s.PowerIterationNormalizer.QR)).build()
    p.fit(x, omega=om.astype(np.float32))
    ref64 = O.sparse_pca_fit(A.astype(np.float64), wl["k"], omega=om, mask=mask, n_oversamples=wl["p"],
                             n_power_iterations=wl["q"])
    return {"value": sample_rows / dt, "unit": "cells/s", "cores": thr, "kind": "port",
            "sample": f"first {sample_rows} rows of the workload matrix, one fit_transform: {dt:.2f} s "
                      "(numpy/scipy oracle; LAPACK multi-threaded, scipy CSR products single-threaded)",
            "parity_on_sample": {"sigma_rel_err_vs_f64_oracle": O.rel_err(p.singular_values_, ref64.singular_values),
                                 "largest_principal_angle_rad": O.largest_principal_angle(p.components_, ref64.components)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SALG_BENCH_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rows", type=int, default=40_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()

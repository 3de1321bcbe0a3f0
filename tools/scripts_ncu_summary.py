"""Summarise an ncu report: key raw metrics and the top stall lines.  usage: scripts_ncu_summary.py rep [kernel-substr]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if sub not in name: continue
    print("==", name[:80])
    for k, v, u in zip(h, r, rows[1]):
        if any(k.startswith(x) or x in k for x in keys) and v not in ("", "0") and "stalled" not in k and ".per_second" not in k and ".max" not in k.replace("elapsed.max","") and ".min" not in k:
            print(f"  {k} [{u}] = {v}")
    st = [(k, float(v.replace(",", ""))) for k, v in zip(h, r) if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v]
    for k, v in sorted(st, key=lambda x: -x[1])[:8]: print(f"  stall {k[34:-24]} = {v:.2f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for b in blocks:
    if sub not in b["name"]: continue
    h = b["rows"][0]; data = b["rows"][1:]
    si = h.index("# Samples"); srci = h.index("Source")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in data if r[si].isdigit())
    print("-- source hot spots, total samples", tot)
    for r in sorted(data, key=lambda r: -int(r[si]) if r[si].isdigit() else 0)[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        st = sorted(((h[i][6:], int(r[i])) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), key=lambda x: -x[1])[:3]
        print(f"  {r[si]:>6} {r[srci].strip()[:64]:64s} {st}")
    break

#!/usr/bin/env python3
"""Summary of one kernel of an .ncu-rep: headline metrics, warp-stall reasons, hottest SASS lines.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active", "launch__registers_per_thread",
          "launch__grid_size", "launch__block_size", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct"]:
    print(k, d.get(k))
st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v)) for k, (v, u) in d.items()
      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v not in ("", "n/a")]
tot = sum(v for _, v in st)
print("stalls:", ", ".join(f"{k} {v / tot:.1%}" for k, v in sorted(st, key=lambda x: -x[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; hdr = rows[hi]
end = his[1] - 3 if len(his) > 1 else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
isrc, iss, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
I = lambda x: int(x) if x.isdigit() else 0
tot = sum(I(r[iss]) for r in data)
for n, r in enumerate(data): r.append(n)
print("samples", tot, "sass lines", len(data))
for r in sorted(sorted(data, key=lambda r: -I(r[iss]))[:top], key=lambda r: r[-1]):
    print(f"{r[-1]:5d} {I(r[iss]) / max(tot, 1):6.1%} {r[iex]:>10} {r[isrc][:110]}")

#!/usr/bin/env python3
"""Benchmark of the simplex pivot path (BASELINE.json metric: pivots/s and rank-1-update GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): C4 of BASELINE.json -- the fp64 32768 x 65536 tableau of the seeded dense
family (17.2 GB, far larger than L2, so no flush is needed between iterations), generated on the
device, row-sharded over the N GPUs (strong scaling: the tableau is fixed, each rank holds m/N rows).
A step is `--pivots-per-step` consecutive simplex pivots (Dantzig pricing + ratio test + update, the
reference's Simplex.solve loop); value = pivots/s over all ranks, device-timed with CUDA events on the
launching stream, max over ranks.

The loop runs pivots in BLOCKS (csrc/lps_f64_blk.cuh): a look-ahead panel decides up to 16 pivots from
O(m+n) numbers, one pass over the tableau applies them all, bit-identical to pivot-by-pivot updates.
The dominant kernel is that fused pass, k_blk_update: it reads and writes every cell ONCE per block.

Extra objects on the JSON line:
  roofline      k_blk_update: the bytes one launch has to move -- 16(m_loc+1)(n+1) for the cells plus the
                block's pivot rows and columns, K(8(m_loc+1) + 8(n+1)) -- / its mean CUDA-event duration in
                the timed region, against MEASURED_PEAKS.json's hbm_gbs.  `rank1_equivalent_gbs` restates
                the same time in SURVEY.md section 8(d)'s per-pivot bytes (what a pivot-at-a-time loop
                would have had to move for the pivots this launch applied); `rank1` is the rank-1 kernel
                k_update_f64 itself (north star: >= 70 % of the HBM peak), timed over a short leg of the
                pivot-at-a-time loop before the main measurement
  e2e           the same metric through the public API with HOST buffers: every step uploads the
                tableau from pinned host memory, runs the pivots, reads back basis, objective and
                basic solution
  cpu_baseline  the CPU oracle (C restatement of the reference's float path, all host cores) on the
                same tableau, a few pivots (rank 0, N = 1 only)
--impl reference times that CPU oracle alone (the reference itself is single-threaded pure Python
and cannot hold this tableau: SURVEY.md section 8d); one pivot of the full tableau per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries the ONE JSON line of the contract.  Libraries write there too (NCCL's version banner on the first
# communicator, for one): file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

METRIC, UNIT = "pivots_per_sec", "pivots/s"


def algorithmic_bytes(m_rows: int, n: int) -> int:
    """fp64 bytes one pivot must move over (m_rows+1) x (n+1) cells."""
    return 16 * (m_rows + 1) * (n + 1) + 8 * (m_rows + 1) + 8 * (n + 1)


def block_bytes(m_rows: int, n: int, k: int) -> int:
    """fp64 bytes ONE fused launch applying k pivots must move: every cell read and written once, plus the k pivot
    columns and normalised pivot rows of the panel."""
    return 16 * (m_rows + 1) * (n + 1) + k * (8 * (m_rows + 1) + 8 * (n + 1))


def base_config(a) -> dict:
    """The workload, named identically by both arms."""
    m, n = a.m, 2 * a.m
    return {"workload": f"C4 fp64 {m}x{n} dense tableau, seed {a.seed}", "m": m, "n": n,
            "pivots_per_step": a.pivots_per_step,
            "l2": f"tableau ({(m + 1) * (n + 1) * 8 / 1e9:.1f} GB total) is larger than L2; no flush needed"
                  if (m + 1) * (n + 1) * 8 > (256 << 20) else "tableau fits L2: a cache-resident workload by construction"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel: str):
    """dram bytes per launch of a kernel at C4 on one GPU from the committed `ncu --set full` capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "update_traffic.json")) as f:
            d = json.load(f)
        return d.get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                      text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_mem_available() -> int:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 1 << 62


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's float path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_rows_that_fit(m: int, n: int) -> int:
    need = (m + 1) * (n + 1) * 8
    avail = host_mem_available()
    rows = m
    while rows > 64 and (rows + 1) * (n + 1) * 8 > 0.6 * avail:
        rows //= 2
    return rows


def time_cpu_pivots(T: np.ndarray, m_rows: int, pivots: int):
    """Seconds per pivot of the oracle's solve loop (pricing + ratio test + rank-1 update) on T."""
    from oracle import c_oracle as C
    basis = np.arange(T.shape[1] - 1 - m_rows, T.shape[1] - 1, dtype=np.int32)
    t0 = time.perf_counter()
    st, k, _ = C.solve(T, basis, max_pivots=pivots, want_trace=False)
    dt = time.perf_counter() - t0
    return dt / max(k, 1), k


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as C
    m, n = a.m, 2 * a.m
    rows = cpu_rows_that_fit(m, n)
    T = np.empty((rows + 1, n + 1), dtype=np.float64)
    # a row slab of the same tableau when the host cannot hold all of it: per-pivot work is
    # proportional to the rows, so the full-size rate is the slab's rate times rows/m
    C.fill_dense(T, m, 0, a.seed, 1, 9)
    scale = (rows + 1) / (m + 1)
    basis = np.arange(m, m + rows, dtype=np.int32)
    times = []
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        C.solve(T, basis, max_pivots=1, want_trace=False)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
    sec = float(np.mean(times)) / scale
    B = a.pivots_per_step
    sample = f"bounded sample: 1 pivot (pricing + ratio test + rank-1 update) timed per step instead of {B}, on the " \
             f"{rows} x {n} " + ("full tableau" if rows == m else f"row slab of the {m} x {n} tableau, time scaled by rows") \
             + f", {a.steps} steps; ms_per_step = that pivot's time x {B}"
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3 * B, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": base_config(a),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": C.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is single-threaded pure Python (fractions.Fraction) and cannot hold this tableau; "
                "this arm is oracle/pivot_oracle.c, the C restatement of its float path, on all host cores",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200.sharded import ShardedF64Solver, shard_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m, n, B = a.m, 2 * a.m, a.pivots_per_step
    row0, m_loc = shard_rows(m, rank, world)
    solver = ShardedF64Solver(m, n, rank, world, device=local, exchange=a.exchange, block=a.block)
    solver_exchange = solver.exchange            # "nccl" when the peers' exchange blocks could not be mapped
    eng = solver.eng
    K = solver.block                             # pivots per block; 0: the pivot-at-a-time loop
    basis = np.arange(m, 2 * m, dtype=np.int32)
    peak, peak_src = measured_peak()

    # ---- short leg of the pivot-at-a-time loop: the rank-1 kernel k_update_f64 against the roofline ----
    rank1 = None
    if K and a.rank1_pivots > 0:
        eng.fill_dense(a.seed, 1, 9)
        eng.set_block_size(0)
        solver.block = 0
        solver.begin(basis.copy(), None, -1, 0)
        solver.enqueue(3)
        solver.poll()
        eng.prof_reset()
        barrier()
        solver.enqueue(a.rank1_pivots)
        solver.poll()
        r1_ms, r1_n, _ = eng.prof_read()
        solver.end(basis.copy(), None, 0)
        r1_avg = max_over_ranks(r1_ms / max(r1_n, 1))
        r1_bytes = algorithmic_bytes(m_loc, n)
        r1_ach = r1_bytes / (r1_avg * 1e-3) / 1e9
        rank1 = {"kernel": "k_update_f64", "achieved": r1_ach, "frac": r1_ach / peak, "bytes_per_launch": r1_bytes,
                 "ms_per_launch": r1_avg, "launches_timed": int(r1_n), "traffic": recorded_traffic("k_update_f64") if world == 1 else None}
        eng.set_block_size(K)
        solver.block = K

    eng.fill_dense(a.seed, 1, 9)
    solver.begin(basis, None, -1, 0)

    def step():
        solver.enqueue(B)
        return solver.poll()

    for _ in range(a.warmup):
        step()
    _, piv0 = solver.poll()
    sampler = ClockSampler(local) if rank == 0 else None
    eng.prof_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(eng.stream)                      # CUDA events on the stream the kernels are launched on
    for _ in range(a.steps):
        status, piv1 = step()
    e1.record(eng.stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    upd_ms, upd_launches, all_launches = eng.prof_read()
    pivots = piv1 - piv0
    value = pivots / (ms * 1e-3)
    kernel = "k_blk_update" if K else "k_update_f64"
    per_launch = pivots / max(upd_launches, 1)   # pivots one launch applied (K, fewer only at the end of a solve)
    bytes_launch = block_bytes(m_loc, n, K) if K else algorithmic_bytes(m_loc, n)
    upd_avg_ms = max_over_ranks(upd_ms / max(upd_launches, 1))
    achieved = bytes_launch / (upd_avg_ms * 1e-3) / 1e9 if upd_launches else None
    solver.end(basis, None, 0)

    # ---- e2e: host buffers through the public API, upload + pivots + read-back every step ----------
    host = torch.empty((m_loc + 1, n + 1), dtype=torch.float64, pin_memory=True)
    eng.fill_dense(a.seed, 1, 9)
    N.check(eng.L.lps_download_f64(eng.h, host.data_ptr(), n + 1))
    host_np = host.numpy()
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    h2d = (m_loc + 1) * (n + 1) * 8
    d2h_box = [0]
    parts = []                                   # per e2e step: (upload + setup, pivots, read-back) seconds on this rank

    if world == 1:
        from linear_program_solver_b200 import Simplex, Tableau
        solver.close()

        def e2e_step():
            t0 = time.perf_counter()
            t = Tableau.from_grid(host_np, mode="fp64", device=local)
            s = Simplex(t)                      # reads b, uploads, checks canonical form on the device
            t1 = time.perf_counter()
            s.solve(max_pivots=B)
            t2 = time.perf_counter()
            z = s.getObjValue()
            x = s.getBFS()
            d2h_box[0] = 8 * (n + 1) + 8 * (m + 1) + 4 * m + 4 * n + 12 * B
            t3 = time.perf_counter()
            t._drop_engine()                    # device buffers released inside the timed step (not part of read_back)
            parts.append((t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3))
            return z, len(x)
    else:
        def e2e_step():
            t0 = time.perf_counter()
            N.check(eng.L.lps_upload_f64(eng.h, host.data_ptr(), n + 1))
            eng.sync()
            t1 = time.perf_counter()
            bs = np.arange(m, 2 * m, dtype=np.int32)
            st_, k_, _ = solver.solve(bs, None, max_pivots=B)
            t2 = time.perf_counter()
            z = -eng.read_row(m_loc)[n]
            x = eng.read_col(n)
            d2h_box[0] = 8 * (n + 1) + 8 * (m_loc + 1) + 4 * m
            t3 = time.perf_counter()
            parts.append((t1 - t0, t2 - t1, t3 - t2))
            return z, len(x)

    e2e_step()                                   # warm-up (engine creation, pinned mappings)
    parts.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = e2e_steps * B / e2e_s

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on the same host tableau ----------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import c_oracle as C
        sec1, _ = time_cpu_pivots(host_np, m, 1)
        extra = int(max(1, min(16, 12.0 / max(sec1, 1e-3))))
        sec, k = time_cpu_pivots(host_np, m, extra)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": C.num_threads(), "kind": "port",
               "sample": f"{k} consecutive pivots of the full {m} x {n} tableau after 1 warm-up pivot "
                         "(oracle/pivot_oracle.c, pthread row-parallel)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(base_config(a), sharding=f"rows/{world}", rows_per_gpu=m_loc,
                           loop=(f"blocks of {K} pivots: look-ahead panel + one fused pass over the tableau" if K
                                 else "one rank-1 update per pivot"),
                           exchange=("none (one rank)" if world == 1 else
                                     "peer stores over NVLink inside the pivot kernels" if solver_exchange == "p2p" else
                                     "one NCCL all-gather per pivot"),
                           status_after=int(status), pivots_timed=int(pivots)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": recorded_traffic(kernel) if world == 1 else None,
                         "kernel": kernel, "bytes_per_launch": bytes_launch, "ms_per_launch": upd_avg_ms,
                         "launches_timed": int(upd_launches), "pivots_per_launch": per_launch,
                         "peak_source": peak_src,
                         # the same launch time in SURVEY 8(d)'s per-pivot bytes: what a pivot-at-a-time loop would
                         # have had to move for the pivots one launch applies (exceeds the HBM peak by construction:
                         # the block reads and writes the tableau once for all of them)
                         "rank1_equivalent_gbs": (algorithmic_bytes(m_loc, n) * per_launch / (upd_avg_ms * 1e-3) / 1e9)
                         if upd_launches else None,
                         "loop_frac": bytes_launch * upd_launches / (ms * 1e-3) / 1e9 / peak,
                         "rank1": rank1},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * world if world > 1 else h2d,
                    "d2h_bytes_per_step": d2h_box[0], "steps": e2e_steps, "pivots_per_step": B,
                    "seconds_per_step": {"upload_and_setup": float(np.mean([p[0] for p in parts])),
                                         "pivots": float(np.mean([p[1] for p in parts])),
                                         "read_back": float(np.mean([p[2] for p in parts])),
                                         "release": float(np.mean([p[3] for p in parts if len(p) > 3] or [0.0]))},
                    "api": "Tableau.from_grid + Simplex + solve(max_pivots)" if world == 1 else
                           "ShardedF64Solver: lps_upload_f64 + solve + read-back per rank"},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--m", type=int, default=32768, help="constraint rows (n = 2m); default is config C4")
    ap.add_argument("--pivots-per-step", type=int, default=4096,
                    help="consecutive simplex pivots per step (C4 takes 193 374 pivots to optimality: profiles/r02d_c4_time_to_optimal.json)")
    ap.add_argument("--block", type=int, default=None, help="pivots per block (default: the library's, 16; 0 = one rank-1 update per pivot)")
    ap.add_argument("--rank1-pivots", type=int, default=24, help="length of the pivot-at-a-time leg that times k_update_f64 (0: skip)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the ranks exchange the leaving-row candidates and the pivot row")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Config C5 of BASELINE.json: a batch of independent 64 x 128 exact node LPs through k_bnb_nodes.

    python tools/bench_batch.py [--count 4096] [--m 64] [--cluster 0,4,8] [--reps 5] [--cpu-seconds 10]

Prints one JSON line per cluster width: device time of the kernel (CUDA events on its stream, inputs
resident in HBM), node LPs/s and pivots/s, the shared-memory traffic the updates need
(8 P (m+1)(n+1) bytes per pivot: every residue read once and written once) against the aggregate
shared-memory bandwidth, the end-to-end rate through `solve_batch` with host buffers (H2D of the
grids, kernel, D2H + decode of every objective), and the CPU oracle (the reference's algorithm in
Python Fractions, 1 core) on a bounded sample of the same LPs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_two_tier(grids, primes, art_cap, cl, reps, m, n, device=None, barrier=None):
    """Device time of solve_batch's two launches: half the primes with the per-LP capacity check, then the LPs that
    asked for more primes on the full set.  None when the plan has no shorter tier."""
    from linear_program_solver_b200 import _native as NN
    from linear_program_solver_b200.batch import BatchExactSolver
    count = grids.shape[0]
    short = len(primes) // 2
    short += short & 1
    if not (2 <= short < len(primes)):
        return None
    ps = sorted(primes)
    s1 = BatchExactSolver(count, m, n, ps[:short], art_cap=art_cap, cluster=cl, trace_cap=0, device=device)
    s1.set_capacity_check(True)
    s1.upload(grids)
    s1.run(64 * (m + n)); s1.sync()
    t1 = []
    for _ in range(reps):
        if barrier:
            barrier()
        s1.run(64 * (m + n)); s1.sync()
        t1.append(s1.info()["last_run_ms"])
    i1 = s1.info()
    st1 = s1.download(want_numbers=False).status
    s1.close()
    redo = np.nonzero(st1 == NN.NEED_PRIMES)[0]
    t2 = [0.0]
    if redo.size:
        s2 = BatchExactSolver(int(redo.size), m, n, ps, art_cap=art_cap, cluster=cl, trace_cap=0, device=device)
        s2.upload(np.ascontiguousarray(grids[redo]))
        s2.run(64 * (m + n)); s2.sync()
        t2 = []
        for _ in range(reps):
            s2.run(64 * (m + n)); s2.sync()
            t2.append(s2.info()["last_run_ms"])
        s2.close()
    ms = float(np.median(t1)) + float(np.median(t2))
    return {"primes_first": short, "cluster_first": i1["cluster"], "resident_first": i1["resident_clusters"],
            "ms_first": float(np.median(t1)), "solved_again": int(redo.size), "ms_second": float(np.median(t2)),
            "ms": ms, "node_lps_per_s": count / (ms * 1e-3), "optimal_first": int((st1 == 1).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=4096)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--cluster", default="0")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--family", choices=["hashed", "seeded"], default="hashed")
    ap.add_argument("--timing", action="store_true", help="print the kernel's per-phase cycle shares (profiling aid)")
    ap.add_argument("--gpus", type=int, default=1, help="under torchrun: every rank solves its own --count LPs (weak scaling)")
    a = ap.parse_args()
    if a.gpus > 1:
        return main_multi(a)

    import torch
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.batch import BatchExactSolver, plan_batch, solve_batch

    m, n, count = a.m, 2 * a.m, a.count
    t0 = time.perf_counter()
    if a.family == "hashed":
        grids = np.stack([problems.hashed_dense_grid(m, seed) for seed in range(count)]).astype(np.int32)
    else:
        grids = np.stack([np.asarray(problems.dense_lp_grid(m, seed), dtype=np.int32) for seed in range(count)])
    gen_s = time.perf_counter() - t0
    art_cap, primes, bits = plan_batch(grids)
    P = len(primes)
    sm_count = torch.cuda.get_device_properties(0).multi_processor_count
    smem_peak = sm_count * 128 * 1.965e9 / 1e9          # GB/s: 128 B/clk/SM at the B200's 1965 MHz boost clock

    cpu = None
    if a.cpu_seconds > 0:
        from oracle import lpsol_oracle as O
        done, piv, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < a.cpu_seconds and done < count:
            g = grids[done]
            res = O.simplex([int(v) for v in g[m, :n]], [[int(v) for v in row[:n]] for row in g[:m]], [int(v) for v in g[:m, n]])
            piv += len(res["pivots"])
            done += 1
        dt = time.perf_counter() - t0
        cpu = {"value": done / dt, "unit": "node LPs/s", "pivots_per_s": piv / dt, "cores": 1, "kind": "port",
               "sample": f"{done} of the {count} LPs, oracle/lpsol_oracle.py (the reference's Fraction simplex), {dt:.1f} s"}

    for cl in [int(x) for x in a.cluster.split(",")]:
        solver = BatchExactSolver(count, m, n, primes, art_cap=art_cap, cluster=cl, trace_cap=0)
        solver.upload(grids)
        solver.run(64 * (m + n))
        solver.sync()                                     # warm-up
        times = []
        for _ in range(a.reps):
            solver.run(64 * (m + n))
            solver.sync()
            times.append(solver.info()["last_run_ms"])
        info = solver.info()
        if a.timing:
            solver.timing(read=False)
            solver.run(64 * (m + n))
            solver.sync()
            tm = solver.timing()
            tot = tm.sum()
            npv = solver.download(want_numbers=False).npiv[:, :3].sum()
            print("phase cycles per pivot (CTA 0 of each cluster):", {name: round(float(tm[:, k].sum()) / npv, 1)
                  for k, name in enumerate(solver.PHASES)}, "total/pivot", round(float(tot) / npv, 1), flush=True)
        res = solver.download()
        pivots = int(res.npiv[:, :3].sum())
        ok = int((res.status == 1).sum())
        ms = float(np.median(times))
        solver.close()
        # two tiers (what solve_batch does): half the primes with the per-LP capacity check, then the LPs that
        # asked for more primes again on the full set; device time of both launches
        tier = time_two_tier(grids, primes, art_cap, cl, a.reps, m, n)
        # end to end with host buffers through the public call
        solve_batch(grids[:8], cluster=cl).objectives()      # warm-up: torch's allocator and copy kernels, first use
        e2e_all = []
        for _ in range(3):
            t0 = time.perf_counter()
            r2 = solve_batch(grids, cluster=cl)
            objs = r2.objectives()
            e2e_all.append(time.perf_counter() - t0)
        e2e_s = float(np.median(e2e_all))
        smem_bytes = 8.0 * P * (m + 1) * (n + 1) * pivots
        line = {
            "workload": f"C5: {count} independent {m}x{n} exact node LPs ({a.family} dense family, 1..9 coefficients)",
            "kernel": "k_bnb_nodes", "cluster": info["cluster"], "primes": P, "bound_bits": bits,
            "rows_per_cta": info["rows_per_cta"], "smem_bytes_per_cta": info["smem_bytes"],
            "resident_clusters": info["resident_clusters"], "ms": ms, "ms_all": times,
            "node_lps_per_s": count / (ms * 1e-3), "pivots": pivots, "pivots_per_s": pivots / (ms * 1e-3),
            "optimal": ok,
            "smem_roofline": {"bound": "shared memory", "achieved": smem_bytes / (ms * 1e-3) / 1e9, "peak": smem_peak,
                              "unit": "GB/s", "frac": smem_bytes / (ms * 1e-3) / 1e9 / smem_peak,
                              "bytes_per_pivot": 8 * P * (m + 1) * (n + 1)},
            "e2e": {"node_lps_per_s": count / e2e_s, "seconds": e2e_s, "seconds_all": e2e_all, "h2d_bytes": int(grids.nbytes),
                    "d2h_bytes": int(count * (4 + 16 + 4 * m + (m + 2) * (1 + 4 * P))),
                    "api": "batch.solve_batch + objectives()"},
            "gen_seconds": gen_s,
        }
        if tier:
            line["single_tier_node_lps_per_s"] = line["node_lps_per_s"]
            line["two_tier"] = tier
            line["node_lps_per_s"] = max(line["node_lps_per_s"], tier["node_lps_per_s"])
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)


def main_multi(a):
    """Weak scaling over the GPUs of one box: rank g solves LPs with seeds g*count .. (g+1)*count-1, no collective on
    the data path; device-timed kernel (max over ranks) and the end-to-end call, both summed over ranks."""
    import torch
    import torch.distributed as dist
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.batch import BatchExactSolver, plan_batch, solve_batch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == a.gpus
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, n, count = a.m, 2 * a.m, a.count
    grids = np.stack([problems.hashed_dense_grid(m, rank * count + s) for s in range(count)]).astype(np.int32)
    art_cap, primes, bits = plan_batch(grids, device=local)
    tier = time_two_tier(grids, primes, art_cap, int(a.cluster.split(",")[0]), a.reps, m, n, device=local, barrier=dist.barrier)
    if tier is None:
        solver = BatchExactSolver(count, m, n, primes, art_cap=art_cap, cluster=int(a.cluster.split(",")[0]), trace_cap=0, device=local)
        solver.upload(grids)
        solver.run(64 * (m + n))
        solver.sync()
        times = []
        for _ in range(a.reps):
            dist.barrier()
            solver.run(64 * (m + n))
            solver.sync()
            times.append(solver.info()["last_run_ms"])
        ok = int((solver.download(want_numbers=False).status == 1).sum())
        solver.close()
    else:
        times, ok = [tier["ms"]], tier["optimal_first"]
    dist.barrier()
    t0 = time.perf_counter()
    r2 = solve_batch(grids, device=local)
    objs = r2.objectives()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([float(np.median(times)), e2e_s, float(ok)], dtype=torch.float64, device="cuda")
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"workload": f"C5 x {world} GPUs: {count} independent {m}x{n} exact node LPs per GPU", "n_gpus": world,
                          "scaling": "weak", "primes": len(primes), "ms_max_over_ranks": float(mx[0]),
                          "node_lps_per_s": world * count / (float(mx[0]) * 1e-3), "optimal": int(sm[2]),
                          "e2e": {"node_lps_per_s": world * count / float(mx[1]), "seconds_max_over_ranks": float(mx[1])}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

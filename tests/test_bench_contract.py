"""bench.py's reference arm (the CPU leg of the contract) at a small size, as the driver launches it: ONE JSON line on
stdout with the contract's keys, nothing from any rank but rank 0, exit code 0.  The GPU arm needs a B200 and is
exercised by the driver itself (profiles/*bench*.json are its lines)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--m",
                           "256", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_contract_line():
    p = _run({"RANK": "0", "WORLD_SIZE": "2"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pivots_per_sec" and d["unit"] == "pivots/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 1
    B = d["config"]["pivots_per_step"]                     # the step both arms name; the CPU arm times 1 pivot of it
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3 * B) < 1e-6 * 1e3 * B
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C4 fp64 256x512") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256 x 512" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    p = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_algorithmic_bytes_are_the_surveys_figures():
    """16(m+1)(n+1) + 8(m+1) + 8(n+1) bytes per fp64 pivot -- 128.1 MB at C2, 34.36 GB at C4
    (5.26 ms at the measured 6531.9 GB/s: the 190 pivots/s ceiling the bench's roofline fraction is taken against)."""
    code = ("import bench, os\n"
            "os.write(bench._REAL_STDOUT, ('%d %d' % (bench.algorithmic_bytes(2000, 4000), "
            "bench.algorithmic_bytes(32768, 65536))).encode())\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-1500:]
    c2, c4 = map(int, p.stdout.split())
    assert c2 == 16 * 2001 * 4001 + 8 * 2001 + 8 * 4001 and round(c2 / 1e6, 1) == 128.1
    assert c4 == 16 * 32769 * 65537 + 8 * 32769 + 8 * 65537 and round(c4 / 1e9, 2) == 34.36
    assert 189.0 < 6531.9e9 / c4 < 191.0


def test_block_bytes_and_shared_config():
    """One fused launch applying k pivots moves every cell once plus k pivot rows and columns; both arms name the
    workload with the same config keys."""
    code = ("import bench, os, json, argparse\n"
            "a = argparse.Namespace(m=32768, seed=0, pivots_per_step=1024)\n"
            "os.write(bench._REAL_STDOUT, json.dumps([bench.block_bytes(32768, 65536, 16), bench.block_bytes(2000, 4000, 1), "
            "bench.base_config(a)]).encode())\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-1500:]
    b16, b1, cfg = json.loads(p.stdout)
    assert b16 == 16 * 32769 * 65537 + 16 * (8 * 32769 + 8 * 65537)
    assert b1 == 16 * 2001 * 4001 + 8 * 2001 + 8 * 4001            # k = 1: SURVEY 8(d)'s per-pivot figure
    assert cfg["workload"] == "C4 fp64 32768x65536 dense tableau, seed 0" and cfg["pivots_per_step"] == 1024
    assert set(cfg) == {"workload", "m", "n", "pivots_per_step", "l2"}

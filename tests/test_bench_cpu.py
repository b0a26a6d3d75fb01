"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the reference path)
prints ONE JSON line with the keys the driver reads, and the algorithmic-byte model of DESIGN.md section 4."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torchrun exports; the arm must still use every thread
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "plumbing_10k_256", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("fwd+bwd frames/sec") and d["value"] > 0 and d["steps"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "oracle" in cb["sample"]
    assert d["config"]["workload"] == "plumbing_10k_256" and d["config"]["image"] == [256, 256]
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    assert cb["cores"] == avail                       # OMP_NUM_THREADS=1 was overridden


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_worked_example():
    """SURVEY.md section 8d's worked example: P = P_vis = 1 M, N = 8 M, 1080p, C = 3, degree 3, 45 key bits."""
    sys.path.insert(0, ROOT)
    import bench
    d = bench.algorithmic_bytes(1_000_000, 1_000_000, 8_000_000, 1080, 1920, 3, 3, 13)
    assert d["preprocess_fwd"] == 1_000_000 * (236 + 48)
    assert d["blend_fwd"] == 8_000_000 * 48 + 1920 * 1080 * (4 * 5 + 8)
    assert d["tile_sort"] == 8_000_000 * 12 * 2
    total = sum(d.values())
    assert 1.5e9 < total < 3.2e9                      # the survey's 3.0 GB uses 6 passes of a 64-bit sort

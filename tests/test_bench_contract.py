"""CPU: the bench.py JSON contract of the reference arm (the CUDA arm needs a GPU and is exercised by
the driver), and the roofline bookkeeping helpers."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_reference(*extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
           "--workload", "configs[2]", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = run_reference()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "d2q9_glups" and d["unit"] == "GLUPS"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["value"] > 0 and d["steps"] == 2 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert run_reference("--gpus", "2", env=env) == []


def test_committed_traffic_entries():
    """roofline.traffic comes from a committed ncu capture of ONE launch of the dominant kernel; an
    entry only applies to the kernel it was measured on (a stale figure for another kernel is worse
    than none -- round-1 verdict)."""
    sys.path.insert(0, ROOT)
    import bench
    e = bench.committed_traffic("configs[3]", "alb::march2_kernel")
    cells = 32768 * 16384
    assert e["kernel"] == "alb::march2_kernel" and os.path.exists(os.path.join(ROOT, e["source"].split()[0]))
    # one launch performs two updates per cell (144 B algorithmic) while reading and writing the state
    # about once: between 72 and 90 B of DRAM traffic per cell
    assert 72.0 * cells * 0.9 < e["dram_bytes_per_launch"] < 90.0 * cells
    assert e["algorithmic_bytes_per_launch"] == 2 * 72.0 * cells
    assert bench.committed_traffic("configs[3]", "alb::step_kernel<MODE_STEP>") is None
    assert bench.committed_traffic("no such workload", "alb::march2_kernel") is None

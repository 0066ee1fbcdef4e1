"""bench.py contract pieces that need no GPU: the reference arm (CPU restatement timed on the host cores) prints one JSON line
with the keys the driver reads; the algorithmic-bytes formula of SURVEY 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n_side", "12", "--cpu_seconds", "1"],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "md_steps_per_s" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["n_beads"] == 12 ** 3 and "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    lh = bench.l_half()
    assert abs(lh - 38.81) < 0.02                       # (2 pi / 3)(rc + skin)^3 rho at rc 2.5, skin 0.3, rho 0.8442
    assert abs((128 + 4 * lh) - 283.25) < 0.1           # bytes per bead-step (C2 283.3 B)

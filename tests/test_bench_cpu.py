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


def test_cpu_leg_runs_in_its_own_process(tmp_path):
    """The cpu_baseline leg of the GPU arm: `bench.py --impl cpu_leg --state <pickle>` restarts the oracle from the pickled host
    snapshot (pinned OpenMP threads live in that process only) and prints the cpu_baseline object."""
    import pickle
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WorkloadC2(12)
    s = dict(wl.system()); s["lists_now"] = {}; s["excl_now"] = s["exclusions"]; s["step"] = 0
    p = tmp_path / "snap.pkl"
    with open(p, "wb") as f:
        pickle.dump(s, f, protocol=4)
    for extra in ([], ["--no_reactions"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "cpu_leg", "--state", str(p), "--n_side", "12", "--steps", "6", "--warmup", "1"] + extra,
                           capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert d["kind"] == "port" and d["steps"] == 6 and d["value"] > 0 and len(d["segment_rates"]) == 3
        assert ("no reaction pass" in d["sample"]) == bool(extra)


def test_replicated_workload_is_periodic_copies_of_the_recorded_system():
    """--workload c5 (rim135 recorded from the chemlab driver, then replicated k^3 times): every per-interaction energy of the
    k = 2 system is exactly 8x the energy of the recorded system, and it holds 8x the beads, tuples and exclusions."""
    sys.path.insert(0, ROOT)
    import numpy as np
    from chemlab_b200 import synthetic
    from oracle.engine_adapter import OracleEngine
    res = {}
    for k in (1, 2):
        wl = synthetic.make_workload("c5", k, example_root=os.path.join(ROOT, "tests", "golden"))
        sysd = wl.system()
        o = OracleEngine(sysd["box"], wl.rc, wl.skin, seed=1)
        o.set_particles(sysd["ids"], sysd["type"], sysd["pos"], sysd["mass"], vel=sysd["vel"], q=sysd.get("q"), state=sysd["state"], res_id=sysd["resid"])
        h = wl.setup(o, sysd)
        o.compute_forces()
        res[k] = (sysd["n"], {name: o.energy(v) for name, v in h["energies"].items()}, {name: o.list_size(v) for name, (v, ar) in h["lists"].items()},
                  len(o.get_exclusions()), float(np.abs(o.get_particles(fields=("force",))["force"]).max()))
    assert res[2][0] == 8 * res[1][0] and res[2][3] == 8 * res[1][3]
    assert all(res[2][2][name] == 8 * res[1][2][name] for name in res[1][2])
    for name, e1 in res[1][1].items():
        assert abs(res[2][1][name] - 8 * e1) <= 1e-9 * max(1.0, abs(8 * e1)), (name, e1, res[2][1][name])
    assert abs(res[2][4] - res[1][4]) <= 1e-9 * res[1][4]

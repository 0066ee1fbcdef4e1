#!/usr/bin/env python
"""bench.py -- MD steps/s and pair-interactions/s of chemlab's reactive-MD hot path on B200.

Workload (BASELINE.json configs[1], SURVEY 8d "config 2"): synthetic 1,000,000-bead LJ reactive melt of
A-L-A trimers, rho=0.8442, rc=2.5, skin=0.3, dt=0.005, kT=1, gamma=1, pair potential = LJ tabulated on
1500 rows (dr=0.002, linear interpolation), harmonic bonds K=30 r0=0.97, harmonic angle 180 deg,
step-growth reaction A(1,2)+A(1,2)->A(1):A(1), cutoff 1.2, interval 200, p=0.05, nearest partner.

A "step" is one Velocity-Verlet step of the whole system (neighbour rebuilds and reaction passes included
at their natural frequency).  One JSON line is printed by rank 0:
  value   : steps/s, device-timed (CUDA events on the engine's stream), state resident in HBM
  e2e     : steps/s through the public C-ABI with HOST buffers: upload of the full particle state,
            run in chunks of 200 steps with the observables read back per chunk, final state download
  roofline: pair-force kernel, algorithmic bytes (40 + 4*L_half per particle per launch) over
            its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the fp64 CPU restatement (oracle/) of the same path timed on the host cores
`--impl reference` times that CPU restatement alone (the reference's own ESPResSo++ cannot be built
here: SURVEY 8c) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RC, SKIN, DT, KT, GAMMA, RHO = 2.5, 0.3, 0.005, 1.0, 1.0, 0.8442
INTERVAL, P_ACCEPT = 200, 0.05
SEED = 12347


def l_half():
    return (2.0 * np.pi / 3.0) * (RC + SKIN) ** 3 * RHO


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_engine(sysd, seed=SEED, device=0):
    from chemlab_b200 import Engine
    from chemlab_b200 import synthetic
    e = Engine(sysd["box"], RC, SKIN, seed=seed, device=device)
    e.set_particles(sysd["ids"], sysd["type"], sysd["pos"], sysd["mass"], vel=sysd["vel"], state=sysd["state"], res_id=sysd["resid"])
    h = synthetic.setup_reactive_melt(e, sysd, rc=RC, dt=DT, kT=KT, gamma=GAMMA, interval=INTERVAL, p_accept=P_ACCEPT)
    return e, h


def snapshot(e, sysd, h):
    """Host copy of the full dynamic state of an engine (the e2e leg and the CPU baseline restart from it)."""
    st = e.get_particles(fields=("pos", "vel", "type", "state", "mass", "image"))
    s = dict(sysd)
    s["pos"] = st["pos"] + st["image"] * sysd["box"]
    s["vel"] = st["vel"]; s["type"] = st["type"]; s["state"] = st["state"]; s["mass"] = st["mass"]
    s["react_bonds"] = e.list_get(h["react_list"], 2)
    s["angles_now"] = e.list_get(h["angle_list"], 3)
    s["excl_now"] = e.get_exclusions()
    return s


def restore_into(api, s, h):
    if len(s["react_bonds"]):
        api.list_add(h["react_list"], s["react_bonds"])
    extra = s["angles_now"][len(s["angles"]):]
    if len(extra):
        api.list_add(h["angle_list"], extra)
    api.set_exclusions(s["excl_now"])


def cpu_baseline(s, target_seconds=15.0, threads=None):
    """Time the CPU restatement (oracle/) on the same state: all host threads, bounded sample."""
    from oracle import pyoracle
    from chemlab_b200 import synthetic
    n = s["n"]
    o = pyoracle.Oracle(n, s["box"], RC, SKIN, seed=SEED)
    nt = threads or o.max_threads()
    o.set_threads(nt)
    o.set_particles(s["pos"], s["vel"], s["mass"], None, s["type"], s["state"], s["resid"])
    h = synthetic.setup_reactive_melt(o, s, rc=RC, dt=DT, kT=KT, gamma=GAMMA, interval=INTERVAL, p_accept=P_ACCEPT)
    restore_into(o, s, h)
    o.reaction_general(1, INTERVAL, 1, 0)
    t0 = time.perf_counter(); o.run(1); t1 = time.perf_counter() - t0   # includes the first list build
    t0 = time.perf_counter(); o.run(2); t2 = (time.perf_counter() - t0) / 2
    nsteps = int(max(3, min(200, target_seconds / max(t2, 1e-6))))
    t0 = time.perf_counter(); o.run(nsteps); dt = time.perf_counter() - t0
    return {"value": nsteps / dt, "unit": "steps/s", "cores": nt, "kind": "port",
            "sample": "%d MD steps of the same %d-bead workload (fp64 C restatement, OpenMP x%d; first step incl. list build %.2fs)" % (nsteps, n, nt, t1),
            "seconds": dt, "steps": nsteps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--equil", type=int, default=1000, help="untimed equilibration steps before warm-up (reactions off)")
    ap.add_argument("--n_side", type=int, default=100, help="beads per box edge (100 -> 1,000,000 beads)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--cpu_seconds", type=float, default=15.0)
    ap.add_argument("--option", action="append", default=[], help="engine option name=value")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from chemlab_b200 import synthetic
    workload = "C2 synthetic %d-bead LJ reactive trimer melt (tabulated LJ 1500 rows + harmonic bonds/angles + step-growth reaction)" % (a.n_side ** 3)
    config = {"workload": workload, "n_beads": a.n_side ** 3, "rho": RHO, "rc": RC, "skin": SKIN, "dt": DT, "kT": KT, "gamma": GAMMA,
              "reaction_interval": INTERVAL, "p_accept": P_ACCEPT, "nearest": True,
              "l2": "per-step working set (pos+vel+force+lists ~0.36 GB at 1M beads) exceeds the 126 MB L2; no explicit flush",
              "parallelism": "slab%d" % a.gpus if a.gpus > 1 else "single"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        sysd = synthetic.trimer_melt(a.n_side, rho=RHO, seed=12345, kT=KT)
        s = dict(sysd); s["react_bonds"] = np.zeros((0, 2), np.int64); s["angles_now"] = sysd["angles"]; s["excl_now"] = sysd["exclusions"]
        cb = cpu_baseline(s, target_seconds=min(60.0, max(5.0, a.cpu_seconds * 2)))
        line = {"impl": "reference", "metric": "md_steps_per_s", "value": cb["value"], "unit": "steps/s", "n_gpus": a.gpus, "steps": cb["steps"],
                "warmup": 3, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "CPU restatement of the reference algorithm (oracle/), NOT ESPResSo++ itself: cgchemlab/espressopp is not vendored and cannot be built here (SURVEY 8c); bounded sample, rate is per MD step of the full workload"}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the engine has no CPU fallback"}))
        return 2
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = local

    sysd = synthetic.trimer_melt(a.n_side, rho=RHO, seed=12345, kT=KT)
    n = sysd["n"]
    from chemlab_b200 import Engine as _E
    e = _E(sysd["box"], RC, SKIN, seed=SEED, device=device)
    for kv in a.option:
        k, v = kv.split("=")
        e.set_option(k, float(v))
    if world > 1:
        e.join()          # one engine per rank = one z-slab; torch.distributed only carries the NCCL id
    e.set_particles(sysd["ids"], sysd["type"], sysd["pos"], sysd["mass"], vel=sysd["vel"], state=sysd["state"], res_id=sysd["resid"])
    h = synthetic.setup_reactive_melt(e, sysd, rc=RC, dt=DT, kT=KT, gamma=GAMMA, interval=INTERVAL, p_accept=P_ACCEPT)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # untimed equilibration (lattice start -> melt), then enable reactions, then warm-up
    if a.equil > 0:
        e.run(a.equil)
    e.reaction_general(1, INTERVAL, 1, 0)
    if a.warmup > 0:
        e.run(a.warmup)
    # the read-back is collective on a multi-rank engine: every rank takes the snapshot
    snap = snapshot(e, sysd, h) if not (a.no_e2e and (a.no_cpu_baseline or world > 1)) else None

    e.reset_timers()
    e.set_option("pair_event_timing", 1)
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    e.run(a.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    tm, cn = e.timers()
    t_dev = tm["total"]
    pair_ms = e.get_option("pair_kernel_ms")
    pair_launches = e.get_option("pair_kernel_launches")
    if dist is not None:
        # timing = max over ranks; extensive counters = sum over ranks
        tt = torch.tensor([t_dev, t_wall, pair_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_wall, pair_ms = float(tt[0]), float(tt[1]), float(tt[2])
        cc = torch.tensor([cn["list_entries"], cn["launches"], cn["ghosts"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cn["list_entries"], cn["launches"], cn["ghosts"] = int(cc[0]), int(cc[1]), int(cc[2])
    e.set_option("pair_event_timing", 0)
    e.energy(h["nb"])                       # also counts the pairs inside the force cutoff
    _, cn2 = e.timers()
    interacting = cn2["interacting_pairs"]
    kin = e.kinetics()
    nbonds_new = e.list_size(h["react_list"])

    line = None
    if rank == 0:
        steps_per_s = a.steps / t_dev
        peak, peak_src = measured_peaks()
        lh = l_half()
        n_per_gpu = n / world
        pair_bytes = (40.0 + 4.0 * lh) * n_per_gpu            # pos 16 + list 4*L_half + force 24 per particle
        step_bytes = (128.0 + 4.0 * lh) * n_per_gpu
        t_pair = (pair_ms * 1e-3 / pair_launches) if pair_launches else None
        roof = {"bound": "hbm", "kernel": "k_pair_forces", "achieved": (pair_bytes / t_pair / 1e9) if t_pair else None, "peak": peak,
                "unit": "GB/s", "frac": (pair_bytes / t_pair / 1e9 / peak) if t_pair else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this workload, from the committed
                # `ncu --set full` capture (profiles/r1g_ncu_full_raw.csv: 293.9 MB + 20.8 MB); only quoted for the profiled case
                "traffic": 314.7e6 if (world == 1 and a.n_side == 100) else None, "traffic_source": "profiles/r1g_ncu_full_raw.csv",
                "binding_roof": "shared-memory wavefronts (56.5 M per launch, 52 % bank conflicts of the two random 16-byte gathers per pair); not HBM",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": pair_bytes, "kernel_ms": (t_pair * 1e3) if t_pair else None, "launches_timed": pair_launches,
                "kernel_share_of_step": (pair_ms * 1e-3 / t_dev) if t_dev else None,
                "step_achieved": step_bytes * steps_per_s / 1e9, "step_frac": step_bytes * steps_per_s / 1e9 / peak,
                "step_algorithmic_bytes": step_bytes}
        line = {"metric": "md_steps_per_s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "roofline": roof, "clocks": clocks, "gpu_launches": cn["launches"],
                "pair_interactions_per_s": interacting * steps_per_s, "interacting_pairs_per_step": interacting,
                "list_pairs_per_particle": cn["list_entries"] / 2.0 / n, "ns_per_day_at_dt_ps": steps_per_s * DT * 86.4,
                "rebuilds": cn["rebuilds"], "reaction_passes": cn["reaction_passes"], "reaction_events": cn["reaction_events"],
                "new_bonds_total": nbonds_new, "temperature": float(kin[1]), "wall_s": t_wall, "equil_steps": a.equil,
                "ghost_beads_total": cn["ghosts"], "pair_threads": e.get_option("pair_threads"), "pair_grid": e.get_option("pair_grid"),
                "pair_nv": e.get_option("pair_nv"), "pair_smem": e.get_option("pair_smem"), "home_max": e.get_option("home_max"), "tile_max": e.get_option("tile_max"),
                "buckets_s": {k: v for k, v in tm.items() if v > 0} if e.get_option("timers") else None}

    # e2e: public API with HOST buffers -- a fresh engine (one per rank when N > 1) restarted from the host snapshot:
    # upload of the particle state, run in chunks with observables read back, final state download
    if not a.no_e2e:
        chunk = INTERVAL
        barrier()
        # engine creation and the NCCL rendezvous are one-off setup, not data movement: outside the timed region
        e2 = _E(snap["box"], RC, SKIN, seed=SEED, device=device)
        if world > 1:
            e2.join()
        barrier()
        t0 = time.perf_counter()
        e2.set_particles(snap["ids"], snap["type"], snap["pos"], snap["mass"], vel=snap["vel"], state=snap["state"], res_id=snap["resid"])
        t_a = time.perf_counter() - t0
        h2 = synthetic.setup_reactive_melt(e2, snap, rc=RC, dt=DT, kT=KT, gamma=GAMMA, interval=INTERVAL, p_accept=P_ACCEPT)
        t_b = time.perf_counter() - t0
        restore_into(e2, snap, h2)
        e2.reaction_general(1, INTERVAL, 1, 0)
        t_up = time.perf_counter() - t0
        if os.environ.get("CLB_BENCH_VERBOSE"):
            print("e2e upload: set_particles %.3fs, force field + lists %.3fs, restore %.3fs" % (t_a, t_b - t_a, t_up - t_b), file=sys.stderr)
        done = 0
        obs = []
        while done < a.steps:
            m = min(chunk, a.steps - done)
            e2.run(m); done += m
            obs.append((e2.kinetics()[1], e2.energy(h2["nb"]), e2.energy(h2["bonds"]), e2.energy(h2["angles"]), e2.energy(h2["react_bonds"])))
        t_run = time.perf_counter() - t0 - t_up
        out = e2.get_particles(fields=("pos", "vel", "type", "state", "image"))
        barrier()
        t_e2e = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt[0])
        h2d = n * (8 + 4 + 24 + 24 + 8 + 4 + 4) + snap["bonds"].size * 8 + snap["angles_now"].size * 8 + snap["excl_now"].size * 8 + 3 * 1500 * 8
        d2h = n * (24 + 24 + 4 + 4 + 12) + len(obs) * 5 * 8
        if rank == 0:
            line["e2e"] = {"value": a.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
                           "seconds": t_e2e, "upload_s": t_up, "run_s": t_run, "download_s": t_e2e - t_up - t_run,
                           "what": "(engine create%s outside) full state upload + run in %d-step chunks with T/Epot read back per chunk + final state download (N > 1: every rank uploads and downloads the full state)" % (" + NCCL join" if world > 1 else "", chunk)}
        e2.close()
    elif rank == 0:
        line["e2e"] = None

    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(snap, target_seconds=a.cpu_seconds)
    if rank == 0:
        print(json.dumps(line))
    e.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- MD steps/s and pair-interactions/s of chemlab's reactive-MD hot path on B200.

Default workload (BASELINE.json configs[1], SURVEY 8d "config 2"): synthetic 1,000,000-bead LJ reactive melt of
A-L-A trimers, rho=0.8442, rc=2.5, skin=0.3, dt=0.005, kT=1, gamma=1, pair potential = LJ tabulated on
1500 rows (dr=0.002, linear interpolation), harmonic bonds K=30 r0=0.97, harmonic angle 180 deg,
step-growth reaction A(1,2)+A(1,2)->A(1):A(1), cutoff 1.2, interval 200, p=0.05, nearest partner.
The start state is the periodic replication of an EQUILIBRATED 8000-bead tile (chemlab_b200/data/melt_tile_20.npz),
identical for the GPU arm, its cpu_baseline leg and the `--impl reference` arm.  `--workload c3|c4|c5` selects the
multi-table configurations (replicated hyperbranched / dacron / rim135 systems, chemlab_b200/synthetic.py).

A "step" is one Velocity-Verlet step of the whole system (neighbour rebuilds and reaction passes included at their
natural frequency).  One JSON line is printed by rank 0:
  value   : steps/s over EXACTLY --steps steps, device-timed (CUDA events on the engine's stream), state resident in HBM
  e2e     : steps/s through the public C-ABI with (pinned) HOST buffers: upload of the full particle state, the same
            number of steps with the observables read back per chunk, final state download -- all inside the timed region
  e2e_steady : the same engine, a further chunk of steps + observables with no state round trip
  roofline: pair-force kernel, algorithmic bytes (40 + 4*L_half per particle per launch) over its CUDA-event
            duration, against MEASURED_PEAKS.json hbm_gbs
  reaction_pass_ms / steps_per_s_amortised : one reaction pass timed on its own; steps/s with one pass per `interval` steps
  parity  : GPU engine vs the fp64 CPU oracle ON THE BENCHMARKED STATE: Verlet pair set, forces, per-interaction energies,
            then one reaction pass on both (events, bond lists, types, states)
  cpu_baseline: the fp64 CPU restatement (oracle/) of the same path timed on the host cores (all of them)
`--impl reference` times that CPU restatement alone for --warmup + --steps steps (the reference's own ESPResSo++ cannot
be built here: SURVEY 8c) and prints the same line with "impl": "reference".
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


def l_half(rc=RC, skin=SKIN, rho=RHO):
    return (2.0 * np.pi / 3.0) * (rc + skin) ** 3 * rho


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def host_threads():
    """Cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arms ignore it on purpose."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


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
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1)] or [r for (_, r) in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload C2
class WorkloadC2:
    name = "c2"
    rc, skin, dt, kT, gamma, rho, interval, p_accept, nearest = RC, SKIN, DT, KT, GAMMA, RHO, INTERVAL, P_ACCEPT, True

    def __init__(self, n_side):
        self.n_side = n_side
        self.replicated = (n_side % 20 == 0)
        self.description = ("C2 synthetic %d-bead LJ reactive trimer melt (tabulated LJ 1500 rows + harmonic bonds/angles + step-growth reaction; %s)"
                            % (n_side ** 3, "replicated equilibrated 8000-bead tile" if self.replicated else "jittered lattice start"))

    def system(self):
        from chemlab_b200 import synthetic
        if self.replicated:
            return synthetic.replicated_melt(self.n_side)
        return synthetic.trimer_melt(self.n_side, rho=RHO, seed=12345, kT=KT)

    @property
    def n(self):
        return self.n_side ** 3

    def setup(self, api, sysd):
        from chemlab_b200 import synthetic
        return synthetic.setup_reactive_melt(api, sysd, rc=RC, dt=DT, kT=KT, gamma=GAMMA, interval=INTERVAL, p_accept=P_ACCEPT)

    def l_half(self):
        return l_half()

    def config(self, gpus):
        return {"workload": self.description, "n_beads": self.n_side ** 3, "rho": RHO, "rc": RC, "skin": SKIN, "dt": DT, "kT": KT, "gamma": GAMMA,
                "reaction_interval": INTERVAL, "p_accept": P_ACCEPT, "nearest": True,
                "l2": "per-step working set (pos+vel+force+lists ~0.4 GB at 1M beads) exceeds the 126 MB L2; no explicit flush",
                "parallelism": "slab%d" % gpus if gpus > 1 else "single"}


def make_workload(a):
    if a.workload == "c2":
        return WorkloadC2(a.n_side)
    from chemlab_b200 import synthetic
    return synthetic.make_workload(a.workload, a.scale, example_root=os.path.join(ROOT, "tests", "golden"))


# ------------------------------------------------------------------------------------------------ state hand-over
def snapshot(e, sysd, h, pinned=False):
    """Host copy of the full dynamic state of an engine (the e2e leg, the parity oracle and the CPU baseline restart from it)."""
    st = e.get_particles(fields=("pos", "vel", "type", "state", "mass", "image"))
    s = dict(sysd)
    s["pos"] = st["pos"] + st["image"] * sysd["box"]
    s["vel"] = st["vel"]; s["type"] = st["type"]; s["state"] = st["state"]; s["mass"] = st["mass"]
    s["lists_now"] = {k: e.list_get(v, a) for k, (v, a) in h["lists"].items()}
    s["excl_now"] = e.get_exclusions()
    s["step"] = e.step()
    if pinned:
        import torch
        for k in ("pos", "vel", "mass", "type", "state", "resid", "ids") + (("q",) if "q" in s else ()):
            t = torch.from_numpy(np.ascontiguousarray(s[k])).pin_memory()
            s[k] = t.numpy(); s.setdefault("_keep", []).append(t)
    return s


def restore_into(api, s, h):
    """Tuples created by reactions since the start (appended behind the initial ones) and the current exclusions."""
    for k, (lst, ar) in h["lists"].items():
        now, init = s["lists_now"].get(k), h["initial"].get(k, 0)
        if now is not None and len(now) > init:
            api.list_add(lst, now[init:])
    api.set_exclusions(s["excl_now"])


def upload(api, s):
    api.set_particles(s["ids"], s["type"], s["pos"], s["mass"], vel=s["vel"], q=s.get("q"), state=s["state"], res_id=s["resid"])


def oracle_from(wl, s, threads):
    """The CPU oracle (through its Engine-shaped adapter) holding state s with the workload's force field and lists."""
    from oracle.engine_adapter import OracleEngine
    o = OracleEngine(s["box"], wl.rc, wl.skin, seed=SEED)
    upload(o, s)
    o.set_threads(threads)
    h = wl.setup(o, s)
    restore_into(o, s, h)
    o.set_option("step", s.get("step", 0))
    return o, h


def cpu_run(wl, s, steps, warmup, threads, reactions=True):
    """CPU restatement (oracle/) on state s: `warmup` untimed steps (the first includes the list build), then EXACTLY `steps`
    timed steps in up to 3 segments (median and spread of the segment rates are reported)."""
    o, h = oracle_from(wl, s, threads)
    if reactions:
        o.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
    t0 = time.perf_counter(); o.run(max(1, warmup)); t_warm = time.perf_counter() - t0
    nseg = 3 if steps >= 6 else 1
    seg = [steps // nseg + (1 if k < steps % nseg else 0) for k in range(nseg)]
    rates, total = [], 0.0
    for m in seg:
        t0 = time.perf_counter(); o.run(m); dt = time.perf_counter() - t0
        total += dt; rates.append(m / dt)
    return {"value": steps / total, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": "%d MD steps (after %d warm-up steps, %.2fs incl. the list build) of the same %d-bead workload from the same state (fp64 C restatement, OpenMP x%d, pinned threads; %s)"
                      % (steps, max(1, warmup), t_warm, s["n"], threads, "reaction passes at their interval" if reactions else "no reaction pass, like the GPU arm's timed window"),
            "seconds": total, "steps": steps, "segment_rates": rates, "median_rate": float(np.median(rates)),
            "spread": (max(rates) - min(rates)) / float(np.median(rates)) if len(rates) > 1 else 0.0}


def cpu_leg_subprocess(a, snap, steps, warmup, reactions=True):
    """The cpu_baseline leg in its own process (same interpreter, same workload flags, `--impl cpu_leg`): pinned OpenMP threads
    there, none of the pinning in the process that drives the GPU."""
    import pickle
    import subprocess
    import tempfile
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    fd, path = tempfile.mkstemp(suffix=".pkl", prefix="clb_snap_", dir=d)
    try:
        with os.fdopen(fd, "wb") as f:
            pickle.dump({k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in snap.items() if k != "_keep"}, f, protocol=4)
        env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "cpu_leg", "--state", path, "--steps", str(steps), "--warmup", str(warmup),
               "--workload", a.workload, "--scale", str(a.scale), "--n_side", str(a.n_side)] + ([] if reactions else ["--no_reactions"])
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1800)
        for ln in reversed(r.stdout.splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": "cpu_leg failed: " + (r.stderr or r.stdout)[-400:]}
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass


def bounded_cpu_steps(wl, n, threads, seconds):
    """Steps that fit the CPU time budget (about 1.5 us per bead-step per core for C2 on the round-1 boxes)."""
    est = 1.6e-6 * n * (wl.l_half() / 38.8) / max(1, min(threads, 32)) + 1e-4
    return int(max(3, min(200, seconds / est)))


# ------------------------------------------------------------------------------------------------ parity
def sorted_pair_keys(p):
    p = np.asarray(p)
    k = (p[:, 0].astype(np.uint64) << np.uint64(32)) | p[:, 1].astype(np.uint64)
    k.sort()
    return k


def rel_force_err(f, fref):
    d = np.linalg.norm(f - fref, axis=1)
    nr = np.linalg.norm(fref, axis=1)
    rms = np.sqrt((nr ** 2).mean())
    return float((d / np.maximum(nr, rms)).max())


def parity_block(e, wl, sysd, h, rank, threads, do_reaction=True, barrier=None):
    """GPU engine vs CPU oracle on the engine's CURRENT (benchmarked) state.  Collective on a multi-rank engine: every rank
    makes the engine calls, rank 0 runs the oracle and compares.  Returns (dict on rank 0 | None, reaction_pass_seconds)."""
    t_start = time.perf_counter()
    e.decompose()                                    # lists of the CURRENT positions (the run's last rebuild is a few steps old)
    snap = snapshot(e, sysd, h)
    pe = e.pairs()                                   # canonical sorted (id_a < id_b) rows
    e.compute_forces()
    fe = e.get_particles(fields=("force",))["force"]
    en_e = {k: e.energy(v) for k, v in h["energies"].items()}
    out = None
    o = None
    if rank == 0:
        o, ho = oracle_from(wl, snap, threads)
        ko = sorted_pair_keys(o.pairs_raw())
        ke = (pe[:, 0].astype(np.uint64) << np.uint64(32)) | pe[:, 1].astype(np.uint64)
        pairs_equal = bool(len(ko) == len(ke) and np.array_equal(ko, ke))
        pair_diff = None
        if not pairs_equal:                         # diagnostics: which side lists what the other does not
            miss = np.setdiff1d(ko, ke, assume_unique=False); extra = np.setdiff1d(ke, ko, assume_unique=False)
            dup = int(len(ke) - len(np.unique(ke)))
            pair_diff = {"oracle_pairs": int(len(ko)), "missing_in_engine": int(len(miss)), "extra_in_engine": int(len(extra)), "duplicates_in_engine": dup,
                         "first_missing": [[int(k >> np.uint64(32)), int(k & np.uint64(0xffffffff))] for k in miss[:4]],
                         "first_extra": [[int(k >> np.uint64(32)), int(k & np.uint64(0xffffffff))] for k in extra[:4]]}
        o.compute_forces()
        fo = o.get_particles(fields=("force",))["force"]
        en_o = {k: o.energy(v) for k, v in ho["energies"].items()}
        e_err = max(abs(en_e[k] - en_o[k]) / max(abs(en_o[k]), 1e-300) for k in en_o if en_o[k] != 0.0 or en_e[k] != 0.0) if en_o else 0.0
        out = {"state": "benchmarked state after the timed steps (%d beads, step %d)" % (snap["n"], snap["step"]),
               "pairs": int(len(ke)), "pairs_equal": pairs_equal, "pair_diff": pair_diff, "force_rel_err": rel_force_err(fe, fo), "force_bar": 1e-6,
               "energy_rel_err": float(e_err), "energy_bar": 1e-8, "energies": {k: en_e[k] for k in en_e}}
    t_react = None
    if do_reaction:
        e.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
        if barrier is not None:
            barrier()                                 # rank 0 has just run the oracle: the other ranks must not time their wait for it
        t0 = time.perf_counter()
        nev = e.react_now()
        t_react_first = time.perf_counter() - t0
        st = e.get_particles(fields=("type", "state", "mass"))
        lists_e = {k: e.list_get(v, a) for k, (v, a) in h["lists"].items()}
        ex_e = e.get_exclusions()
        if rank == 0:
            o.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
            nev_o = o.react()
            so = o.get_particles(fields=("type", "state", "mass"))

            def canon(a):
                a = np.asarray(a, np.int64)
                return a[np.lexsort(a.T[::-1])] if len(a) else a
            lists_equal = all(np.array_equal(canon(lists_e[k]), canon(o.list_get(ho["lists"][k][0], ho["lists"][k][1]))) for k in lists_e)
            out["reaction"] = {"events": int(nev), "events_oracle": int(nev_o), "bond_lists_equal": bool(lists_equal),
                               "types_equal": bool(np.array_equal(st["type"], so["type"])), "states_equal": bool(np.array_equal(st["state"], so["state"])),
                               "masses_equal": bool(np.array_equal(st["mass"], so["mass"])),
                               "exclusions_equal": bool(np.array_equal(canon(ex_e), canon(o.get_exclusions())))}
    if do_reaction:
        # the pass compared above is the first one of this engine when the timed window held none (buffers grow to their working
        # size, the replicated tile yields 3x the events of any later pass): a second pass is the one that is timed
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        e.react_now()
        t_react = time.perf_counter() - t0
        if rank == 0:
            out["reaction"]["first_pass_ms"] = 1e3 * t_react_first
    if rank == 0:
        r = out.get("reaction", {})
        out["green"] = bool(out["pairs_equal"] and out["force_rel_err"] < 1e-6 and out["energy_rel_err"] < 1e-8 and
                            all(v for k, v in r.items() if k.endswith("_equal")) and r.get("events", 0) == r.get("events_oracle", 0))
        out["oracle"] = "oracle/chemlab_oracle.c (fp64 CPU restatement; PARITY UNPINNED against ESPResSo++ itself: SURVEY 8c, REFERENCE_UNVERIFIED.md)"
        out["seconds"] = time.perf_counter() - t_start
    return out, t_react


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--equil", type=int, default=0, help="untimed extra steps before warm-up (reactions off); the replicated tile needs none")
    ap.add_argument("--n_side", type=int, default=100, help="C2: beads per box edge (100 -> 1,000,000 beads; multiples of 20 use the equilibrated tile)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=int, default=0, help="c3/c4/c5: replications per box edge (0 = the BASELINE size)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--no_parity", action="store_true")
    ap.add_argument("--cpu_seconds", type=float, default=15.0)
    ap.add_argument("--option", action="append", default=[], help="engine option name=value")
    ap.add_argument("--state", default=None, help="(--impl cpu_leg) pickled host snapshot written by the GPU arm")
    ap.add_argument("--no_reactions", action="store_true", help="(--impl cpu_leg) MD steps only: the GPU arm's timed window held no reaction pass")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    threads = host_threads()
    # the CPU arms use every core this process may run on (set before libgomp loads).  Thread PINNING is for the processes that
    # only run the CPU restatement (`--impl reference`, `--impl cpu_leg`): in a process that drives a GPU, OMP_PROC_BIND pins the
    # main thread to the first core and every thread created later (NCCL proxy, CUDA workers) inherits that single-core mask --
    # measured on the 2-GPU box as a constant 8 ms (one scheduler slice) per NCCL round and 0.4 s reaction passes.
    os.environ["OMP_NUM_THREADS"] = str(threads)
    if a.impl in ("reference", "cpu_leg"):
        os.environ.setdefault("OMP_PROC_BIND", "close")
        os.environ.setdefault("OMP_PLACES", "cores")
    else:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")      # the parity oracle's idle workers must not spin beside the CUDA host thread
    wl = make_workload(a)
    config = wl.config(a.gpus)

    if a.impl == "cpu_leg":
        # child of the GPU arm: the cpu_baseline leg on the pickled snapshot, alone in its process with pinned threads
        import pickle
        with open(a.state, "rb") as f:
            snap = pickle.load(f)
        print(json.dumps(cpu_run(wl, snap, a.steps, a.warmup, threads, reactions=not a.no_reactions)))
        return 0

    if a.impl == "reference":
        if rank != 0:
            return 0
        sysd = wl.system()
        s = dict(sysd); s["lists_now"] = {}; s["excl_now"] = sysd["exclusions"]; s["step"] = 0
        steps = a.steps
        cap = bounded_cpu_steps(wl, sysd["n"], threads, 120.0)
        note_cap = ""
        if steps > cap:
            note_cap = "; --steps %d capped to %d so that the CPU run ends within a few minutes" % (steps, cap)
            steps = cap
        cb = cpu_run(wl, s, steps, a.warmup, threads)
        line = {"impl": "reference", "metric": "md_steps_per_s", "value": cb["value"], "unit": "steps/s", "n_gpus": a.gpus, "steps": steps,
                "warmup": max(1, a.warmup), "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "CPU restatement of the reference algorithm (oracle/), NOT ESPResSo++ itself: cgchemlab/espressopp is not vendored and cannot be built here (SURVEY 8c); every step is a full MD step of the whole workload" + note_cap}
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
        if not dist.is_initialized():         # the recorded chemlab driver of the c3/c4/c5 workloads may have joined the group already
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = local

    sysd = wl.system()
    n = sysd["n"]
    from chemlab_b200 import Engine as _E

    def new_engine():
        e_ = _E(sysd["box"], wl.rc, wl.skin, seed=SEED, device=device)
        for kv in a.option:
            k, v = kv.split("=")
            e_.set_option(k, float(v))
        if world > 1:
            e_.join()          # one engine per rank = one slab; torch.distributed only carries the NCCL id
        return e_
    e = new_engine()
    upload(e, sysd)
    h = wl.setup(e, sysd)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    if a.equil > 0:
        e.reaction_general(0, wl.interval, 1 if wl.nearest else 0, 0)
        e.run(a.equil)
    e.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
    if a.warmup > 0:
        e.run(a.warmup)
    # host snapshot of the warmed-up state: the e2e leg and the CPU baseline restart from it (collective read-back)
    snap = snapshot(e, sysd, h, pinned=True) if not (a.no_e2e and (a.no_cpu_baseline or world > 1)) else None

    e.reset_timers()
    e.set_option("pair_event_timing", 1)
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    t_wall0 = time.perf_counter()
    e.run(a.steps)
    barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    if rank == 0 and t_wall < 0.3:
        time.sleep(0.1)
    clocks = sampler.stop(t_wall0 - 0.02, t_wall1 + 0.05) if rank == 0 else None
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
    rl_names = [k for k, (v, ar) in h["lists"].items() if v in h.get("react_lists", [h.get("react_list")])]
    nbonds_new = sum(e.list_size(h["lists"][k][0]) - h["initial"].get(k, 0) for k in rl_names)

    # parity on the benchmarked state (+ one reaction pass on both sides, which is also the timed reaction pass)
    parity, t_react = (None, None)
    if not a.no_parity:
        parity, t_react = parity_block(e, wl, sysd, h, rank, threads, barrier=barrier)
    else:
        e.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
        e.react_now()                     # first pass of this engine (buffers grow, 3x the events): untimed
        barrier()
        t0 = time.perf_counter(); e.react_now(); t_react = time.perf_counter() - t0
    if dist is not None and t_react is not None:
        tt = torch.tensor([t_react], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_react = float(tt[0])

    line = None
    if rank == 0:
        steps_per_s = a.steps / t_dev
        peak, peak_src = measured_peaks()
        lh = wl.l_half()
        n_per_gpu = n / world
        pair_bytes = (40.0 + 4.0 * lh) * n_per_gpu            # pos 16 + list 4*L_half + force 24 per particle
        step_bytes = (128.0 + 4.0 * lh) * n_per_gpu
        t_pair = (pair_ms * 1e-3 / pair_launches) if pair_launches else None
        traffic, traffic_src = profiled_traffic(wl, world, n)
        roof = {"bound": "hbm", "kernel": "k_pair_forces", "achieved": (pair_bytes / t_pair / 1e9) if t_pair else None, "peak": peak,
                "unit": "GB/s", "frac": (pair_bytes / t_pair / 1e9 / peak) if t_pair else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "binding_roof": "issue slots (55 instructions per listed pair around 16 fp64 operations, 60 % issue utilisation at 25 warps/SM) and shared-memory gather wavefronts; not HBM (DESIGN.md 3.1)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": pair_bytes, "kernel_ms": (t_pair * 1e3) if t_pair else None, "launches_timed": pair_launches,
                "kernel_share_of_step": (pair_ms * 1e-3 / t_dev) if t_dev else None,
                "step_achieved": step_bytes * steps_per_s / 1e9, "step_frac": step_bytes * steps_per_s / 1e9 / peak,
                "step_algorithmic_bytes": step_bytes}
        npass_win = cn["reaction_passes"]
        amort = None
        if t_react is not None:
            t_md = max(t_dev - npass_win * t_react, 1e-9)
            amort = a.steps / (t_md + (a.steps / wl.interval) * t_react)
        line = {"metric": "md_steps_per_s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "roofline": roof, "clocks": clocks, "gpu_launches": cn["launches"],
                "pair_interactions_per_s": interacting * steps_per_s, "interacting_pairs_per_step": interacting,
                "list_pairs_per_particle": cn["list_entries"] / 2.0 / n, "ns_per_day_at_dt_ps": steps_per_s * wl.dt * 86.4,
                "rebuilds": cn["rebuilds"], "reaction_passes": npass_win, "reaction_events": cn["reaction_events"],
                "reaction_pass_ms": (1e3 * t_react) if t_react is not None else None,
                "steps_per_s_amortised": amort,
                "amortised_note": "steps/s with exactly one reaction pass per %d steps (the timed window held %d); reaction_pass_ms is one pass timed on its own (host wall clock, synchronous call)" % (wl.interval, npass_win),
                "new_bonds_total": nbonds_new, "temperature": float(kin[1]), "wall_s": t_wall, "equil_steps": a.equil,
                "ghost_beads_total": cn["ghosts"], "parity": parity}

    # the first engine is done (timed run, parity): its device blocks go to the engine's block cache and are reused below
    pair_opts = {k: e.get_option(k) for k in ("pair_threads", "pair_grid", "pair_nv", "pair_smem", "home_max", "tile_max", "pair_kernel", "pair_rep",
                                               "pair_table_rows", "pair_tables_resident", "pair_tables_resident_weight", "block_cells", "timers", "comm_peer")}
    if rank == 0:
        line.update({k: pair_opts[k] for k in pair_opts if k != "timers"})
        line["buckets_s"] = {k: v for k, v in tm.items() if v > 0} if pair_opts["timers"] else None
    e.close()
    # e2e: public API with HOST buffers -- a fresh engine (one per rank when N > 1) restarted from the pinned host snapshot:
    # upload of the particle state, the same number of steps with observables read back per chunk, final state download
    if not a.no_e2e:
        chunk = wl.interval
        barrier()
        # engine creation and the NCCL rendezvous are one-off setup, not data movement: outside the timed region
        e2 = new_engine()
        outbuf = {k: torch.empty(sh, dtype=dt_, pin_memory=True).numpy() for k, sh, dt_ in
                  (("pos", (n, 3), torch.float64), ("vel", (n, 3), torch.float64), ("image", (n, 3), torch.int32), ("type", (n,), torch.int32), ("state", (n,), torch.int32))}
        barrier()
        t0 = time.perf_counter()
        upload(e2, snap)
        t_a = time.perf_counter() - t0
        h2 = wl.setup(e2, snap)
        t_b = time.perf_counter() - t0
        restore_into(e2, snap, h2)
        e2.set_option("step", snap["step"])
        e2.reaction_general(1, wl.interval, 1 if wl.nearest else 0, 0)
        t_up = time.perf_counter() - t0
        done = 0
        obs = []

        def observe():
            obs.append((e2.kinetics()[1],) + tuple(e2.energy(v) for v in h2["energies"].values()))
        while done < a.steps:
            m = min(chunk, a.steps - done)
            e2.run(m); done += m
            observe()
        t_run = time.perf_counter() - t0 - t_up
        nobs = len(obs)
        e2.get_particles(fields=("pos", "vel", "type", "state", "image"), out=outbuf)
        barrier()
        t_e2e = time.perf_counter() - t0
        # steady state: a further chunk of steps + observables, no state round trip
        barrier()
        t1 = time.perf_counter()
        done = 0
        while done < a.steps:
            m = min(chunk, a.steps - done)
            e2.run(m); done += m
            observe()
        barrier()
        t_steady = time.perf_counter() - t1
        if os.environ.get("CLB_BENCH_VERBOSE"):
            print("e2e upload: set_particles %.4fs, force field + lists %.4fs, restore %.4fs; run %.4fs; download %.4fs; steady %.4fs"
                  % (t_a, t_b - t_a, t_up - t_b, t_run, t_e2e - t_up - t_run, t_steady), file=sys.stderr)
        if dist is not None:
            tt = torch.tensor([t_e2e, t_steady], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e, t_steady = float(tt[0]), float(tt[1])
        nlist = sum(v.size for v in snap["lists_now"].values())
        h2d = n * (8 + 4 + 24 + 24 + 8 + 4 + 4) + nlist * 8 + snap["excl_now"].size * 8 + h2.get("table_bytes", 3 * 1500 * 8)
        d2h = n * (24 + 24 + 4 + 4 + 12) + nobs * (1 + len(h2["energies"])) * 8
        if rank == 0:
            line["e2e"] = {"value": a.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
                           "seconds": t_e2e, "upload_s": t_up, "run_s": t_run, "download_s": t_e2e - t_up - t_run,
                           "what": "(engine create%s outside) full state upload from pinned host memory + force field, lists, exclusions + %d steps in %d-step chunks with T/Epot read back per chunk + final state download (N > 1: every rank uploads and downloads the full state)"
                                   % (" + NCCL join" if world > 1 else "", a.steps, chunk)}
            line["e2e_steady"] = {"value": a.steps / t_steady, "unit": "steps/s", "seconds": t_steady,
                                  "what": "the same engine, a further %d steps in %d-step chunks with T/Epot read back per chunk, no state round trip" % (a.steps, chunk)}
        e2.close()
    elif rank == 0:
        line["e2e"] = None

    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        ksteps = bounded_cpu_steps(wl, n, threads, a.cpu_seconds)
        line["cpu_baseline"] = cpu_leg_subprocess(a, snap, ksteps, 3, reactions=cn["reaction_passes"] > 0)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def profiled_traffic(wl, world, n=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the pair kernel on this workload, from the committed
    `ncu --set full` capture named in profiles/traffic.json (a profile cannot be taken inside a timed run)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        ent = t.get("%s_n%d" % (wl.name, world))
        if ent and (n is None or ent.get("n_beads") in (None, n)):
            return float(ent["bytes_per_launch"]), "from profile: " + ent["source"]
    except Exception:
        pass
    return None, "no ncu capture committed for this workload / rank count"


if __name__ == "__main__":
    sys.exit(main())

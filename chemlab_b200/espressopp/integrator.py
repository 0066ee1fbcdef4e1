"""espressopp.integrator.*: VelocityVerlet and the extensions chemlab attaches to it
(src/start_simulation.py:165-167,330-354,395-444,566-569,728-797; src/chemlab/reaction_setup.py:81-163,417-427;
src/chemlab/reaction_post_process.py:76-115,380-426)."""
import math

import numpy as np

from ._context import not_in_scope


class VelocityVerlet:
    """integrator.VelocityVerlet(system): .dt .step .run(n) .addExtension(ext) .getTimers()."""
    def __init__(self, system):
        self._system = system
        self._ctx = system._ctx
        self._ctx.integrator = self
        self.dt = 0.001
        self._extensions = []
        self._step_offset = 0
        self._wall = 0.0

    # -- extensions
    def addExtension(self, ext):
        self._extensions.append(ext)
        if hasattr(ext, "_connect"):
            ext._connect(self)

    def getNumberOfExtensions(self):
        return len(self._extensions)

    def getExtension(self, k):
        return self._extensions[k]

    def _attach(self, e):
        for ext in self._extensions:
            if hasattr(ext, "_attach"):
                ext._attach(e)

    @property
    def step(self):
        e = self._ctx.engine
        return (e.step() if e is not None else 0) + self._step_offset

    @step.setter
    def step(self, v):
        e = self._ctx.engine
        self._step_offset = int(v) - (e.step() if e is not None else 0)

    def run(self, nsteps):
        import time
        e = self._ctx.require_engine()
        e.set_dt(self.dt)
        thermo = [x for x in self._extensions if isinstance(x, LangevinThermostat)]
        if thermo:
            thermo[-1]._push(e)
        else:
            e.set_langevin(0, 1.0, 1.0)
        caps = [x for x in self._extensions if isinstance(x, CapForce)]
        e.set_cap_force(caps[-1].capForce if caps else -1.0)
        react = [x for x in self._extensions if isinstance(x, ChemicalReaction) and x._connected]
        for x in self._extensions:
            if isinstance(x, ChemicalReaction):
                x._push(e, enabled=x._connected)
        if not react:
            e.reaction_general(0, 1, 1, 0)
        hosted = [x for x in self._extensions if hasattr(x, "_host_interval")]   # ExtAnalyze, ATRPActivator
        t0 = time.time()
        done = 0
        nsteps = int(nsteps)
        while done < nsteps:
            chunk = nsteps - done
            now = self.step
            for x in hosted:
                k = x._host_interval()
                if k > 0:
                    chunk = min(chunk, k - (now % k))
            # one integrator.run of the reference = one run entry (force recalculation + heat-up), however many host actions
            (e.run if done == 0 else e.run_continue)(chunk)
            done += chunk
            now += chunk
            for x in hosted:
                k = x._host_interval()
                if k > 0 and now % k == 0:
                    x._host_action(self)
        self._wall += time.time() - t0

    def getTimers(self):
        """One list of (name, seconds) pairs per rank, the structure src/tools.py:51-79 consumes: 'timeRun' (skipped there), one
        'f<i>' per registered interaction (label = system.getNameOfInteraction(i)), communication, integration, resort and
        reaction buckets.  The engine times pair and bonded forces as two buckets (clb_timers): the pair bucket is spread evenly
        over the Verlet-list interactions, the bonded bucket over the fixed-list interactions."""
        e = self._ctx.engine
        if e is None:
            return [[]]
        t, c = e.timers()
        g = lambda k: float(t.get(k, 0.0))
        inters = [i for i, _ in self._ctx.interactions]
        is_nb = [hasattr(i, "_vl") or type(i).__name__.startswith("VerletList") for i in inters]
        n_nb, n_b = max(1, sum(is_nb)), max(1, len(inters) - sum(is_nb))
        rows = [("timeRun", g("total"))]
        rows += [("f%d" % k, g("pair") / n_nb if nb else g("bonded") / n_b) for k, nb in enumerate(is_nb)]
        rows += [("timeComm1", g("comm")), ("timeInt1", 0.5 * g("integrate")), ("timeInt2", 0.5 * g("integrate")), ("timeResort", g("neighbour")),
                 ("timeReaction", g("reaction"))]
        return [rows]


class LangevinThermostat:
    """LangevinThermostat(system): .temperature (= T*kB) .gamma .add_valid_types(types): src/start_simulation.py:330-336."""
    def get_timers(self):
        return []

    def __init__(self, system):
        self._system = system
        self.temperature = 1.0
        self.gamma = 1.0
        self._types = []

    def add_valid_types(self, types):
        self._types.extend(int(t) for t in types)

    def _push(self, e):
        e.set_langevin(1, float(self.temperature), float(self.gamma), self._types)


class TopologyParticleProperties:
    """TopologyParticleProperties(type, mass, q, state | incr_state, lambda_adr): reaction_setup.py:137-163."""
    def __init__(self, type=None, mass=None, q=None, state=None, incr_state=None, lambda_adr=None, res_id=None, **kw):
        self.type, self.mass, self.q, self.state, self.incr_state = type, mass, q, state, incr_state


class PostProcessChangeProperty:
    """PostProcessChangeProperty().add_change_property(old_type, props): the reactant itself (nb_level 0)."""
    def __init__(self):
        self._rules = []

    def add_change_property(self, type_id, props, nb_level=0):
        self._rules.append((int(type_id), props, int(nb_level)))


class PostProcessChangeNeighboursProperty(PostProcessChangeProperty):
    """PostProcessChangeNeighboursProperty(tm).add_change_property(type, props, nb_level): reaction_post_process.py:76-115."""
    def __init__(self, topology_manager=None):
        super().__init__()


class _ReactionCutoff:
    def __init__(self, cutoff):
        self.cutoff = float(cutoff)
        self.min_cutoff = 0.0


class Reaction:
    """integrator.Reaction(type_1, type_2, delta_1, delta_2, min_state_*, max_state_*, rate, fpl, cutoff): reaction_setup.py:81-113."""
    def __init__(self, type_1, type_2, delta_1, delta_2, min_state_1, max_state_1, min_state_2, max_state_2, rate, fpl, cutoff, **kw):
        self.type_1, self.type_2, self.delta_1, self.delta_2 = int(type_1), int(type_2), int(delta_1), int(delta_2)
        self.min_state_1, self.max_state_1 = int(min_state_1), int(max_state_1)
        self.min_state_2, self.max_state_2 = int(min_state_2), int(max_state_2)
        self._rate = float(rate)
        self.fpl = fpl
        self._cutoff = _ReactionCutoff(cutoff)
        self.intramolecular = False
        self.intraresidual = False
        self.is_virtual = False
        self._active = True
        self._post = []          # (postprocess, which)
        self._h = None
        self._engine = None

    rate = property(lambda s: s._rate)

    @rate.setter
    def rate(self, v):            # r.rate = exp(-dE/kT) mid-run: src/start_simulation.py:785-796
        self._rate = float(v)
        if self._h is not None:
            self._engine.reaction_set_rate(self._h, self._rate)

    active = property(lambda s: s._active)

    @active.setter
    def active(self, v):
        self._active = bool(v)
        if self._h is not None:
            self._engine.reaction_set_active(self._h, int(self._active))

    def get_reaction_cutoff(self):
        return self._cutoff

    def set_reaction_cutoff(self, rc):
        raise NotImplementedError("ReactionCutoffRandom is outside the scope of the B200 engine (SURVEY E21)")

    def add_postprocess(self, pp, which="both"):
        self._post.append((pp, which))

    def _attach_more(self, e):
        pass

    def _attach(self, e):
        if self._h is not None:
            return
        self.fpl._attach(e)
        self._engine = e
        self._h = e.add_reaction(self.type_1, self.type_2, self.delta_1, self.delta_2, self.min_state_1, self.max_state_1,
                                 self.min_state_2, self.max_state_2, self._rate, self._cutoff.cutoff, self.fpl._h,
                                 min_cutoff=self._cutoff.min_cutoff, intramolecular=int(bool(self.intramolecular)),
                                 intraresidual=int(bool(self.intraresidual)), is_virtual=int(bool(self.is_virtual)),
                                 active=int(self._active))
        self._attach_more(e)
        side_of = {"type_1": 1, "type_2": 2, "both": 3, None: 3}
        for pp, which in self._post:
            if not isinstance(pp, PostProcessChangeProperty):
                raise NotImplementedError("post-process %s is outside the scope of the B200 engine" % type(pp).__name__)
            for old_type, p, lvl in pp._rules:
                mode, val = (1, int(p.state)) if (p.state is not None and lvl > 0) else ((2, int(p.incr_state)) if p.incr_state else (0, 0))
                e.reaction_add_change(self._h, side_of[which], lvl, old_type, int(p.type),
                                      new_mass=float(p.mass) if p.mass is not None else -1.0,
                                      new_q=float(p.q) if p.q is not None else float("nan"), state_mode=mode, state_value=val)


class RestrictReaction(Reaction):
    """integrator.RestrictReaction(...): a Reaction that forms bonds only between the particle pairs named with
    define_connection(b1, b2) -- one call per line of the group's `connectivity_map` (reaction_setup.py:74-75,115-126;
    examples/dacron/restrict/reaction.cfg:26 + connections.list).  `revert` is only set for dissociation reactions
    (:127-128), which stay outside the scope of the engine."""
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._connections = set()
        self.revert = False

    def define_connection(self, b1, b2):
        self._connections.add((min(int(b1), int(b2)), max(int(b1), int(b2))))
        if self._h is not None:
            self._engine.reaction_define_connections(self._h, sorted(self._connections))

    def _attach_more(self, e):
        if self.revert:
            raise NotImplementedError("RestrictReaction.revert (dissociation) is outside the scope of the B200 engine")
        e.reaction_define_connections(self._h, sorted(self._connections))


DissociationReaction = not_in_scope("integrator.DissociationReaction")
ReactionCutoffRandom = not_in_scope("integrator.ReactionCutoffRandom")


class ChemicalReaction:
    """integrator.ChemicalReaction(system, vl, storage, tm, interval): reaction_setup.py:417-427,506; driver :737,:777."""
    def __init__(self, system, vl, storage_, topology_manager, interval):
        self._system = system
        self.interval = int(interval)
        self.nearest_mode = False
        self.max_per_interval = 0
        self.pair_distances_filename = None
        self._reactions = []
        self._connected = False
        self._engine = None

    def add_reaction(self, r):
        self._reactions.append(r)
        if self._engine is not None:
            r._attach(self._engine)

    def _connect(self, integrator_):
        self._connected = True

    def disconnect(self):
        self._connected = False

    def _attach(self, e):
        self._engine = e
        for r in self._reactions:
            r._attach(e)

    def _push(self, e, enabled):
        self._attach(e)
        e.reaction_general(int(enabled), self.interval, int(bool(self.nearest_mode)), int(self.max_per_interval or 0))

    def get_timers(self):
        e = self._engine
        return [[("timeReact", float(e.timers()[0].get("reaction", 0.0)))]] if e is not None else []

    def get_reaction_counters(self):
        if self._engine is None:
            return [0] * len(self._reactions)
        return self._engine.reaction_counters(len(self._reactions)).tolist()

    def save_reaction_counters(self, filename):
        with open(filename, "w") as f:
            for k, c in enumerate(self.get_reaction_counters()):
                f.write("%d %d\n" % (k, c))

    def save_intra_inter_counter(self, filename):
        open(filename, "w").close()


class TopologyManager:
    """integrator.TopologyManager(system): src/start_simulation.py:211-212,395-444."""
    def __init__(self, system):
        self._system = system
        self._ctx = system._ctx
        self._ctx.topology_manager = self
        self._observed, self._triplets, self._quads = [], [], []
        self._initialized = False
        self._engine = None

    def observe_tuple(self, fpl):
        self._observed.append(fpl)
        if self._engine is not None:
            fpl._attach(self._engine); self._engine.topology_observe(fpl._h)

    def observe_triple(self, ftl):
        pass

    observe_quadruple = observe_triple

    def register_tuple(self, fpl, t1, t2):
        pass     # new bonds always go to the reaction's own list (fpl=); the type key is informational

    def register_triplet(self, ftl, t1, t2, t3):
        self._triplets.append((ftl, (int(t1), int(t2), int(t3))))
        if self._engine is not None:
            ftl._attach(self._engine); self._engine.topology_register(ftl._h, (t1, t2, t3))

    def register_quadruplet(self, fql, t1, t2, t3, t4):
        self._quads.append((fql, (int(t1), int(t2), int(t3), int(t4))))
        if self._engine is not None:
            fql._attach(self._engine); self._engine.topology_register(fql._h, (t1, t2, t3, t4))

    def initialize_topology(self):
        self._initialized = True
        if self._engine is not None:
            self._engine.topology_initialize()

    initialize = initialize_topology

    def _attach(self, e):
        if self._engine is not None:
            return
        self._engine = e
        for fpl in self._observed:
            fpl._attach(e); e.topology_observe(fpl._h)
        for ftl, t in self._triplets:
            ftl._attach(e); e.topology_register(ftl._h, t)
        for fql, t in self._quads:
            fql._attach(e); e.topology_register(fql._h, t)
        if self._initialized:
            e.topology_initialize()

    def _connect(self, integrator_):
        pass

    def get_timers(self):
        return []       # list per rank of (name, seconds) pairs (src/start_simulation.py:1040-1048); the graph lives in the engine's reaction bucket

    def get_fixed_pair_list(self, t1, t2):
        return None

    # ---- topology dumps written at the end of a run (src/start_simulation.py:1004-1006).  [EXT] TopologyManager::SaveTopologyToFile /
    # SaveResTopologyToFile / SaveResiduesListToFile define the formats; unverified here (REFERENCE_UNVERIFIED.md U24): one line
    # per node, "id: neighbour ids ..." (bond graph), "res_id: neighbour res_ids ..." (residue graph), "res_id: particle ids ...".
    def _graph(self):
        import collections
        adj = collections.defaultdict(set)
        for fpl in self._observed:
            for a, b in fpl.getAllBonds():
                adj[int(a)].add(int(b)); adj[int(b)].add(int(a))
        return adj

    def _res_ids(self):
        e = self._ctx.require_engine()
        g = e.get_particles(fields=("res_id",))
        return dict(zip(sorted(self._ctx.pid), (int(r) for r in g["res_id"])))

    def save_topology(self, filename):
        adj = self._graph()
        with open(filename, "w") as f:
            for a in sorted(adj):
                f.write("%d: %s\n" % (a, " ".join(str(b) for b in sorted(adj[a]))))

    def save_res_topology(self, filename):
        import collections
        res = self._res_ids()
        radj = collections.defaultdict(set)
        for a, nb in self._graph().items():
            for b in nb:
                if res[a] != res[b]:
                    radj[res[a]].add(res[b])
        with open(filename, "w") as f:
            for r in sorted(radj):
                f.write("%d: %s\n" % (r, " ".join(str(x) for x in sorted(radj[r]))))

    def save_residues(self, filename):
        import collections
        members = collections.defaultdict(list)
        for pid, r in self._res_ids().items():
            members[r].append(pid)
        with open(filename, "w") as f:
            for r in sorted(members):
                f.write("%d: %s\n" % (r, " ".join(str(x) for x in members[r])))


class ExtAnalyze:
    """integrator.ExtAnalyze(observable_or_monitor, interval): src/start_simulation.py:566-569."""
    def get_timers(self):
        return []

    def __init__(self, action, interval):
        self._action, self._interval = action, int(interval)

    def _host_interval(self):
        return self._interval

    def _host_action(self, integrator_):
        a = self._action
        (getattr(a, "perform_action", None) or getattr(a, "dump", None) or a.compute)()


class ATRPActivator:
    """integrator.ATRPActivator(system, interval, num_particles, ratio_activator, ratio_deactivator, delta_catalyst,
    k_activate, k_deactivate) + add_reactive_center: reaction_post_process.py:380-426 (config 1 only).

    The pass itself runs in the engine (clb_atrp_configure / clb_atrp_add_center / clb_atrp_now: candidate scan, counter-based
    random selection of num_particles centres, activation / deactivation draws, property change -- csrc/clb_react.cuh); this
    class holds the parameters, fires the pass every `interval` steps and writes the statistics file [EXT, U22]."""
    def get_timers(self):
        return []

    def __init__(self, system, interval, num_particles, ratio_activator, ratio_deactivator, delta_catalyst, k_activate, k_deactivate):
        self._ctx = system._ctx
        self.interval, self.num_particles = int(interval), int(num_particles)
        self.ratio_activator, self.ratio_deactivator = float(ratio_activator), float(ratio_deactivator)
        self.delta_catalyst, self.k_activate, self.k_deactivate = float(delta_catalyst), float(k_activate), float(k_deactivate)
        self.stats_filename = None
        self.select_from_all = 1
        self._centers = []
        self._pushed = None

    def add_reactive_center(self, type_id, state, is_activator, new_property, delta_state):
        self._centers.append((int(type_id), int(state), bool(is_activator), new_property, int(delta_state)))
        self._pushed = None

    def _host_interval(self):
        return self.interval

    def _push(self, e):
        key = (id(e), self.num_particles, self.ratio_activator, self.ratio_deactivator, self.delta_catalyst, self.k_activate, self.k_deactivate,
               len(self._centers))
        if key == self._pushed:
            return
        e.atrp_configure(self.num_particles, self.ratio_activator, self.ratio_deactivator, self.delta_catalyst, self.k_activate, self.k_deactivate)
        for t, s, needs_deactivator, prop, ds in self._centers:
            # flag "A" (is_activator=False in chemlab's call, reaction_post_process.py:411): a dormant end that meets the
            # ACTIVATOR catalyst, rate k_activate * ratio_activator; flag "DA": an active end that meets the DEACTIVATOR
            e.atrp_add_center(t, s, needs_deactivator, -1 if prop.type is None else int(prop.type), -1.0 if prop.mass is None else float(prop.mass),
                              float("nan") if prop.q is None else float(prop.q), ds)

    def _host_action(self, integrator_):
        e = self._ctx.require_engine()
        self._push(e)
        (n_act, n_deact), (self.ratio_activator, self.ratio_deactivator) = e.atrp_now()
        self._pushed = (id(e), self.num_particles, self.ratio_activator, self.ratio_deactivator, self.delta_catalyst, self.k_activate, self.k_deactivate,
                        len(self._centers))
        if self.stats_filename:
            with open(self.stats_filename, "a") as f:
                f.write("%d %d %d %.6f %.6f\n" % (integrator_.step, n_act, n_deact, self.ratio_activator, self.ratio_deactivator))


class CapForce:
    """integrator.CapForce(system, capForce): src/start_simulation.py:320-324 (--max_force).  Force vectors longer than capForce
    are scaled back to that length after every force evaluation, before the thermostat (clb_set_cap_force, U26)."""
    def get_timers(self):
        return []

    def __init__(self, system, capForce, **kw):
        self._system = system
        self.capForce = float(capForce)


for _n in ("StochasticVelocityRescaling", "BerendsenThermostat", "Isokinetic", "LangevinBarostat", "BerendsenBarostat",
           "FixedListDynamicResolution", "BasicDynamicResolution", "FixDistances", "PostProcessReleaseParticles",
           "PostProcessJoinParticles", "PostProcessRemoveNeighbourBond", "PostProcessChangePropertyByTopologyManager",
           "ChangeInRegion", "ChangeParticleType", "ReactionConstraintNeighbourState"):
    globals()[_n] = not_in_scope("integrator." + _n)
del _n

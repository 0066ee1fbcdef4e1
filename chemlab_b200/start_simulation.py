#!/usr/bin/env python
"""start_simulation.py -- chemlab's driver on the B200 engine:  python -m chemlab_b200.start_simulation @params

Same command line, same input files (`params` arg-file, GROMACS-like .top/.itp, .gro, INI reaction file, optional
hooks.py) and same output files as the reference driver (src/start_simulation.py), re-authored for Python 3 on
top of `chemlab_b200.espressopp`.  Stages follow the reference one to one:

  setup     :77-197   units, topology, coordinates, System / storage / integrator, exclusions, VerletList
  topology  :211-212  TopologyManager
  hooks     :214-228  hooks.py from the working directory (hook_init_reaction, hook_at_step, ...)
  reactions :241-272  reaction_parser + SetupReactions
  forces    :298-310  non-bonded, bonds, angles, dihedrals
  thermostat:326-376  Langevin only (others are outside the engine's scope and raise)
  observers :446-569  SystemMonitor CSV
  main loop :728-797  integrator.run(integrator_step) with reaction start/stop and the conversion stop criterion
  outputs   :800-1081 final .gro, bond/angle/dihedral lists, reaction counters, benchmark record
"""
import math
import os
import random
import sys
import time

import numpy as np

from . import espressopp
from .chemlab import app_args, files_io, gromacs_topology, reaction_parser, reaction_setup, tools


def _load_hooks(path="hooks.py"):
    """hooks.py is executed in the driver's namespace (reference: execfile, :220-228)."""
    hooks = {}
    if os.path.exists(path):
        ns = {"espressopp": espressopp, "__name__": "hooks"}
        # user hook files start with `import espressopp` (examples/*/hooks.py): inside this driver that name is this package's
        # surface, unless a real espressopp has already been imported by the caller
        import sys
        sys.modules.setdefault("espressopp", espressopp)
        for sub in ("analysis", "integrator", "interaction", "storage", "bc", "esutil", "io", "tools"):
            if hasattr(espressopp, sub):
                sys.modules.setdefault("espressopp." + sub, getattr(espressopp, sub))
        with open(path) as f:
            text = f.read()
        try:
            code = compile(text, path, "exec")
        except SyntaxError:               # the hooks shipped with the reference's examples are Python-2 user code:
            import re                     # print statements are rewritten, anything beyond that has to be ported by hand
            fixed = "\n".join(re.sub(r"^(\s*)print (?!\()(.*)$", r"\1print(\2)", line) for line in text.split("\n"))
            try:
                code = compile(fixed, path, "exec")
                print("Note: %s uses Python-2 print statements; they were rewritten as calls" % path)
            except SyntaxError as e:
                raise RuntimeError("%s is not valid Python 3 (%s, line %s): port the hook file (print statements, "
                                   "random.sample on sets, ...)" % (path, e.msg, e.lineno)) from e
        ns.setdefault("xrange", range)
        exec(code, ns)
        # the five hooks of the reference (:215-228) plus two names earlier versions of this driver accepted
        for name in ("hook_init_reaction", "hook_postsetup_reaction", "hook_at_step", "hook_before_sim", "hook_end",
                     "hook_postsetup_interaction", "hook_setup_interactions"):
            if callable(ns.get(name)):
                hooks[name] = ns[name]
        print("Loaded hooks: %s" % ", ".join(sorted(hooks)))
    return hooks


def _init_ranks():
    """`torchrun --nproc-per-node N -m chemlab_b200.start_simulation @params` replaces `mpirun -n N start_simulation.py @params`
    (examples/*/run_simulation.pbs): one process per GPU, every rank runs the same driver, rank 0 prints and writes files."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        # the engine itself needs the GPU (clb_create fails without one); a process group that the caller initialised beforehand
        # (e.g. gloo in the CPU tests of the driver's multi-rank host logic) is used as it is
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return dist.get_rank(), world


def main(argv=None):
    rank, world = _init_ranks()
    if rank == 0:
        return _main(argv, rank, world)
    out, sys.stdout = sys.stdout, open(os.devnull, "w")      # only rank 0 talks
    try:
        return _main(argv, rank, world)
    finally:
        sys.stdout.close()
        sys.stdout = out


def _main(argv, rank, world):
    args = app_args._args().parse_args(argv)
    prefix_dir = os.path.dirname(args.output_prefix)
    if prefix_dir:
        os.makedirs(prefix_dir, exist_ok=True)
    if rank == 0:
        app_args._args().save_to_file("%sparams.out" % args.output_prefix, args)

    kb = args.kb if args.kb else 0.0083144621                   # :53-61 (GROMACS units unless overridden)
    mass_factor = args.mass_factor if args.mass_factor else 1.6605402
    lj_cutoff, cg_cutoff = args.lj_cutoff, args.cg_cutoff
    max_cutoff = max(lj_cutoff, cg_cutoff)                      # :77-80
    dt = args.dt
    time0 = time.time()

    if getattr(args, "debug", None):                            # :65-72 -- `--debug logger[:regex],...`; the engine's own stage timer
        import logging                                             # is the environment variable CLB_TRACE=1 (csrc/engine.cu)
        logging.basicConfig()
        for item in args.debug.split(","):
            name_filter = item.split(":")
            print("Activating logger %s" % name_filter[0])
            logging.getLogger(name_filter[0].strip()).setLevel(logging.DEBUG)
            if len(name_filter) == 2:
                logging.getLogger(name_filter[0].strip()).addFilter(app_args.RegexpFilter(name_filter[1]))
    if getattr(args, "check_topology", False):                  # :74-75
        import logging
        logging.getLogger("TopologyManager").setLevel(logging.WARN)
    has_excl_file = args.exclusion_list is not None and os.path.exists(args.exclusion_list)
    gt = gromacs_topology.GromacsTopology(args.top, generate_exclusions=not has_excl_file).read()
    conf = files_io.GROFile(args.conf)
    conf.read()
    box = conf.box

    integrator_step = args.int_step
    if args.trj_collect > 0:
        integrator_step = min(args.int_step, args.trj_collect)  # :100-103
    skin = 0.16 if args.skin == "auto" else float(args.skin)
    rng_seed = args.rng_seed
    if not rng_seed or rng_seed == -1:
        rng_seed = random.randint(10, 1000000)
        if world > 1:                          # every rank must use the same seed
            import torch.distributed as dist
            box_ = [rng_seed]
            dist.broadcast_object_list(box_, src=0)
            rng_seed = box_[0]
        args.rng_seed = rng_seed
    prefix = "%s_%s" % (args.output_prefix, rng_seed)
    print("Skin: %s\nRNG Seed: %s\nBoltzmann constant: %s" % (skin, rng_seed, kb))

    part_prop, particle_list = gromacs_topology.gen_particle_list(conf, gt)
    npart = len(particle_list)
    print("Reads %d particles with properties %s" % (npart, part_prop))
    if args.temperature is None:
        raise RuntimeError("Temperature not defined!")
    temperature = args.temperature * kb                          # :135

    # ---- System / storage / integrator (:148-171)
    system = espressopp.System()
    system.rng = espressopp.esutil.RNG(rng_seed)
    system.skin = skin
    node_grid = (tuple(int(x) for x in args.node_grid.split(",")) if getattr(args, "node_grid", None)
                 else espressopp.tools.decomp.nodeGrid(1))
    cell_grid = espressopp.tools.decomp.cellGrid(box, node_grid, max_cutoff, skin)
    print("Cell grid: %s, node grid: %s" % (cell_grid, node_grid))
    system.bc = espressopp.bc.OrthorhombicBC(system.rng, box)
    system.storage = espressopp.storage.DomainDecomposition(system, node_grid, cell_grid)
    integrator = espressopp.integrator.VelocityVerlet(system)
    integrator.dt = dt
    system.integrator = integrator
    part_prop = list(part_prop)
    mi = part_prop.index("mass")
    # masses go to the engine UNSCALED, as in the reference (:168); mass_factor only enters the density print (:128) and the
    # Maxwell-Boltzmann draw (:136-146).  The reaction post-processes later write the same unscaled topology masses.
    density = sum(p[mi] for p in particle_list) * mass_factor / (box[0] * box[1] * box[2])
    print("Density: %s kg/m^3\nBox: %s nm" % (density, list(box)))
    if getattr(args, "gen_velocity", False):                    # :136-146
        vx, vy, vz = espressopp.tools.velocities.gaussian(temperature, npart, [p[mi] * mass_factor for p in particle_list], kb=1.0, seed=rng_seed)
        vel = [espressopp.Real3D(a, b, c) for a, b, c in zip(vx, vy, vz)]
        print("Generating velocities from Maxwell-Boltzmann distribution T=%s (%s)" % (args.temperature, temperature))
        particle_list = [tuple(p) + (v,) for p, v in zip(particle_list, vel)]
        part_prop.append("v")
    elif any(a.velocity is not None for a in conf.atoms.values()):
        # gen_particle_list (gromacs_topology.py:1426) carries no 'v' property: the reference starts from zero velocities
        print("Note: velocities in %s are not used (no gen_velocity): the run starts from zero velocities, as in the reference" % args.conf)
    system.storage.addParticles(particle_list, *part_prop)
    system.storage.decompose()

    # ---- exclusions + Verlet list (:174-197)
    if has_excl_file:
        exclusions = [tuple(int(x) for x in l.split()) for l in open(args.exclusion_list) if l.strip()]
        print("Read exclusion list from %s (total: %d)" % (args.exclusion_list, len(exclusions)))
        if len(exclusions) == 0 and gt.bonds and not getattr(args, "do_not_exclude_bonds", False):
            raise RuntimeError("Exclusion list in %s is empty" % args.exclusion_list)
        gt.exclusions = exclusions
    exclusions = sorted(tuple(x) for x in gt.exclusions)
    if exclusions and rank == 0:                                  # :182-187: rewritten whenever the list is non-empty
        with open("exclusion_%s.list" % os.path.basename(args.top).split(".")[0], "w") as f:
            f.write("\n".join("%d %d" % tuple(p) for p in exclusions))
    print("Excluded pairs from LJ interaction: %d" % len(exclusions))
    dynamic_exclude = espressopp.DynamicExcludeList(integrator, exclusions)
    verletlist = espressopp.VerletList(system, cutoff=max_cutoff, exclusionlist=dynamic_exclude)

    # ---- topology manager, hooks, reactions (:211-272)
    topology_manager = espressopp.integrator.TopologyManager(system)
    system.topology_manager = topology_manager
    hooks = _load_hooks()
    ar = None
    chem_fpls, reactions, ext_to_integrator = [], [], []
    sc = None
    ar_interval = 0
    if args.reactions:
        if not os.path.exists(args.reactions):
            raise RuntimeError("Reaction config %s not found" % args.reactions)
        cfg = reaction_parser.parse_config(args.reactions)
        sc = reaction_setup.SetupReactions(system, verletlist, gt, topology_manager, cfg, args)
        ar, chem_fpls, reactions, ext_to_integrator = sc.setup_reactions()
        ar_interval = sc.ar_interval
        if rank == 0:                                            # :261-263
            import shutil
            shutil.copyfile(args.reactions, "%s_%s" % (prefix, os.path.basename(args.reactions)))
        integrator_step = min(integrator_step, ar_interval)      # :265-267
        print("Set up %d reactions in %d groups, interval %d" % (len(reactions), len(chem_fpls), ar_interval))
        if "hook_postsetup_reaction" in hooks:                   # :272
            hooks["hook_postsetup_reaction"](system, integrator, gt, args, ar)
    sim_step = args.run // integrator_step                       # :103,:268 (Python-2 integer division)
    dynamic_types = sc.dynamic_types if sc else set()
    print("Dynamic type ids: %s" % sorted(dynamic_types))       # :276

    # ---- force field (:298-310)
    cr_observs = dict(sc.cr_observs) if sc else {}
    maximum_conversion = []                                      # :280-287: its observables join cr_observs, hence the energy CSV
    if args.maximum_conversion:
        maximum_conversion = tools.get_maximum_conversion(args, system, chem_fpls, gt, cr_observs)
    cr_observs, _ = gromacs_topology.set_nonbonded_interactions(system, gt, verletlist, lj_cutoff, getattr(args, "coulomb_cutoff", None),
                                                                cg_cutoff, tables=getattr(args, "table_groups", None), cr_observs=cr_observs)
    dyn_fpl, static_fpl, _ = gromacs_topology.set_bonded_interactions(system, gt, dynamic_types)
    dyn_ftl, static_ftl = gromacs_topology.set_angle_interactions(system, gt, dynamic_types)
    dyn_fql, static_fql = gromacs_topology.set_dihedral_interactions(system, gt, dynamic_types)
    gromacs_topology.set_pair_interactions(system, gt, args, dynamic_types)
    print("Interactions: %s" % ", ".join(system.getNameOfInteraction(k) for k in range(system.getNumberOfInteractions())))

    # ---- cap force (:320-324)
    if getattr(args, "max_force", -1) is not None and float(getattr(args, "max_force", -1)) > -1:
        integrator.addExtension(espressopp.integrator.CapForce(system, float(args.max_force)))
        print("Cap force to %s" % args.max_force)

    # ---- thermostat (:326-376)
    if args.thermostat != "lv":
        raise NotImplementedError("thermostat %r: only the Langevin thermostat (lv) is inside the engine's scope (SURVEY E20)" % args.thermostat)
    thermostat = espressopp.integrator.LangevinThermostat(system)
    thermostat.temperature = temperature
    thermostat.gamma = args.thermostat_gamma
    integrator.addExtension(thermostat)
    if getattr(args, "barostat", None) and getattr(args, "pressure", None):
        raise NotImplementedError("barostats are outside the engine's scope (SURVEY E20)")

    # ---- dynamic exclusions and topology bookkeeping (:378-444)
    for f in chem_fpls:
        if not getattr(args, "do_not_exclude_bonds", False):     # :425-430
            dynamic_exclude.observe_tuple(f.fpl)
        topology_manager.observe_tuple(f.fpl)
    for lst in list(dyn_fpl.values()) + list(static_fpl):
        topology_manager.observe_tuple(lst)
    for lst in dyn_ftl.values():
        dynamic_exclude.observe_triple(lst)
    for lst in dyn_fql.values():
        dynamic_exclude.observe_quadruple(lst)
    # registered type tuples: every dynamic angle/dihedral parameter set becomes a template for generated tuples
    for arity, dyn, params in ((3, dyn_ftl, gt.angleparams), (4, dyn_fql, gt.dihedralparams)):
        for key, lst in dyn.items():
            for pt, p in params.items():
                if int(p["func"]) != key.func or not (set(pt) & set(dynamic_types)):
                    continue
                if arity == 3:
                    topology_manager.register_triplet(lst, *pt)
                else:
                    topology_manager.register_quadruplet(lst, *pt)
    topology_manager.initialize_topology()
    integrator.addExtension(topology_manager)

    # ---- observables (:446-569): same columns, in the same order, as the reference's energy CSV
    energy_file = "%s_energy_%s.csv" % (args.output_prefix, rng_seed)
    monitor = espressopp.analysis.SystemMonitor(system, integrator, espressopp.analysis.SystemMonitorOutputCSV(energy_file))
    temp_obs = espressopp.analysis.Temperature(system)
    monitor_filter = args.system_monitor_filter.split(",") if getattr(args, "system_monitor_filter", None) else None   # :458-460
    monitor.add_observable("T", temp_obs)
    monitor.add_observable("Ekin", espressopp.analysis.KineticEnergy(system, temp_obs))
    if getattr(args, "store_pressure", False):
        print("Note: store_pressure is ignored (analysis.Pressure is outside the engine's scope, SURVEY E20)")
    for label, interaction in sorted(system.getAllInteractions().items()):       # :471-480: sorted by label; the filter only hides columns from info()
        visible = True if not monitor_filter else any(v in label for v in monitor_filter)
        monitor.add_observable(label, espressopp.analysis.PotentialEnergy(system, interaction), visible)
    for (cr_type, _, ts), obs in cr_observs.items():                             # :481-489
        name = "cr_" + "_".join(str(x) for x in (cr_type if isinstance(cr_type, (tuple, list)) else [cr_type]))
        monitor.add_observable(name if ts is None else "%s_%s" % (name, ts), obs)
    for i, f in enumerate(chem_fpls):
        monitor.add_observable("count_%d" % i, espressopp.analysis.NFixedPairListEntries(system, f.fpl))
    if getattr(args, "count_tuples", False):                                     # :503-538
        A = espressopp.analysis
        for prefix_, statics, dynamics, cls in (("bcount", static_fpl, dyn_fpl, A.NFixedPairListEntries), ("acount", static_ftl, dyn_ftl, A.NFixedTripleListEntries),
                                                ("qcount", static_fql, dyn_fql, A.NFixedQuadrupleListEntries)):
            for n_, lst in enumerate(list(statics) + list(dynamics.values())):
                monitor.add_observable("%s_%d" % (prefix_, n_), cls(system, lst))
        monitor.add_observable("vl_excl", A.NExcludeListEntries(system, verletlist))
    if getattr(args, "count_types", None):                                       # :544-549
        for sym in args.count_types.split(","):
            tid = gt.atomsym_atomtype[sym]
            print("Observer %-9s (%s)" % (sym, tid))
            monitor.add_observable("num_type_%s_%s" % (sym, tid), espressopp.analysis.ChemicalConversion(system, tid))
    if getattr(args, "count_types_state", None) is not None:                     # :551-559
        for ts in args.count_types_state.split(","):
            type_name, state = ts.split(":")
            monitor.add_observable("st_%s_%d" % (type_name, int(state)),
                                   espressopp.analysis.ChemicalConversionTypeState(system, gt.atomsym_atomtype[type_name], int(state)))
    # :278, :565-569: cr_interval = min(integrator step, reaction interval) = the integrator step (itself capped by the reaction
    # interval, :265-267); data are collected every min(cr_interval, energy_collect) steps
    cr_interval = min(integrator_step, ar_interval) if ar is not None else integrator_step
    if args.energy_collect > 0:
        energy_collect = min(cr_interval, args.energy_collect)
        integrator.addExtension(espressopp.integrator.ExtAnalyze(monitor, energy_collect))
        print("Configured system analysis, collect data every %d steps" % energy_collect)

    if getattr(args, "gro_trj_collect", None):                   # :684-696
        dump_trj = espressopp.io.DumpGRO(system, integrator, filename=("%s_traj.gro" % prefix) if rank == 0 else os.devnull, unfolded=True, append=True)
        integrator.addExtension(espressopp.integrator.ExtAnalyze(dump_trj, int(args.gro_trj_collect)))
        print("Set gro trajectory saver, save every %d steps" % int(args.gro_trj_collect))

    # ---- start/stop of the reactions in units of outer iterations (:700-721)
    # :636-640, :661-665 -- rounded UP to whole outer iterations
    k_enable_reactions = int(math.ceil(args.start_ar / float(integrator_step))) if (ar is not None and args.start_ar >= 0) else -1
    stop_ar = getattr(args, "stop_ar", -1)
    k_stop_reactions = int(math.ceil(stop_ar / float(integrator_step))) if (ar is not None and stop_ar is not None and stop_ar >= 0) else -1
    print("Running %d steps as %d x integrator.run(%d); reactions start at outer step %d" % (args.run, sim_step, integrator_step, k_enable_reactions))
    espressopp.analysis.CMVelocity(system).reset()

    rate_file = open("%s_new_rates.csv" % prefix, "w") if (getattr(args, "rate_arrhenius", False) and rank == 0) else None     # :712-714

    # ---- main loop (:705-800)
    monitor.dump()                                               # :705 -- the row of step 0
    eq_run = int(args.eq_steps / sim_step) if (maximum_conversion and getattr(args, "eq_steps", 0) > 0 and sim_step) else 0      # :286-287
    total_time0 = time.time()
    integrator_loop = 0.0
    if "hook_before_sim" in hooks:                               # :726
        hooks["hook_before_sim"](system, integrator, ar, gt)
    reactions_enabled = False
    stop_simulation = False
    energy0, bonds0 = 0.0, 0
    for k in range(sim_step):
        monitor.info()
        if k == k_enable_reactions and ar is not None:           # :735-755
            print("Enabling chemical reactions at step %d" % integrator.step)
            integrator.addExtension(ar)
            for ext in ext_to_integrator:
                integrator.addExtension(ext)
            reactions_enabled = True
            # :742-746 -- the INPUT configuration object is updated (unfolded positions + velocities) and written, unconditionally
            print("Save configuration before start of the reaction, filename: %s_before_reaction_confout.gro" % prefix)
            conf.update_position(system, unfolded=True)
            if rank == 0:
                conf.write("%s_before_reaction_confout.gro" % prefix, with_velocity=True)
            if "hook_init_reaction" in hooks:                    # :748-750
                print("Processing hook_init_reaction")
                if not hooks["hook_init_reaction"](system, integrator, ar, gt, args):
                    raise RuntimeError("hook_init_reaction return False")
        if reactions_enabled:                                    # :757-776
            if not stop_simulation:
                for obs, stop_value in maximum_conversion:
                    val = obs.compute()
                    if val >= stop_value:
                        print("Reaches %s of the conversion => Stop simulation" % val)
                        stop_simulation = True
            if stop_simulation:
                if eq_run == 0:
                    break
                eq_run -= 1
            if getattr(args, "rate_arrhenius", False):
                bonds0 = sum(f.fpl.totalSize() for f in chem_fpls)
                energy0 = monitor.potential_energy
            if k == k_stop_reactions or stop_simulation:
                ar.disconnect()
        t0 = time.time()
        integrator.run(integrator_step)                           # :780 -- 100 % of the compute
        integrator_loop += time.time() - t0
        if "hook_at_step" in hooks:
            hooks["hook_at_step"](system, integrator, ar, gt, args, k * integrator_step)     # :783
        if getattr(args, "rate_arrhenius", False) and reactions_enabled:      # :785-796: k = exp(-dE/kT) per new bond
            delta_bonds = sum(f.fpl.totalSize() for f in chem_fpls) - bonds0
            if delta_bonds > 0:
                energy_delta = (monitor.potential_energy - energy0) / float(delta_bonds)
                new_rate = math.exp(-energy_delta / temperature)
                print("%d\tChange reaction rate, delta_E=%s, new_k=%s, delta_bonds=%d" % (k * integrator_step, energy_delta, new_rate, delta_bonds))
                if rate_file is not None:
                    rate_file.write("%d %e\n" % (k * integrator_step, new_rate))
                for r_ in reactions:
                    r_.rate = new_rate
    total_time = time.time() - total_time0
    if rate_file is not None:
        rate_file.close()
    if "hook_end" in hooks:                                      # :800
        hooks["hook_end"](system, integrator, ar, gt, args)
    monitor.info()                                               # :802

    # ---- outputs (:800-1081)
    e = system._ctx.require_engine()
    g = e.get_particles(fields=("pos", "image", "type", "state", "res_id", "mass", "q"))
    ids = sorted(system._ctx.pid)
    id2type = {v: k for k, v in gt.atomsym_atomtype.items()}
    chem_bonds = [np.asarray(f.fpl.getAllBonds(), np.int64).reshape(-1, 2) for f in chem_fpls]
    if rank == 0:
        # :1008-1012 -- the input configuration object with the folded end positions: title, atom and residue names as read (the
        # current types are in _state.dat and _output_topol.top).  GROFile.update_position(unfolded=False) of the reference leaves
        # the velocities alone (files_io.py:276-279): they are those of the input file, or those stored when the
        # before-reaction file was written
        for k, pid in enumerate(ids):
            conf.atoms[pid] = conf.atoms[pid]._replace(position=tuple(g["pos"][k]))
        conf.write("%s_confout.gro" % prefix, with_velocity=True)
        print("Wrote end configuration to: %s_confout.gro" % prefix)
        np.savetxt("%s_state.dat" % prefix, np.column_stack([ids, g["type"], g["state"], g["res_id"]]), fmt="%d", header="id type state res_id")
        for i, bonds_i in enumerate(chem_bonds):
            np.savetxt("%s_bonds_chem_%d.dat" % (prefix, i), bonds_i, fmt="%d")
        # :897-988 -- one row per tuple: ids, func, parameters and where it comes from (static list, type-dispatched "dynamic" list,
        # reaction list); a tuple whose current types have no parameters is flagged like in the reference
        type_of = {pid: int(g["type"][k]) for k, pid in enumerate(ids)}
        name_of = lambda pid: id2type.get(type_of[pid], "?")

        def lookup(params, tup):
            key = tuple(type_of[x] for x in tup)
            return params.get(key) or params.get(key[::-1])
        tuple_lines = {}
        for label, getter, statics, dynamics, params in (("bonds", "getAllBonds", static_fpl, dyn_fpl, gt.bondparams), ("angles", "getAllTriples", static_ftl, dyn_ftl, gt.angleparams),
                                                         ("dihedrals", "getAllQuadruples", static_fql, dyn_fql, gt.dihedralparams)):
            lines = []
            for lst in statics:
                func, pr = lst.params
                lines += [[*t, func, *pr, "; static"] for t in getattr(lst, getter)()]
            for lst in dynamics.values():
                for t in getattr(lst, getter)():
                    q = lookup(params, t)
                    lines.append([*t, q["func"], *q["params"], "; dynamic"] if q else [*t, "; MISSING params type: %s dynamic" % "-".join(name_of(x) for x in t)])
            if label == "bonds":
                for bonds_i in chem_bonds:
                    for t in bonds_i.tolist():
                        q = lookup(params, t)
                        names = "-".join(name_of(x) for x in t)
                        lines.append([*t, ("%s %s ; chem %s" % (q["func"], " ".join(str(x) for x in q["params"]), names)) if q else ("; chem MISSING params type: %s" % names)])
            tuple_lines[label] = lines
            with open("%s_%s.dat" % (prefix, label), "w") as f:
                f.writelines(" ".join(str(x) for x in row) + "\n" for row in lines)
        # final topology (:834-994): the force-field sections of the input, every particle with its current type, mass, charge and
        # residue, and the tuple rows written above (static and type-dispatched lists, reaction bonds) with their parameters
        top_atoms = {pid: dict(gt.atoms[pid], type_id=int(g["type"][k]), mass=float(g["mass"][k]), charge=float(g["q"][k]), chain_idx=int(g["res_id"][k]))
                     for k, pid in enumerate(ids) if pid in gt.atoms}
        gt.gt.write_system("%s_output_topol.top" % prefix, top_atoms, tuple_lines["bonds"], tuple_lines["angles"], tuple_lines["dihedrals"],
                           {v: k_ for k_, v in gt.atomsym_atomtype.items()})
        print("Write output topology: %s_output_topol.top" % prefix)
        if ar is not None:                                           # :1027-1036
            ar.save_reaction_counters("%s_reaction_counters" % prefix)
            with open("%s_reaction_counters" % prefix, "a") as f:
                f.write("\n\nReaction index\n")
                for ridx in sorted(sc.reaction_index):
                    f.write("%s %s\n" % (ridx, sc.reaction_index[ridx]))
            ar.save_intra_inter_counter("%s_intra_inter_counters" % prefix)
        with open("%s_benchmark.csv" % prefix, "a") as f:          # record format of the reference (:997-998): nranks NPart total loop
            f.write("%d %d %.6f %.6f\n" % (world, npart, total_time, integrator_loop))
    # :1014-1017 -- the whole configuration through io.DumpGRO (collective read, rank 0 writes)
    espressopp.io.DumpGRO(system, integrator, filename=("%s_whole_confout.gro" % prefix) if rank == 0 else os.devnull, unfolded=False, append=False).dump()
    # :1004-1006 -- bond graph, residue graph and residue membership at the end of the run (collective read, rank 0 writes)
    tm_files = [("save_topology", "%s_topology.dat"), ("save_res_topology", "%s_res_topology.dat"), ("save_residues", "%s_residue_list.dat")]
    for meth, pattern in tm_files:
        getattr(topology_manager, meth)(pattern % prefix if rank == 0 else os.devnull)
    if rank == 0:
        import pickle                                               # :1040-1076 -- per-bucket timers of the run, averaged over the ranks
        ext_timers = {}
        for k_ in range(integrator.getNumberOfExtensions()):
            ext = integrator.getExtension(k_)
            tmr = tools.average_timers(ext.get_timers()) if hasattr(ext, "get_timers") else {}
            if tmr:
                ext_timers["%s_%d" % (type(ext).__name__, id(ext))] = tmr
        with open("%s_benchmark.pck" % prefix, "wb") as f:
            pickle.dump({"traj_timers": {}, "topol_timers": tools.average_timers(topology_manager.get_timers()),
                         "integrator_timers": tools.get_integrator_timers(integrator.getTimers(), system),
                         "extension_timers": ext_timers, "verlet_list": tools.average_timers(verletlist.get_timers())}, f)
    timers = tools.get_integrator_timers(integrator.getTimers(), system)
    e_t, e_c = e.timers()
    timers.update({"n_" + k: v for k, v in e_c.items()})
    print("final: steps=%d total=%.3fs integratorLoop=%.3fs (%.1f steps/s) setup=%.3fs" %
          (integrator.step, total_time, integrator_loop, integrator.step / max(integrator_loop, 1e-9), total_time0 - time0))
    print("engine timers/counters: %s" % {k: (round(v, 4) if isinstance(v, float) else v) for k, v in timers.items()})
    return dict(system=system, integrator=integrator, ar=ar, topology=gt, chem_fpls=chem_fpls, reactions=reactions, prefix=prefix,
                steps=integrator.step, integrator_loop=integrator_loop, monitor=monitor)


if __name__ == "__main__":
    main(sys.argv[1:])

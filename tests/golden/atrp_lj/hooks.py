# hooks.py for the atrp_lj plumbing config (Python 3, seeded).  Same job as the hook shipped with the reference example
# (examples/atrp_lj/hooks.py:28-80): at the moment the reactions are switched on, activate 20 random trimers --
# first MA end -> FA (state 2, the growing radical end), other MA end -> PA, middle ML -> PL (state 2).
import collections
import random


def hook_init_reaction(system, integrator, ar, topol, args):
    rnd = random.Random(int(args.rng_seed) + 1)
    name2type = topol.atomsym_atomtype
    res2pids = collections.defaultdict(list)
    for pid in system.storage.getAllParticleIDs():
        res2pids[system.storage.getParticle(pid).res_id].append(pid)
    chosen = rnd.sample(sorted(res2pids), 20)
    for res_id in chosen:
        first = True
        for pid in sorted(res2pids[res_id]):
            p = system.storage.getParticle(pid)
            if p.type == name2type["MA"]:
                new = "FA" if first else "PA"
                system.storage.modifyParticle(pid, "type", name2type[new])
                if first:
                    system.storage.modifyParticle(pid, "state", 2)
                system.storage.modifyParticle(pid, "mass", topol.gt.atomtypes[new]["mass"])
                first = False
            elif p.type == name2type["ML"]:
                system.storage.modifyParticle(pid, "type", name2type["PL"])
                system.storage.modifyParticle(pid, "mass", topol.gt.atomtypes["PL"]["mass"])
                system.storage.modifyParticle(pid, "state", 2)
    print("Activated %d trimers" % len(chosen))
    return True      # the driver stops when the hook does not return a true value (src/start_simulation.py:749-750)

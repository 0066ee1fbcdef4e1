# hooks.py for examples/pccg_lj/chemical_reactions (Python 3, seeded).  Same job as the shipped hook file
# (examples/pccg_lj/chemical_reactions/hooks.py:44-108): an angle histogram accumulated at every outer step and written at the
# end, and -- when the reactions are switched on -- 20 random divinyl monomers activated: first MA bead -> FA (state 2), the
# other -> PA.
import random

import numpy

import espressopp


class AngleDistribution:
    def __init__(self):
        self.histogram = numpy.zeros(100)
        self.steps = 0


angle_dist = AngleDistribution()


def hook_before_sim(system, integrator, ar, gt):
    analysis_angles = espressopp.analysis.AngleDistribution(system)
    analysis_angles.load_from_topology_manager(system.topology_manager)
    angle_dist.analysis_angles = analysis_angles


def hook_at_step(system, integrator, ar, gt, args, step):
    angle_dist.histogram += numpy.array(angle_dist.analysis_angles.compute(100))
    angle_dist.steps += 1


def hook_end(system, integrator, ar, gt, args):
    theta = numpy.arange(0.0, numpy.pi, numpy.pi / 100)
    numpy.savetxt("output_angle.csv", numpy.column_stack((theta, angle_dist.histogram)))


def hook_init_reaction(system, integrator, ar, topol, args):
    rnd = random.Random(int(args.rng_seed) + 1)
    name2type = topol.atomsym_atomtype
    number_to_activate = 20
    res_id2pids = {k: (i, i + 1) for k, i in enumerate(range(1, 4001, 2), 1)}
    res_ids = rnd.sample(range(1, 2001), number_to_activate)
    for res_id in res_ids:
        activated_monomer = False
        for pid in res_id2pids[res_id]:
            p = system.storage.getParticle(pid)
            if p.type == name2type["MA"]:
                new = "PA" if activated_monomer else "FA"
                system.storage.modifyParticle(pid, "type", name2type[new])
                if not activated_monomer:
                    system.storage.modifyParticle(pid, "state", 2)
                system.storage.modifyParticle(pid, "mass", topol.gt.atomtypes[new]["mass"])
                activated_monomer = True
    system.storage.decompose()
    types = [system.storage.getParticle(pid).type for r in res_ids for pid in res_id2pids[r]]
    assert types.count(name2type["FA"]) == number_to_activate and types.count(name2type["PA"]) == number_to_activate
    print("Activated %d monomers" % number_to_activate)
    return True

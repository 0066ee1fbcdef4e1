"""espressopp.interaction.* as used by chemlab (src/chemlab/gromacs_topology.py:463-1411, reaction_setup.py:444-467).

Potentials are plain records; interaction objects own an engine interaction handle once attached and forward
setPotential to the C-ABI.  Potential kinds outside north_star are constructible stubs."""
import numpy as np

from ._context import not_in_scope


def read_pot(filename):
    """Reads an ESPResSo++ table ('r e f' rows written by tools/convert_gromacs2espp.py:84,107)."""
    data = np.loadtxt(filename, comments="#", ndmin=2)
    if data.shape[1] < 3 or len(data) < 2:
        raise RuntimeError("table %s: expected rows 'r energy force'" % filename)
    return data[:, 0].copy(), data[:, 1].copy(), data[:, 2].copy()


# ------------------------------------------------------------------ potentials (records)
class _Pot:
    kind = None

    def params(self):
        return ()


class Tabulated(_Pot):
    """Tabulated(itype, filename, cutoff): gromacs_topology.py:696-707,919-925."""
    kind = "Tabulated"

    def __init__(self, itype=1, filename=None, cutoff=None, **kw):
        self.itype, self.filename, self.cutoff = int(itype), filename, cutoff


class TabulatedAngular(Tabulated):
    kind = "TabulatedAngular"


class TabulatedDihedral(Tabulated):
    kind = "TabulatedDihedral"


class LennardJones(_Pot):
    """LennardJones(epsilon, sigma, cutoff, shift='auto'): gromacs_topology.py:715-721."""
    kind = "LennardJones"

    def __init__(self, epsilon=1.0, sigma=1.0, cutoff=2.5, shift="auto", **kw):
        self.epsilon, self.sigma, self.cutoff, self.shift = float(epsilon), float(sigma), float(cutoff), shift

    def params(self):
        """as a pair-list potential (1-4 pairs): {epsilon, sigma, cutoff, shift}; shift 'auto' = U_LJ(cutoff)."""
        sr6 = (self.sigma / self.cutoff) ** 6
        shift = 4.0 * self.epsilon * (sr6 * sr6 - sr6) if self.shift == "auto" else float(self.shift or 0.0)
        return (self.epsilon, self.sigma, self.cutoff, shift)


class MixedTabulated(_Pot):
    """MixedTabulated(itype, tab1, tab2, cr_obs | mix_value, cutoff): gromacs_topology.py:757-790."""
    kind = "MixedTabulated"

    def __init__(self, itype=1, tab1=None, tab2=None, cr_observation=None, mix_value=None, cutoff=None, **kw):
        self.itype, self.tab1, self.tab2, self.cutoff = int(itype), tab1, tab2, cutoff
        self.cr_observation, self.mix_value = cr_observation, mix_value


class Harmonic(_Pot):
    """Harmonic(K, r0): U = K (r-r0)^2 (chemlab halves the GROMACS K: gromacs_topology.py:918)."""
    kind = "Harmonic"

    def __init__(self, K=1.0, r0=0.0, cutoff=None, shift=0.0, **kw):
        self.K, self.r0 = float(K), float(r0)

    def params(self):
        return (self.K, self.r0)


class FENE(_Pot):
    kind = "FENE"

    def __init__(self, K=1.0, r0=0.0, rMax=1.0, **kw):
        self.K, self.r0, self.rMax = float(K), float(r0), float(rMax)

    def params(self):
        return (self.K, self.r0, self.rMax)


class FENELennardJones(_Pot):
    """FENELennardJones(K, r0, rMax, sigma, epsilon): [ bondtypes ] func 9 (gromacs_topology.py:935-961; doc/topology.rst:72-79)."""
    kind = "FENELennardJones"

    def __init__(self, K=1.0, r0=0.0, rMax=1.0, sigma=1.0, epsilon=1.0, **kw):
        self.K, self.r0, self.rMax, self.sigma, self.epsilon = float(K), float(r0), float(rMax), float(sigma), float(epsilon)

    def params(self):
        return (self.K, self.r0, self.rMax, self.sigma, self.epsilon)


class AngularHarmonic(_Pot):
    """AngularHarmonic(K, theta0): gromacs_topology.py:1073."""
    kind = "AngularHarmonic"

    def __init__(self, K=1.0, theta0=0.0, **kw):
        self.K, self.theta0 = float(K), float(theta0)

    def params(self):
        return (self.K, self.theta0)


class Cosine(_Pot):
    """Cosine(K, theta0): U = K (1 + cos(theta - theta0)) [EXT, U29].  chemlab passes the topology K unchanged (angletypes func 11,
    gromacs_topology.py:1082); doc/topology.rst:99-104,160 writes 1/2 K with the footnote "internally divided by 2.0" -- the wording
    it uses for the harmonic angle, whose K chemlab itself halves (:1073) -- so the documentation describes a halving the code
    does not do.  The code path is followed."""
    kind = "Cosine"

    def __init__(self, K=1.0, theta0=0.0, **kw):
        self.K, self.theta0 = float(K), float(theta0)

    def params(self):
        return (self.K, self.theta0)


class DihedralHarmonic(_Pot):
    """DihedralHarmonic(K, phi0): U = 1/2 K (phi - phi0)^2 [EXT, U30] -- the formula of doc/topology.rst:123-129 for dihedraltypes
    func 12, whose K chemlab passes unchanged (gromacs_topology.py:1199-1202; no "divided by 2" footnote in the table, :177).  The
    engine's kind 8 evaluates K' (phi - phi0)^2 (include/chemlab_b200.h): K' = K / 2 is handed down."""
    kind = "DihedralHarmonic"

    def __init__(self, K=1.0, phi0=0.0, **kw):
        self.K, self.phi0 = float(K), float(phi0)

    def params(self):
        return (0.5 * self.K, self.phi0)


# ------------------------------------------------------------------ non-bonded interactions over the Verlet list
class _VerletListInteraction:
    _nb_kind = None

    def __init__(self, vl):
        self._vl = vl
        self._ctx = vl._system._ctx
        self._h = None
        self._pots = {}

    def setPotential(self, type1, type2, potential):
        self._pots[(int(type1), int(type2))] = potential
        if self._h is not None:
            self._push(int(type1), int(type2), potential)

    def getPotential(self, type1, type2):
        return self._pots.get((type1, type2), self._pots.get((type2, type1)))

    def _attach(self, e):
        if self._h is None:
            self._h = e.add_nonbonded(self._nb_kind)
            for (t1, t2), p in self._pots.items():
                self._push(t1, t2, p)

    def computeEnergy(self):
        return self._ctx.require_engine().energy(self._h)


class VerletListTabulated(_VerletListInteraction):
    """gromacs_topology.py:512."""
    _nb_kind = "Tabulated"

    def _push(self, t1, t2, p):
        self._ctx.engine.nb_set_tabulated(self._h, t1, t2, self._ctx.table(p), p.cutoff)


class VerletListLennardJones(_VerletListInteraction):
    """gromacs_topology.py:511."""
    _nb_kind = "LennardJones"

    def _push(self, t1, t2, p):
        self._ctx.engine.nb_set_lj(self._h, t1, t2, p.epsilon, p.sigma, p.cutoff, 1 if p.shift == "auto" else 0)


class VerletListMixedTabulated(_VerletListInteraction):
    """gromacs_topology.py:757-790 (func 10 with a conversion observable, func 12 with a constant mix)."""
    _nb_kind = "MixedTabulated"

    def _push(self, t1, t2, p):
        class _T:  # table records for the two files
            def __init__(s, f, it): s.filename, s.itype = f, it
        a, b = self._ctx.table(_T(p.tab1, p.itype)), self._ctx.table(_T(p.tab2, p.itype))
        obs = p.cr_observation
        if obs is not None:
            self._ctx.engine.nb_set_mixed(self._h, t1, t2, a, b, obs.compute(), obs.type_id, obs.total, p.cutoff)
        else:
            self._ctx.engine.nb_set_mixed(self._h, t1, t2, a, b, float(p.mix_value), -1, 1.0, p.cutoff)


# ------------------------------------------------------------------ bonded interactions over fixed tuple lists
class _FixedListInteraction:
    _typed = 0

    def __init__(self, system, flist, potential=None):
        self._ctx = system._ctx
        self._list = flist
        self._h = None
        self._pot = potential
        self._typed_pots = {}

    def setPotential(self, potential=None, **types):
        if types:   # type1=, type2=[, type3=, type4=]
            key = tuple(int(types[k]) for k in ("type1", "type2", "type3", "type4") if k in types)
            self._typed_pots[key] = potential
            if self._h is not None:
                self._push(key, potential)
        else:
            self._pot = potential
            if self._h is not None:
                self._push((), potential)

    def getFixedPairList(self):
        return self._list

    getFixedTripleList = getFixedQuadrupleList = getFixedPairList

    def _push(self, types, p):
        table = self._ctx.table(p) if isinstance(p, Tabulated) else -1
        self._ctx.engine.bonded_set_potential(self._h, types, p.kind, p.params(), table)

    def _attach(self, e):
        if self._h is None:
            self._list._attach(e)
            self._h = e.add_bonded(self._list._h, self._typed)
            if self._pot is not None:
                self._push((), self._pot)
            for key, p in self._typed_pots.items():
                self._push(key, p)

    def computeEnergy(self):
        return self._ctx.require_engine().energy(self._h)


class _TypedFixedListInteraction(_FixedListInteraction):
    _typed = 1


FixedPairListHarmonic = FixedPairListTabulated = FixedPairListFENE = FixedPairListFENELennardJones = FixedPairListLennardJones = _FixedListInteraction
FixedPairListTypesHarmonic = FixedPairListTypesTabulated = FixedPairListTypesFENE = FixedPairListTypesFENELennardJones = FixedPairListTypesLennardJones = _TypedFixedListInteraction
FixedTripleListAngularHarmonic = FixedTripleListTabulatedAngular = FixedTripleListCosine = _FixedListInteraction
FixedTripleListTypesAngularHarmonic = FixedTripleListTypesTabulatedAngular = FixedTripleListTypesCosine = _TypedFixedListInteraction
FixedQuadrupleListTabulatedDihedral = FixedQuadrupleListDihedralHarmonic = _FixedListInteraction
FixedQuadrupleListTypesTabulatedDihedral = FixedQuadrupleListTypesDihedralHarmonic = _TypedFixedListInteraction

class CoulombTruncated(_Pot):
    """CoulombTruncated(prefactor, cutoff): gromacs_topology.py:866-878 (prefactor 138.935485 * fudgeQQ)."""
    kind = "CoulombTruncated"

    def __init__(self, prefactor=1.0, cutoff=None, **kw):
        self.prefactor, self.cutoff = float(prefactor), cutoff


class VerletListCoulombTruncated:
    """The `coulomb` interaction chemlab registers whenever coulomb_cutoff > 0 (gromacs_topology.py:866-878; rim135, dacron,
    pccg_lj ship coulomb_cutoff=0.9).  Every shipped coarse-grained system is NEUTRAL bead by bead (all q = 0), so the term is
    identically zero: this class keeps the label and the energy column (0.0) without touching the pair kernel.  A system with a
    non-zero charge raises NotImplementedError at set-up: charged pairs are outside north_star."""
    def __init__(self, vl):
        self._vl = vl
        self._ctx = vl._system._ctx
        self._h = None
        self._pots = {}
        self._zero_energy = True

    def setPotential(self, type1, type2, potential):
        self._pots[(int(type1), int(type2))] = potential

    def _attach(self, e):
        q = np.asarray(self._ctx.props.get("q", []), float)
        if q.size and np.any(q != 0.0):
            raise NotImplementedError("VerletListCoulombTruncated: the system holds charged particles (%d with q != 0); Coulomb pair "
                                      "forces are outside the scope of the B200 engine (SURVEY 8f rank 3)" % int((q != 0.0).sum()))

    def computeEnergy(self):
        return 0.0


# constructible, NotImplementedError when attached/used (gromacs_topology.py:513-514 builds two of them always)
for _n in ("VerletListTabulatedCapped", "VerletListLennardJonesEnergyCapped", "VerletListMultiTabulated", "VerletListMultiMixedTabulated",
           "VerletListScaleTabulated", "VerletListDynamicResolutionTabulated",
           "VerletListDynamicResolutionLennardJones", "TabulatedCapped", "LennardJonesEnergyCapped", "MultiTabulated",
           "MultiMixedTabulated", "ScaleTabulated", "FixedPairListLambdaHarmonic", "FixedPairListLambdaTabulated",
           "FixedTripleListLambdaAngularHarmonic", "FixedTripleListLambdaTabulatedAngular", "FixedQuadrupleListLambdaTabulatedDihedral",
           "DihedralRB", "DihedralHarmonicNCos", "FixedQuadrupleListDihedralRB", "FixedQuadrupleListDihedralHarmonicNCos",
           "FixedQuadrupleListTypesDihedralRB", "FixedQuadrupleListTypesDihedralHarmonicNCos", "ParticlePairScaling"):
    globals()[_n] = not_in_scope("interaction." + _n)
del _n

"""TEST INFRASTRUCTURE: run the REFERENCE's own driver -- /root/reference/src/start_simulation.py with src/chemlab/*, src/tools.py,
src/app_args.py -- on top of this repo's espressopp surface (oracle backend, no GPU).

The reference is Python 2.  Its nine source files are read where they lie, converted in memory by a handful of textual rules
(print statements, dict.iter*(), cPickle / ConfigParser names, `except X, e`, the two integer divisions that compute the number of
outer iterations) and written to a scratch directory OUTSIDE the repository together with a module of Python-2 builtins
(list-returning map / filter / zip, xrange, execfile).  `import espressopp` then resolves to chemlab_b200.espressopp -- the one-line
switch INTEGRATION.md describes -- and the driver runs unchanged: arg file, topology, reactions, force field, thermostat, observers,
main loop, products.  What the image lacks is stood in for here: h5py and the two H5MD dumpers (no HDF5 library), the MPI module.

usage: python tests/ref_driver_harness.py <example directory> <working directory> [driver arguments ...]
Runs only where /root/reference is mounted (this container)."""
import os
import re
import shutil
import sys
import tempfile
import types

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["start_simulation.py", "tools.py", "app_args.py", "chemlab/__init__.py", "chemlab/reaction_parser.py", "chemlab/reaction_setup.py",
         "chemlab/reaction_post_process.py", "chemlab/files_io.py", "chemlab/gromacs_topology.py"]
SHIM = '''
import builtins as _b
def map(f, *a): return list(_b.map(f, *a))
def filter(f, a): return list(_b.filter(f, a))
def zip(*a): return list(_b.zip(*a))
xrange = range
def execfile(path, g=None, l=None):
    with open(path) as fh:
        code = compile(fh.read(), path, "exec")
    exec(code, g if g is not None else {}, l)
import collections as _c, collections.abc as _ca, configparser as _cp, functools as _ft
if not hasattr(_cp, "SafeConfigParser"):
    _cp.SafeConfigParser = _ft.partial(_cp.ConfigParser, strict=False, interpolation=None)
for _n in ("Iterable", "Mapping", "Sequence"):
    if not hasattr(_c, _n): setattr(_c, _n, getattr(_ca, _n))
'''


def convert(text):
    out = []
    for line in text.split("\n"):
        m = re.match(r"^(\s*)print (?!\()(.*)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()").replace(".iterkeys()", ".keys()")
        line = re.sub(r"^(\s*)import cPickle\s*$", r"\1import pickle as cPickle", line)
        line = re.sub(r"^(\s*)import ConfigParser\s*$", r"\1import configparser as ConfigParser", line)
        line = re.sub(r"except (\w+), (\w+):", r"except \1 as \2:", line)
        line = line.replace("sim_step = args.run / integrator_step", "sim_step = args.run // integrator_step")     # Python-2 integer division
        out.append(line)
    return "from _py2shim import *\n" + "\n".join(out)


def build(tmp):
    os.makedirs(os.path.join(tmp, "chemlab"), exist_ok=True)
    for d in (tmp, os.path.join(tmp, "chemlab")):
        with open(os.path.join(d, "_py2shim.py"), "w") as f:
            f.write(SHIM)
    for name in FILES:
        with open(os.path.join(tmp, name), "w") as f:
            f.write(convert(open(os.path.join(REF, name)).read()))
    return tmp


class _H5Group(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}

    def _walk(self, k, create=False):
        node = self
        for part in [p for p in k.split("/") if p]:
            if not dict.__contains__(node, part):
                if not create:
                    raise KeyError(k)
                dict.__setitem__(node, part, _H5Group())
            node = dict.__getitem__(node, part)
        return node

    def create_group(self, name): return self._walk(name, True)
    def create_dataset(self, name, *a, **k): return self._walk(name, True)
    def __getitem__(self, k): return self._walk(k)

    def __contains__(self, k):
        try:
            self._walk(k)
            return True
        except KeyError:
            return False


class _H5File(_H5Group):
    def __init__(self, *a, **k): super().__init__()
    def close(self): pass
    def flush(self): pass


class DumpH5MD:
    def __init__(self, system, filename, **kw): self.ndump = 0
    def dump(self, *a): self.ndump += 1
    def flush(self): pass
    def close(self): pass
    def getTimers(self): return []


class DumpTopology:
    def __init__(self, system, integrator, traj): self.observed, self.static = [], []
    def observe_tuple(self, lst, name): self.observed.append(name)
    observe_triple = observe_quadruple = observe_tuple
    def add_static_tuple(self, lst, name): self.static.append(name)
    add_static_triple = add_static_quadruple = add_static_tuple
    def dump(self): pass
    def update(self): pass
    perform_action = dump
    def get_timers(self): return []


def main(argv):
    example, work, driver_args = argv[0], argv[1], argv[2:]
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    tmp = build(tempfile.mkdtemp(prefix="chemlab_ref_driver_"))
    sys.path.insert(0, os.path.join(tmp, "chemlab")); sys.path.insert(0, tmp)      # the reference's implicit relative imports
    import chemlab_b200.espressopp as es
    import chemlab_b200.espressopp._context as C
    if os.environ.get("CHEMLAB_HARNESS_BACKEND", "oracle") != "gpu":      # "gpu": the CUDA engine itself (needs a B200 and the reference tree)
        from oracle.engine_adapter import OracleEngine
        C.Engine = OracleEngine
    sys.modules["espressopp"] = es
    for sub in ("analysis", "integrator", "interaction", "storage", "bc", "esutil", "io", "tools"):
        sys.modules["espressopp." + sub] = getattr(es, sub)
    sys.modules["espressopp.tools.convert"] = es.tools.convert
    sys.modules["espressopp.tools.convert.gromacs"] = es.tools.convert.gromacs
    h5py = types.ModuleType("h5py"); h5py.File = _H5File
    sys.modules["h5py"] = h5py
    es.io.DumpH5MD, es.io.DumpTopology = DumpH5MD, DumpTopology
    mpi = types.ModuleType("MPI")
    mpi.COMM_WORLD = types.SimpleNamespace(size=1, rank=0)
    sys.modules["MPI"] = mpi
    if os.path.abspath(example) != os.path.abspath(work):
        shutil.copytree(example, work)
    os.chdir(work)
    sys.argv = ["start_simulation.py"] + list(driver_args)
    import start_simulation as reference_driver
    try:
        reference_driver.main()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("REFERENCE_DRIVER_OK")


if __name__ == "__main__":
    main(sys.argv[1:])

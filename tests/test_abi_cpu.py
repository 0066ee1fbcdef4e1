"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/chemlab_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from chemlab_b200 import _lib, build
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound():
    from chemlab_b200 import _lib as binding
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "chemlab_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 45
    for name in sorted(declared):
        assert hasattr(L, name), "libchemlab_b200.so does not export %s" % name
        assert name in binding.SIGNATURES, "ctypes binding misses %s" % name
    assert set(binding.SIGNATURES) <= declared
    assert L.clb_abi_version() == 1


def test_reaction_spec_layout_matches_header():
    from chemlab_b200._lib import ReactionSpec
    # 8 int32 + 3 double + 5 int32, natural alignment (see struct clb_reaction_spec)
    assert C.sizeof(ReactionSpec) == 8 * 4 + 3 * 8 + 5 * 4 + 4
    assert ReactionSpec.rate.offset == 32 and ReactionSpec.list.offset == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chemlab_b200 import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine([10.0, 10.0, 10.0], 2.5, 0.3)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    bad = []
    for dp, dn, fn in os.walk(os.path.join(ROOT, "chemlab_b200")):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", src, flags=re.M):
                    bad.append(f)
    assert not bad, "product files reference the oracle: %s" % bad


def test_every_entry_point_rejects_a_null_engine():
    """Error convention of the boundary: a call on a NULL engine handle returns an error code (or 0 / NULL for the value
    getters) instead of touching memory.  Runs without a GPU, in a child process so that a crash is reported as a failure."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent("""
        import sys, ctypes as C
        sys.path.insert(0, %r)
        from chemlab_b200 import _lib
        L = _lib.load()
        for name, (res, args) in _lib.SIGNATURES.items():
            vals = [a(0) if a in (C.c_int, C.c_int64, C.c_double, C.c_uint64, C.c_int32) else None for a in args]
            print(name, flush=True)
            r = getattr(L, name)(*vals)
            if name.startswith("clb_") and res is C.c_int and args and args[0] is C.c_void_p and name not in ("clb_nccl_unique_id",):
                assert r != 0, (name, r)
        print("NULL_OK")
        """ % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "NULL_OK" in p.stdout, (p.stdout.splitlines()[-2:], p.stderr[-1500:])

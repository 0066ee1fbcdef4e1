"""Build the CUDA engine in-tree: nvcc -> chemlab_b200/lib/libchemlab_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "engine.cu")
OUT = os.path.join(HERE, "lib", "libchemlab_b200.so")


def _newest_source():
    d = os.path.join(HERE, "csrc")
    files = [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(HERE, "..", "include", "chemlab_b200.h")]
    return max(os.path.getmtime(f) for f in files)


def nccl_flags():
    """nccl.h is needed for the types only: libnccl is resolved with dlopen at clb_comm_init time
    (csrc/engine_comm.inl), so the engine neither links against it nor needs it for single-GPU use."""
    if os.path.exists("/usr/include/nccl.h"):
        return ["-ldl"]
    try:
        import nvidia.nccl as m
        inc = os.path.join(os.path.dirname(m.__file__), "include")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return ["-I" + inc, "-ldl"]
    except Exception:
        pass
    raise RuntimeError("nccl.h not found")


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC] + nccl_flags()
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)

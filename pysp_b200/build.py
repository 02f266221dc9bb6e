"""Build libpysp_b200.so in-tree with nvcc for sm_100a (`python -m pysp_b200.build`)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "pysp_b200.cu")
OUT = os.path.join(HERE, "libpysp_b200.so")
# -fmad=false: every float32 op on the path is rounded individually (see csrc/pysp_common.cuh)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(HERE, "..", "include", "pysp_b200.h")]


def build(force=False, verbose=False, out=OUT, defines=()):
    """`out` / `defines` build an A/B variant next to the product library (tools/kbench.py); the package itself
    only ever loads libpysp_b200.so."""
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in sources()):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    subprocess.check_call(cmd, cwd=os.path.join(HERE, "csrc"))
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))

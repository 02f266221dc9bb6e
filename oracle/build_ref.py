"""Recipe: build the reference's Cython homogeneity-map extension with gcc.

TEST INFRASTRUCTURE ONLY.  The reference (`/root/reference`, read-only) ships an MSVC-only
setup.py (setup.py:12-18); this recipe cythonizes `debayer/ahd_homogeneity_cython.pyx` from where
it lies and writes every output (generated C, shared object) into `oracle/_ref/` (git-ignored).
Nothing from the reference is copied into the repository.  Flags: -O2 -fopenmp -ffp-contract=off
(no fast-math: the count map must be IEEE so that it can serve as a parity pin).

Usage:  python oracle/build_ref.py [/root/reference]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(reference_root="/root/reference", force=False, module=("debayer", "ahd_homogeneity_cython")):
    """`module` = (sub-package, name): the homogeneity map (default) or ("dng_warp_corr", "dng_warp_rectilinear_coords"),
    the reference's other native component (the DNG WarpRectilinear coordinate table)."""
    pyx = os.path.join(reference_root, module[0], module[1] + ".pyx")
    if not os.path.exists(pyx):
        return None
    import numpy
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, module[1] + ext)
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(pyx):
        return so
    c_file = os.path.join(OUT, module[1] + ".c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file])
    cmd = ["/usr/bin/gcc", "-shared", "-fPIC", "-O2", "-fopenmp", "-ffp-contract=off",
           "-Wno-unused-function", "-Wno-cpp",
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           c_file, "-o", so]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    print(build(root, force=True))
    print(build(root, force=True, module=("dng_warp_corr", "dng_warp_rectilinear_coords")))

"""Recipe: make the UNMODIFIED reference importable on the GPU box, for the CPU arm of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container only.  The bench contract's reference
arm (`bench.py --impl reference`) has to time the reference's own code on the GPU box's host cores, so this recipe --
run by `__graft_entry__.build()` when the reference is mounted -- installs the reference's Python package under
`baseline/_ref/pySP/` (git-ignored: never committed; not gpurun-ignored: it travels with the snapshot) together with
its compiled Cython extension (`oracle/build_ref.py` -> `oracle/_ref/`).  The reference's own setup.py cannot be used
(MSVC flags and back-slash paths, setup.py:12-18; no package metadata), so the install is a plain copy of its `*.py` /
`*.pyx` files; nothing is modified.  `oracle/ref_harness.py` imports it from there when `/root/reference` is absent.

Usage:  python oracle/install_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "..", "baseline", "_ref", "pySP")


def install(reference_root="/root/reference"):
    if not os.path.exists(os.path.join(reference_root, "debayer", "ahd.py")):
        return None
    dest = os.path.abspath(DEST)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    shutil.copytree(reference_root, dest, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"),
                    copy_function=shutil.copy2)
    return dest


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))

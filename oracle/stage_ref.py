#!/usr/bin/env python
"""Stage the UNMODIFIED reference's hot-path modules into the git-ignored oracle/_ref/ so that the reference itself
(not the port) can be timed on the GPU box's host cores (bench.py --impl reference, cpu_baseline.kind "reference").

The reference is pure Python: "building" it is a file copy.  Nothing is edited; the import shim that supplies the
absent mpi4py / h5py and numpy's removed `np.int` alias is tests/golden/refshim.py, pointed at the staged tree through
PSI_REFERENCE_ROOT.  oracle/_ref/ is listed in .gitignore (reference sources never enter the history) but not in
.gpurunignore (it travels to the GPU box like a built .so).  Run by __graft_entry__.build() whenever /root/reference
is present:
    python oracle/stage_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
MODULES = ("__init__", "tanhsinh", "sizedensity", "power_bump", "psi_dispersion", "complex_roots", "psi_mode",
           "psi_mode_mpi", "complex_roots_mpi", "psi_grid_refine")


def stage(src_root="/root/reference"):
    src = os.path.join(src_root, "psitools")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(DEST, "psitools")
    os.makedirs(dst, exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m + ".py"), os.path.join(dst, m + ".py"))
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as fh:
        fh.write(src + "\n")
    return True


def staged():
    return os.path.isfile(os.path.join(DEST, "psitools", "psi_dispersion.py"))


if __name__ == "__main__":
    ok = stage(*sys.argv[1:2])
    print("staged" if ok else "reference tree not found", DEST)

"""Put the UNMODIFIED reference model package where the GPU box can run it: baseline/_ref/models.

`/root/reference` exists only in the build container; `baseline/_ref/` is git-ignored (the repository's history holds no
reference source) but NOT gpurun-ignored, so the copy travels with the snapshot and `bench.py --impl reference` /
`cpu_baseline` time the reference's stock classes on the GPU box's host cores.  The reference is pure Python: the
"install" is a verbatim copy of its `models/` package (pip cannot install it: the tree has no setup.py / pyproject).
Its one third-party arithmetic dependency, UMNN==1.0, is absent from the image and the wheelhouse; `oracle/UMNN.py`
(restatement, parity unpinned at that boundary) is put on sys.path in its place by baseline/ref_arm.py.

    python baseline/install_ref.py        (called by __graft_entry__.build() when /root/reference is present)
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/models"
DST = os.path.join(HERE, "_ref", "models")


def install(verbose=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)          # GPU box: use what travelled with the snapshot
    if os.path.isdir(DST):
        cmp = filecmp.dircmp(SRC, DST, ignore=["__pycache__"])
        if not (cmp.left_only or cmp.right_only or cmp.diff_files):
            return True
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(HERE, "_ref", "INSTALLED.txt"), "w") as f:
        f.write("verbatim copy of /root/reference/models (AWehenkel/Graphical-Normalizing-Flows), made by baseline/install_ref.py\n")
    if verbose:
        print("installed", DST)
    return True


if __name__ == "__main__":
    ok = install(verbose=True)
    sys.exit(0 if ok else 1)

"""Build the host SIMT-simulator flavour of the kernels (tests only; see csrc/cpu_emu.h)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "graphical-normalizing-flows_b200", "csrc")
OUT = os.path.join(HERE, "libgnf_emu.so")


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    # tc_*.cu (tcgen05 kernels) compile to error-returning stubs under GNF_EMU: they have no host flavour
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-x", "c++", "-std=c++20", "-O2", "-DGNF_EMU", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas",
           "-o", OUT] + srcs
    subprocess.run(cmd, check=True, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force=True))

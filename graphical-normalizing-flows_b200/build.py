"""In-tree build of libgnf_sm100.so: nvcc, sm_100a only (cross-compiles without a GPU)."""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgnf_sm100.so")
OBJ = os.path.join(HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False, dev=False):
    """Product library, or (dev=True) libgnf_sm100_dev.so = the same sources with -DGNF_DEVTOOLS: clock-stamp traces,
    ablation switches and tiling overrides for scripts/ -- never loaded by the package."""
    global OUT, OBJ
    if dev:
        saved = (OUT, OBJ, list(NVCC_FLAGS))
        OUT, OBJ = os.path.join(HERE, "libgnf_sm100_dev.so"), os.path.join(HERE, "build_dev")
        NVCC_FLAGS.append("-DGNF_DEVTOOLS")
        try:
            return _build(force, verbose)
        finally:
            OUT, OBJ = saved[0], saved[1]
            NVCC_FLAGS[:] = saved[2]
    return _build(force, verbose)


def _build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(os.path.dirname(HERE), "include", "gnf.h")]
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=CSRC)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", OUT] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv, dev="--dev" in sys.argv))

"""Build the C-ABI CUDA library ``aline_b200/lib/libaline_b200.so`` in-tree with nvcc (sm_100a only).

    python -m aline_b200.build [--force]

nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored but travels to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# ALINE_BUILD_TRACE=1: a second, instrumented library (per-phase clock stamps in the candidate-query stream; used only
# by tools/trace_q4.py through ALINE_B200_LIB) -- never the product build
TRACE = os.environ.get("ALINE_BUILD_TRACE") == "1"
# ALINE_BUILD_DEFS="-DX -DY": an experimental library libaline_b200_exp.so next to the product one (development A/B runs
# through ALINE_B200_LIB, like the trace build)
EXP_DEFS = os.environ.get("ALINE_BUILD_DEFS", "").split()
LIB = os.path.join(LIB_DIR, "libaline_b200_trace.so" if TRACE else "libaline_b200_exp.so" if EXP_DEFS else "libaline_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC",
] + (["-DALINE_Q4_TRACE"] if TRACE else []) + EXP_DEFS


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libaline_b200.so for sm_100a)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ (one object per file, in parallel) and link the shared library."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build_trace" if TRACE else "build_exp" if EXP_DEFS else "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(f) for f in _deps() if not f.endswith(".cu") or f == src)):
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp", *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libaline_b200.so failed")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

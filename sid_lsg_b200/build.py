"""Builds the C-ABI CUDA library `sid_lsg_b200/_C/libsidlsg.so` for sm_100a with nvcc (in-tree, no JIT cache).

`python -m sid_lsg_b200.build [--force]`.  Objects are rebuilt only when their source (or a header) is newer.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libsidlsg.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", CSRC, "-I", INCLUDE]
# extra compile flags for instrumented builds, e.g. SIDLSG_NVCC_EXTRA=-DSIDLSG_GEMM_TRACE (scripts/trace_gemm.py); use --force
NVCC_FLAGS += os.environ.get("SIDLSG_NVCC_EXTRA", "").split()


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built on this machine")
    return nvcc


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False):
    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if os.path.isdir(INCLUDE):
        headers += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hdr_mtime = max([os.path.getmtime(h) for h in headers] + [os.path.getmtime(__file__)])
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, "obj", src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_mtime):
            jobs.append((s, o))

    def run(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

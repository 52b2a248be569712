"""Builds libLBAudioDetectiveCUDA.so in-tree: nvcc (sm_100a) for the kernels, gcc for the host-side C API layer."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libLBAudioDetectiveCUDA.so")
OBJ = os.path.join(PKG, "build")

CU = ["lbad_extract.cu", "lbad_search.cu", "lbad_synth.cu"]
C = ["LBAudioDetective.c", "LBAudioDetectiveFingerprint.c", "LBAudioDetectiveDatabase.c", "lbad_support.c"]
HDRS = ["lbad_cuda.h", "lbad_common.cuh", "lbad_math.cuh", "lbad_host.h"]

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "186"]
CC_FLAGS = ["-std=gnu11", "-O2", "-fPIC", "-Wall", "-Wextra", "-fvisibility=hidden"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(CSRC, h) for h in HDRS] + [os.path.join(PKG, "..", "include", h) for h in os.listdir(os.path.join(PKG, "..", "include"))]
    srcs = [os.path.join(CSRC, s) for s in CU + C]
    if not force and not _stale(LIB, srcs + hdrs + [os.path.abspath(__file__)]):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for s in CU:
        o = os.path.join(OBJ, s + ".o"); objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s in C:
        o = os.path.join(OBJ, s + ".o"); objs.append(o)
        cmd = [os.environ.get("CC", "gcc")] + CC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), out))
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lm"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(link), r.stdout))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

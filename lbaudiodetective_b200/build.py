"""Builds libLBAudioDetectiveCUDA.so in-tree: nvcc (sm_100a) for the kernels, gcc for the host-side C API layer."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libLBAudioDetectiveCUDA.so")
OBJ = os.path.join(PKG, "build")

CU = ["lbad_extract.cu", "lbad_search.cu", "lbad_synth.cu", "lbad_resample.cu", "lbad_frame.cu"]
C = ["LBAudioDetective.c", "LBAudioDetectiveFingerprint.c", "LBAudioDetectiveDatabase.c", "LBAudioDetectiveFrame.c", "lbad_support.c", "lbad_resample_design.c"]
HDRS = ["lbad_cuda.h", "lbad_common.cuh", "lbad_math.cuh", "lbad_host.h"]

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "186"]
CC_FLAGS = ["-std=gnu11", "-O2", "-fPIC", "-Wall", "-Wextra", "-fvisibility=hidden", "-ffp-contract=off"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_variant(name: str, defines) -> str:
    """Kernel A/B experiments: the same sources with extra -D flags, linked into scripts/_bin/libvariant_<name>.so (never the shipped
    library; bench.py loads one when LBAD_LIBRARY points at it)."""
    out_dir = os.path.join(PKG, "..", "scripts", "_bin"); os.makedirs(out_dir, exist_ok=True)
    obj_dir = os.path.join(out_dir, "obj_" + name); os.makedirs(obj_dir, exist_ok=True)
    objs = []
    for s in CU:
        o = os.path.join(obj_dir, s + ".o"); objs.append(o)
        subprocess.run([_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, s), "-o", o], check=True)
    for s in C:
        o = os.path.join(obj_dir, s + ".o"); objs.append(o)
        subprocess.run([os.environ.get("CC", "gcc")] + CC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o], check=True)
    lib = os.path.join(out_dir, "libvariant_%s.so" % name)
    subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lm"], check=True)
    return lib


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
    if "--variant" in sys.argv:          # python build.py --variant NAME -DFOO=1 -DBAR=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

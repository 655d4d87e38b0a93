"""Builds sdflib_b200/libsdfb200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python -m sdflib_b200.build [--force]

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so travels to
the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).

Per-file flags matter:
  * octree_build.cu / exact_build.cu / exact_query.cu and the EXACT object of octree_query.cu are
    compiled with -fmad=false: they reproduce the reference's x86-64 (FMA-less) float arithmetic bit
    for bit, because octree topology is decided by those floats;
  * the FAST object of octree_query.cu keeps FMA contraction on (Horner evaluation of the leaf
    polynomial, within 1e-5 of the exact order and never used for a topology decision).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsdfb200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_FLAGS = "-fPIC,-fopenmp,-ffp-contract=off,-O3"
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", HOST_FLAGS] + ARCH
NO_FMA = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"]

# (source, object name, extra flags)
UNITS = [
    ("mesh_host.cpp", "mesh_host.o", ["-x", "cu"] + NO_FMA),
    ("bin_io.cpp", "bin_io.o", ["-x", "cu"]),
    ("capi.cpp", "capi.o", ["-x", "cu"]),
    ("shard.cpp", "shard.o", ["-x", "cu"]),
    ("host_mem.cpp", "host_mem.o", ["-x", "cu"]),
    ("query_host.cpp", "query_host.o", ["-x", "cu"]),
    ("multi_device.cpp", "multi_device.o", ["-x", "cu"]),
    ("fixtures.cpp", "fixtures.o", ["-x", "cu"] + NO_FMA),
    ("mesh_device.cu", "mesh_device.o", NO_FMA),
    ("bvh_device.cu", "bvh_device.o", NO_FMA),
    ("octree_build.cu", "octree_build.o", NO_FMA),
    ("octree_cont.cu", "octree_cont.o", NO_FMA),
    ("octree_query.cu", "octree_query_fast.o", []),
    ("octree_query.cu", "octree_query_exact.o", ["-DSDFB_QUERY_EXACT"] + NO_FMA),
]
UNITS += [("exact_build.cu", "exact_build.o", NO_FMA), ("exact_query.cu", "exact_query.o", NO_FMA)]
UNITY_LIB = os.path.join(HERE, "libSdfLibUnity.so")   # the reference's Unity plugin exports on top of the C-ABI


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libsdfb200.so cannot be built")
    return nvcc


def _ccbin():
    # The image exports CXX=/opt/gcc/bin/g++ (no libgomp.spec); the system compiler has OpenMP.
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", "sdfb200.h"), os.path.abspath(__file__)]
    nvcc = _nvcc()
    jobs = []
    for src, obj, extra in UNITS:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, obj)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + _ccbin() + COMMON + extra + ["-Xptxas", "-v", "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose:
                    print(" ".join(cmd[-4:]), file=sys.stderr)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError("nvcc failed on " + cmd[-3])
                with open(cmd[-1] + ".ptxas.log", "w") as f:   # registers / spills, see DESIGN.md
                    f.write(r.stderr)
    objs = [os.path.join(OBJ, o) for _, o, _ in UNITS]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + _ccbin() + ["-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libsdfb200.so failed")
    unity_src = os.path.join(CSRC, "unity_abi.cpp")
    if force or _stale(UNITY_LIB, [unity_src, LIB] + headers):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", unity_src, "-o", UNITY_LIB, "-L" + HERE, "-lsdfb200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libSdfLibUnity.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

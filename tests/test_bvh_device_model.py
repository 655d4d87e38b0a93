"""The device-side BVH build (sdflib_b200/csrc/bvh_build.cuh) run from its CUDA source on the CPU (tests/cpp/simt_bvh_main.cpp:
threads of a CTA = host threads, __syncthreads / warp collectives = barriers, the real launch sequence): it must reproduce
the host builder — which calls libstdc++'s own std::sort routines — node for node and bit for bit, and its data-parallel
restatement of std::sort (Hoare partitions as ordered compactions, heap sort once the depth limit is spent, the final
insertion pass as ranks) must give the library's permutation for any depth limit. Runs without a GPU; the GPU test of the
same kernels is tests/test_gpu_bvh.py."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def exe(tmp_path_factory, sdf):
    out = str(tmp_path_factory.mktemp("simt_bvh") / "simt_bvh_main")
    lib_dir = os.path.join(ROOT, "sdflib_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # small CTAs and a 256-element shared-memory limit: meshes of a few thousand triangles reach the global partition rounds;
    # BVH_HOST_CHAIN=100: the centre sums of levels with >= 100 triangles per node are offered to the host path (odd levels take it)
    cmd = [cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-w", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(lib_dir, "csrc"), "-DBVH_SMALL_MAX=256", "-DBVH_BIG_THREADS=64", "-DBVH_SMALL_THREADS=64", "-DBVH_HOST_CHAIN=100", "-x", "c++",
           os.path.join(ROOT, "tests", "cpp", "simt_bvh_main.cpp"), "-o", out, "-L" + lib_dir, "-lsdfb200", "-Wl,-rpath," + lib_dir, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.mark.parametrize("args", [("0", "0"), ("0", "1"), ("1", "1"), ("2", "0"), ("2", "1"), ("3", "0"), ("3", "1"),
                                  ("2", "1", "1"), ("2", "1", "2"), ("2", "1", "3"), ("2", "1", "17"), ("3", "1", "257"), ("3", "0", "1023")])
def test_tree_equals_host_builder(exe, args):
    r = subprocess.run([exe, "mesh", *args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "identical" in r.stdout, (args, r.stdout, r.stderr)


@pytest.mark.parametrize("n,distinct,depth", [(17, 3, 8), (257, 2, 16), (600, 600, 18), (1500, 12, 0), (1500, 12, 1), (2500, 40, 5),
                                              (4000, 3, 22), (5000, 5000, 24), (300, 1, 16), (256, 7, 2), (1000, 7, 3)])
def test_sort_equals_libstdcxx(exe, n, distinct, depth):
    """depth = remaining introsort depth the range starts with: small values force the heap-sort fallback (in shared memory
    and, for ranges above the shared-memory limit, in global memory)."""
    for seed in (1, 2, 3):
        r = subprocess.run([exe, "sort", str(n), str(distinct), str(depth), str(seed)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and r.stdout.startswith("ok sort"), (n, distinct, depth, seed, r.stdout, r.stderr)

"""The C++ drop-in surface (include/SdfLib/*.h) and the Unity C exports (include/sdfb200_unity.h).

CPU: the README-style program compiles and links against our headers + libsdfb200.so, and fails LOUDLY (exception
text, non-zero exit) on a box without a GPU — there is no CPU fallback. GPU: the program's results equal the ctypes
binding's for the same inputs, bit for bit. glm comes from oracle/shim (a user has the real glm)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal


def compile_dropin(tmp_path):
    exe = str(tmp_path / "dropin_main")
    lib_dir = os.path.join(ROOT, "sdflib_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-o", exe, "-L" + lib_dir, "-lsdfb200", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_dropin_compiles_links_and_has_no_cpu_fallback(sdf, tmp_path):
    exe = compile_dropin(tmp_path)
    if sdf.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, str(tmp_path / "out.bin"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_unity_library_exports_the_reference_symbols(sdf):
    lib = C.CDLL(os.path.join(ROOT, "sdflib_b200", "libSdfLibUnity.so"))
    for name in ("saveSdf", "loadSdf", "createExactOctreeSdf", "createOctreeSdf", "getDistance", "getDistanceAndGradient",
                 "getBBMinPoint", "getBBSize", "getStartGridSize", "getOctreeDataSize", "getOctreeData", "deleteSdf"):
        assert hasattr(lib, name), name      # src/tools/SdfLibUnity/SdfExportFunc.h:16-58
    lib.loadSdf.restype = C.c_void_p
    assert lib.loadSdf(b"/nonexistent/file.bin") is None


def read_vec(f, dtype, cols=None):
    n = int(np.frombuffer(f.read(8), np.uint64)[0])
    item = np.dtype(dtype).itemsize * (cols or 1)
    a = np.frombuffer(f.read(n * item), dtype)
    return a.reshape(n, cols) if cols else a


@pytest.mark.gpu
def test_dropin_program_matches_the_python_binding(sdf, tmp_path):
    exe = compile_dropin(tmp_path)
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(out), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    with open(out, "rb") as f:
        verts, idx, box, pts = read_vec(f, np.float32, 3), read_vec(f, np.uint32), read_vec(f, np.float32), read_vec(f, np.float32, 3)
        d_oct, g_oct, d_ex, g_ex = read_vec(f, np.float32), read_vec(f, np.float32, 3), read_vec(f, np.float32), read_vec(f, np.float32, 3)
        oct_words, ex_nodes, ex_sets, ex_masks = read_vec(f, np.uint32), read_vec(f, np.uint32, 2), read_vec(f, np.uint32), read_vec(f, np.uint8)
        hdr = read_vec(f, np.float32)
    mesh, bb = sdf.Mesh(verts, idx), sdf.BoundingBox(box[:3], box[3:])
    o = sdf.OctreeSdf(mesh, bb, 5, 2, 1e-3, sdf.OctreeSdf.NO_CONTINUITY, 8)
    e = sdf.ExactOctreeSdf(mesh, bb, 5, 2, 16, 8)
    assert np.array_equal(o.getOctreeData(), oct_words) and np.array_equal(e.getOctreeData(), ex_nodes)
    assert np.array_equal(e.getTrianglesSets(), ex_sets) and np.array_equal(e.getTrianglesMasks(), ex_masks)
    d, g = o.getDistance(pts, gradient=True)
    assert_bit_equal(d, d_oct); assert_bit_equal(g, g_oct)
    d, g = e.getDistance(pts, gradient=True)
    assert_bit_equal(d, d_ex); assert_bit_equal(g, g_ex)
    assert_bit_equal(hdr[:2], np.float32([o.info().value_range, o.info().min_border_value]))
    # the files the C++ program saved are the same bytes the Python binding writes
    a = str(tmp_path / "py_octree.bin")
    assert o.saveToFile(a) and open(a, "rb").read() == open(tmp_path / "octree.bin", "rb").read()


@pytest.mark.gpu
def test_unity_exports_round_trip(sdf, tmp_path):
    lib = C.CDLL(os.path.join(ROOT, "sdflib_b200", "libSdfLibUnity.so"))

    class Vec3(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]
    for fn in ("createExactOctreeSdf", "createOctreeSdf", "loadSdf"):
        getattr(lib, fn).restype = C.c_void_p
    lib.getDistance.restype = C.c_float
    lib.getDistanceAndGradient.restype = C.c_float
    lib.getBBMinPoint.restype = Vec3
    lib.getBBSize.restype = Vec3
    lib.getStartGridSize.restype = C.c_uint32
    lib.getOctreeDataSize.restype = C.c_uint32
    v, i = sdf.meshes.isosphere(2)
    v = sdf.meshes.displace(v)
    box = sdf.meshes.bounding_box_with_margin(v)
    args = [v.ctypes.data_as(C.c_void_p), C.c_uint32(len(v)), i.ctypes.data_as(C.c_void_p), C.c_uint32(i.size)] + [C.c_float(x) for x in box]
    ex = C.c_void_p(lib.createExactOctreeSdf(*args, C.c_uint32(2), C.c_uint32(5), C.c_uint32(16), C.c_uint32(1)))
    oc = C.c_void_p(lib.createOctreeSdf(*args, C.c_uint32(2), C.c_uint32(5), C.c_float(1e-3), C.c_uint32(1)))
    assert ex.value and oc.value
    ref_e = sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 5, 2, 16, 1)
    p = np.float32([0.31, -0.22, 0.17])
    g = Vec3()
    d = lib.getDistanceAndGradient(ex, C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]), C.byref(g))
    rd, rg = ref_e.getDistance(p, gradient=True)
    assert np.float32(d) == np.float32(rd) and np.array_equal(np.float32([g.x, g.y, g.z]), rg)
    assert np.float32(lib.getDistance(ex, C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]))) == np.float32(rd)
    assert lib.getStartGridSize(oc) == 4 and lib.getStartGridSize(ex) == 0
    n = lib.getOctreeDataSize(oc)
    data = np.zeros(n, np.uint32)
    lib.getOctreeData(oc, data.ctypes.data_as(C.c_void_p))
    assert n > 64 and data[:64].any()
    mn, sz = lib.getBBMinPoint(oc), lib.getBBSize(oc)
    assert abs(sz.x - sz.y) < 1e-6 and mn.x < v[:, 0].min()
    path = str(tmp_path / "unity.bin").encode()
    lib.saveSdf(ex, path)
    again = C.c_void_p(lib.loadSdf(path))
    assert again.value and np.float32(lib.getDistance(again, C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]))) == np.float32(rd)
    for h in (ex, oc, again):
        lib.deleteSdf(h)

"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/sdfb200.h
declares, validates arguments, reports errors as codes (never exits), and refuses to compute without
a CUDA device (no CPU fallback). Host-side pieces (TriangleData precompute, fixtures) are checked
against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal, golden, displaced_sphere


def test_library_exports_every_declared_symbol(sdf):
    header = open(os.path.join(ROOT, "include", "sdfb200.h")).read()
    declared = set(re.findall(r"\b(sdfb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    from sdflib_b200 import _capi
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = sdf.lib()
    for name in declared:
        assert hasattr(lib, name), f"libsdfb200.so does not export {name}"
    assert lib.sdfb200_version() == 100


def test_info_struct_layout_matches_header(sdf):
    from sdflib_b200 import _capi
    # int32 + 6 f32 + int32 + u32 + 2 f32 + 6 u32 (+pad) + 4 u64 + int32 (+pad)
    assert C.sizeof(_capi.Info) == 112
    assert C.sizeof(_capi.BuildStats) == 7 * 8 + 4 * 8


def test_argument_validation_returns_codes(sdf):
    from sdflib_b200 import _capi
    L = sdf.lib()
    v, i = sdf.meshes.isosphere(1)
    box = np.float32([-2, -2, -2, 2, 2, 2])
    h = C.c_void_p()
    # null mesh
    assert L.sdfb200_build_octree(None, 0, None, 0, _capi.ptr(box), 5, 3, 1, C.c_float(1e-3), C.c_float(0), 1, 1, C.byref(h)) == _capi.ERR_INVALID
    assert b"mesh" in L.sdfb200_last_error()
    # index out of range
    bad = i.copy(); bad[5] = 10 ** 6
    assert L.sdfb200_build_octree(_capi.ptr(v), len(v), _capi.ptr(bad), bad.size, _capi.ptr(box), 5, 3, 1, C.c_float(1e-3), C.c_float(0), 1, 1, C.byref(h)) == _capi.ERR_INVALID
    # degenerate box
    flat = np.float32([0, 0, 0, 1, 0, 1])
    assert L.sdfb200_build_octree(_capi.ptr(v), len(v), _capi.ptr(i), i.size, _capi.ptr(flat), 5, 3, 1, C.c_float(1e-3), C.c_float(0), 1, 1, C.byref(h)) == _capi.ERR_INVALID
    # options of the reference API that are not built yet are reported, not silently changed
    assert L.sdfb200_build_octree(_capi.ptr(v), len(v), _capi.ptr(i), i.size, _capi.ptr(box), 5, 3, 1, C.c_float(1e-3), C.c_float(0), 0, 1, C.byref(h)) == _capi.ERR_UNSUPPORTED   # UNIFORM
    assert h.value is None


def test_file_errors(sdf, tmp_path):
    from sdflib_b200 import _capi
    L = sdf.lib()
    h = C.c_void_p()
    assert L.sdfb200_load(str(tmp_path / "missing.bin").encode(), C.byref(h)) == _capi.ERR_IO
    p = tmp_path / "garbage.bin"
    p.write_bytes(b"\x01\x07\x00\x00\x00" + b"\x00" * 40)
    assert L.sdfb200_load(str(p).encode(), C.byref(h)) == _capi.ERR_IO
    g = golden("small_structures.npz")
    t = tmp_path / "truncated.bin"
    t.write_bytes(g["octree_bin"].tobytes()[:1000])
    assert L.sdfb200_load(str(t).encode(), C.byref(h)) == _capi.ERR_IO
    # SdfFunction::loadFromFile returns nullptr on file errors (src/sdf/SdfFunction.cpp:47-51)
    assert sdf.SdfFunction.loadFromFile(str(tmp_path / "missing.bin")) is None


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(sdf):
    """Without a CUDA device construction and queries fail loudly with SDFB200_ERR_CUDA."""
    from sdflib_b200 import _capi
    v, i = sdf.meshes.isosphere(1)
    with pytest.raises(sdf.SdfB200Error) as e:
        sdf.OctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox([-2, -2, -2], [2, 2, 2]), 4, 2)
    assert e.value.code == _capi.ERR_CUDA
    with pytest.raises(sdf.SdfB200Error) as e:
        sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox([-2, -2, -2], [2, 2, 2]), 4, 1, 16)
    assert e.value.code in (_capi.ERR_CUDA, _capi.ERR_UNSUPPORTED)
    assert sdf.device_count() == 0


def test_host_triangle_data_matches_oracle(sdf, port):
    """The product's sort-based edge pairing reproduces the reference's ordered-map walk bit for bit."""
    from sdflib_b200 import _capi
    for mesh in (displaced_sphere(3), (golden("kernels.npz")["tet_vertices"], golden("kernels.npz")["tet_indices"])):
        v, i = mesh
        out = np.empty((i.size // 3, 37), np.float32)
        _capi.check(sdf.lib().sdfb200_triangle_data(_capi.ptr(_capi.f32(v)), len(v), _capi.ptr(_capi.u32(i)), i.size, _capi.ptr(out)))
        assert_bit_equal(out, port.triangle_data(v, i), "TriangleData")


def test_host_triangle_data_non_manifold(sdf, port):
    """Open / duplicated-vertex meshes go through the non-manifold repair path (TriangleUtils.cpp:292-420)."""
    from sdflib_b200 import _capi
    v, i = displaced_sphere(2)
    # duplicate the vertices of the first 40 triangles so their edges only pair after the merge step
    i = i.copy()
    extra = []
    for t in range(40):
        for k in range(3):
            extra.append(v[i[3 * t + k]])
            i[3 * t + k] = len(v) + len(extra) - 1
    v2 = np.concatenate([v, np.float32(extra)])
    out = np.empty((i.size // 3, 37), np.float32)
    _capi.check(sdf.lib().sdfb200_triangle_data(_capi.ptr(v2), len(v2), _capi.ptr(i), i.size, _capi.ptr(out)))
    assert_bit_equal(out, port.triangle_data(v2, i), "TriangleData (non-manifold)")


def test_fixture_meshes(sdf, port):
    for s in (0, 1, 2, 4):
        v, i = sdf.meshes.isosphere(s)
        pv, pi = port.isosphere(s)
        assert_bit_equal(v, pv); assert np.array_equal(i, pi)
        assert i.size // 3 == 20 * 4 ** s
    v, i = sdf.meshes.config_mesh("M0")
    assert i.size // 3 == 320
    g = sdf.meshes.cell_centre_grid(np.float32([0, 0, 0, 1, 1, 1]), 4)
    assert g.shape == (64, 3) and np.allclose(g[1] - g[0], [0.25, 0, 0]) and np.allclose(g[0], 0.125)


def test_release_cached_memory_is_callable_without_a_device(sdf):
    """Trims the device-block and pinned-block caches; a no-op (and still OK) on a box without a GPU."""
    from sdflib_b200 import _capi
    assert _capi.lib().sdfb200_release_cached_memory() == _capi.OK


def test_bvh_builder_equals_serial_reference_order(sdf, tmp_path):
    """The product BVH (16-byte sort proxies, std::sort's recursion split across threads) against a serial restatement
    of tmd's _build_tree that moves 80-byte records with one std::sort per node: every node identical, for meshes full
    of sort-key ties (plain isospheres) and in generic position, at several thread counts."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "bvh_host_main")
    lib_dir = os.path.join(ROOT, "sdflib_b200")
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    cmd = [nvcc] + ccbin + ["-x", "cu", "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-I" + os.path.join(ROOT, "include"),
                            "-I" + os.path.join(lib_dir, "csrc"), os.path.join(ROOT, "tests", "cpp", "bvh_host_main.cpp"), "-o", exe,
                            "-L" + lib_dir, "-lsdfb200", "-Xlinker", "-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "sort"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.stdout, r.stderr)
    for subdivisions, displace, threads in ((2, 0, 4), (6, 0, 1), (6, 0, 8), (7, 0, 16), (7, 1, 3), (7, 1, 16)):
        r = subprocess.run([exe, str(subdivisions), str(displace), str(threads)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("ok "), (subdivisions, displace, threads, r.stdout, r.stderr)


def test_bin_loader_survives_corrupted_files(sdf, port, tmp_path):
    """Untrusted .bin input: byte flips in the header and in the arrays, truncations and wild 64-bit lengths must come
    back as error codes (or load, when the damage is harmless) — never a crash, hang or exit. Parsing and structure
    validation run on the host before a device is needed, so this runs without a GPU."""
    from sdflib_b200 import _capi, meshes
    L = sdf.lib()
    v, i = meshes.isosphere(2)
    box = np.float32([-1.3] * 3 + [1.3] * 3)
    port.build_octree(v, i, box, 4, 2, use_cache=False).save(str(tmp_path / "o.bin"))
    port.build_exact(v, i, box, 4, 2, min_tris=16, use_cache=False).save(str(tmp_path / "e.bin"))
    rng = np.random.default_rng(5)
    seen = set()
    for name in ("o.bin", "e.bin"):
        raw = (tmp_path / name).read_bytes()
        for t in range(240):
            b = bytearray(raw)
            mode = t % 4
            if mode == 0:
                for _ in range(rng.integers(1, 4)):
                    b[rng.integers(0, 200)] = rng.integers(0, 256)
            elif mode == 1:
                for _ in range(rng.integers(1, 6)):
                    b[rng.integers(0, len(b))] = rng.integers(0, 256)
            elif mode == 2:
                b = b[:rng.integers(0, len(b))]
            else:
                k = int(rng.integers(0, len(b) - 8))
                b[k:k + 8] = np.uint64(rng.integers(0, 2 ** 63)).tobytes()
            p = tmp_path / "m.bin"
            p.write_bytes(bytes(b))
            h = C.c_void_p()
            code = L.sdfb200_load(str(p).encode(), C.byref(h))
            seen.add(code)
            assert code in (_capi.OK, _capi.ERR_IO, _capi.ERR_INVALID, _capi.ERR_CUDA), code
            if h.value:
                L.sdfb200_free(h)
    assert _capi.ERR_IO in seen


def _random_meshes(rng, count):
    """Triangle soups that stress calculateMeshTriangleData: tiny vertex sets (edges shared by many triangles),
    spheres with duplicated vertices (open edges the repair merges) or missing triangles (true boundaries), large
    random soups. Degenerate triangles are left out (0/0 in the frame on every implementation)."""
    from sdflib_b200 import meshes
    for t in range(count):
        kind = t % 4
        if kind == 0:
            nv, nt = rng.integers(4, 12), rng.integers(1, 30)
            v, i = rng.standard_normal((nv, 3)), rng.integers(0, nv, size=nt * 3)
        elif kind == 1:
            v, i = meshes.isosphere(1)
            v, i = v.copy(), i.copy()
            for _ in range(rng.integers(1, 20)):
                c = rng.integers(0, i.size)
                v = np.vstack([v, v[i[c]][None] + (rng.standard_normal(3) * 1e-7)])
                i[c] = len(v) - 1
        elif kind == 2:
            v, i = meshes.isosphere(2)
            i = i.reshape(-1, 3)[rng.random(i.size // 3) > 0.2].ravel()
        else:
            nv, nt = rng.integers(50, 400), rng.integers(50, 2000)
            v, i = rng.standard_normal((nv, 3)), rng.integers(0, nv, size=nt * 3)
        tri = np.asarray(i).reshape(-1, 3)
        tri = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]
        if len(tri):
            yield np.ascontiguousarray(v, np.float32), np.ascontiguousarray(tri.ravel(), np.uint32)


def test_triangle_data_on_random_non_manifold_meshes(sdf, port):
    """The multi-threaded host TriangleData (sorted edge pairing, chunked by edge groups, non-manifold repair) against
    the oracle's serial map-based restatement, bit for bit, on meshes where every branch of the repair is taken."""
    from sdflib_b200 import _capi
    L = sdf.lib()
    for v, i in _random_meshes(np.random.default_rng(3), 80):
        out = np.empty((i.size // 3, 37), np.float32)
        assert L.sdfb200_triangle_data(_capi.ptr(v), len(v), _capi.ptr(i), i.size, _capi.ptr(out)) == _capi.OK
        want = port.triangle_data(v, i)
        same = (out.view(np.uint32) == want.view(np.uint32)) | (np.isnan(out) & np.isnan(want))
        assert same.all(), (len(v), i.size)


def test_oracle_triangle_data_equals_reference_on_random_meshes(port, ref):
    for v, i in _random_meshes(np.random.default_rng(8), 40):
        a, b = port.triangle_data(v, i), ref.triangle_data(v, i)
        assert ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all(), (len(v), i.size)


def test_device_acosf_equals_host_libm_on_every_float(tmp_path):
    """acosfLibm (tri_data_build.cuh: what the device uses for the corner angles of TriangleData) against the host libm's
    acosf — the reference's std::acos(float) — on all 2 130 706 434 floats of [-1, 1] (tests/cpp/acosf_libm_main.cpp)."""
    import os
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "acosf_libm_main")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-fopenmp", "-ffp-contract=off", "-fno-builtin", "-std=c++17", "-w", "-I/usr/local/cuda/include",
           "-I" + os.path.join(ROOT, "sdflib_b200", "csrc"), "-x", "c++", os.path.join(ROOT, "tests", "cpp", "acosf_libm_main.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok 2130706434"), r.stdout[-500:]


def test_octree_file_deeper_than_the_path_bits_is_rejected(sdf, tmp_path):
    """A .bin whose tree goes more than 16 levels below its start grid (the query kernels take child choices from 16 path
    bits) must be refused at load time with a file error — measured by walking the tree, not read from the header."""
    import struct
    from sdflib_b200 import _capi
    import ctypes as C

    def make(levels):
        g3, leaf = 1, 0x80000000
        coeff_at = g3 + 8 * levels                      # one shared 64-word coefficient block behind the child blocks
        words = [0] * (coeff_at + 64)
        words[0] = g3                                    # start slot -> first child block
        for k in range(levels):
            base = g3 + 8 * k
            for c in range(8):
                words[base + c] = leaf | coeff_at
            if k + 1 < levels:
                words[base] = g3 + 8 * (k + 1)           # child 0 keeps descending
        return np.asarray(words, np.uint32)

    def write(path, words, header_depth):
        with open(path, "wb") as f:
            f.write(struct.pack("<BI6fiIff", 1, 1, 0, 0, 0, 1, 1, 1, 1, header_depth, 1.0, 0.0))
            f.write(struct.pack("<Q", words.size))
            f.write(words.tobytes())

    L = sdf.lib()
    h = C.c_void_p()
    deep = str(tmp_path / "deep.bin")
    write(deep, make(17), 3)                             # the header lies about the depth
    assert L.sdfb200_load(deep.encode(), C.byref(h)) == _capi.ERR_IO
    assert b"16 levels" in L.sdfb200_last_error()
    ok = str(tmp_path / "ok.bin")
    write(ok, make(16), 16)
    code = L.sdfb200_load(ok.encode(), C.byref(h))       # accepted by the validator; without a GPU the upload then fails
    assert code in (_capi.OK, _capi.ERR_CUDA)
    if code == _capi.OK:
        L.sdfb200_free(h)

"""The reference's two command-line tools (src/tools/SdfExporter, src/tools/SdfError) rebuilt on the drop-in headers,
and the PLY / OBJ reader that replaces assimp in sdflib::Mesh(std::string).

CPU: the tools compile, parse the reference's flags, read meshes, and fail loudly without a GPU.
GPU: SdfExporter's .bin files are byte-identical to the ones the Python binding writes for the same arguments;
SdfError reports the tri-cubic octree's error against the exact octree below the build threshold's order."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, displaced_sphere

BIN = os.path.join(ROOT, "tools", "bin")


@pytest.fixture(scope="module")
def tools(sdf):
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return BIN


def write_ply(path, v, i, binary):
    with open(path, "wb") as f:
        f.write(("ply\nformat %s 1.0\ncomment test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face %d\nproperty list uchar int vertex_indices\nend_header\n"
                 % ("binary_little_endian" if binary else "ascii", len(v), len(i) // 3)).encode())
        if binary:
            f.write(np.ascontiguousarray(v, np.float32).tobytes())
            for t in range(len(i) // 3):
                f.write(struct.pack("<B3i", 3, *[int(x) for x in i[3 * t:3 * t + 3]]))
        else:
            for p in v:
                f.write(("%.9g %.9g %.9g\n" % tuple(p)).encode())
            for t in range(len(i) // 3):
                f.write(("3 %d %d %d\n" % tuple(i[3 * t:3 * t + 3])).encode())


def write_obj(path, v, i):
    with open(path, "w") as f:
        for p in v:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t in range(len(i) // 3):
            a, b, c = (int(x) + 1 for x in i[3 * t:3 * t + 3])
            f.write("f %d//%d %d//%d %d//%d\n" % (a, a, b, b, c, c))


def test_tools_compile_and_check_their_arguments(tools, tmp_path):
    r = subprocess.run([os.path.join(tools, "SdfExporter")], capture_output=True, text=True)
    assert r.returncode == 1 and "No model_path specified" in r.stderr
    r = subprocess.run([os.path.join(tools, "SdfExporter"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--termination_threshold" in r.stderr and "--min_triangles_per_node" in r.stderr
    r = subprocess.run([os.path.join(tools, "SdfExporter"), "model.ply", "out.bin", "--no_such_flag", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "could not be matched" in r.stderr
    r = subprocess.run([os.path.join(tools, "SdfExporter"), str(tmp_path / "missing.ply"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
    r = subprocess.run([os.path.join(tools, "SdfError"), str(tmp_path / "a.bin"), str(tmp_path / "b.bin")], capture_output=True, text=True)
    assert r.returncode == 1


def test_exporter_reads_meshes_and_has_no_cpu_fallback(sdf, tools, tmp_path):
    if sdf.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    v, i = displaced_sphere(1)
    for name, writer in (("a.ply", lambda p: write_ply(p, v, i, False)), ("b.ply", lambda p: write_ply(p, v, i, True)), ("c.obj", lambda p: write_obj(p, v, i))):
        path = str(tmp_path / name)
        writer(path)
        r = subprocess.run([os.path.join(tools, "SdfExporter"), path, str(tmp_path / "o.bin"), "-d", "3"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr, r.stderr   # the mesh was read; the build needs the GPU


@pytest.mark.gpu
def test_exporter_and_error_tool_on_gpu(sdf, tools, tmp_path):
    v, i = displaced_sphere(3)
    models = {"ascii.ply": lambda p: write_ply(p, v, i, False), "binary.ply": lambda p: write_ply(p, v, i, True), "mesh.obj": lambda p: write_obj(p, v, i)}
    for name, writer in models.items():
        writer(str(tmp_path / name))
    box = sdf.meshes.bounding_box_with_margin(v)   # bbox + 20 % of the largest extent = the tool's default margin
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    # octree, the tool's defaults: depth 8 -> use 6 here, start depth 1, continuity, trapezoidal 1e-3
    want = tmp_path / "want_octree.bin"
    sdf.OctreeSdf(mesh, bb, 6, 1, 1e-3, sdf.OctreeSdf.CONTINUITY, 1).saveToFile(want)
    for name in models:
        out = tmp_path / (name + ".octree.bin")
        r = subprocess.run([os.path.join(tools, "SdfExporter"), str(tmp_path / name), str(out), "-d", "6"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Computation time" in r.stdout
        assert open(out, "rb").read() == open(want, "rb").read(), name
    # every flag spelled out: no_continuity, by-distance rule, start depth 2, two "threads" (per-voxel layout)
    out = tmp_path / "flags.bin"
    r = subprocess.run([os.path.join(tools, "SdfExporter"), str(tmp_path / "binary.ply"), str(out), "--depth=5", "--start_depth", "2",
                        "--algorithm", "no_continuity", "--termination_rule", "by_distance_rule", "--termination_threshold", "2e-3",
                        "--termination_threshold_by_distance", "0.05", "--num_threads", "2", "--bb_margin", "20"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want2 = tmp_path / "want_flags.bin"
    sdf.OctreeSdf(mesh, bb, 5, 2, 2e-3, sdf.OctreeSdf.NO_CONTINUITY, 2, terminationRule=sdf.OctreeSdf.BY_DISTANCE_RULE,
                  terminationRuleParams=[2e-3, 0.05]).saveToFile(want2)
    assert open(out, "rb").read() == open(want2, "rb").read()
    # exact octree with the tool's defaults (depth 5, start depth 1, 32 triangles per node)
    exact = tmp_path / "exact.bin"
    r = subprocess.run([os.path.join(tools, "SdfExporter"), str(tmp_path / "mesh.obj"), str(exact), "--sdf_format", "exact_octree"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want3 = tmp_path / "want_exact.bin"
    sdf.ExactOctreeSdf(mesh, bb, 5, 1, 32, 1).saveToFile(want3)
    assert open(exact, "rb").read() == open(want3, "rb").read()
    # SdfError: tri-cubic octree against the exact field, one million rand() samples
    r = subprocess.run([os.path.join(tools, "SdfError"), str(tmp_path / "ascii.ply.octree.bin"), str(exact), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rmse = float(re.search(r"RMSE: ([0-9.eE+-]+)", r.stdout).group(1))
    mae = float(re.search(r"MAE: ([0-9.eE+-]+)", r.stdout).group(1))
    mx = float(re.search(r"Max error: ([0-9.eE+-]+)", r.stdout).group(1))
    assert "Sdf us per query" in r.stdout and "Exact Sdf us per query" in r.stdout
    assert 0 < mae <= rmse < 2e-3 and mx < 5e-2, r.stdout
    # normalisation path: model scaled into [-1, 1] before the margin is added
    out = tmp_path / "normalized.bin"
    r = subprocess.run([os.path.join(tools, "SdfExporter"), str(tmp_path / "binary.ply"), str(out), "-d", "4", "-n", "--algorithm", "no_continuity"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    loaded = sdf.SdfFunction.loadFromFile(out)
    area = loaded.getSampleArea().as_array()
    assert abs((area[3:] - area[:3]).max() - 2.8) < 1e-4   # 2 (normalised extent) + 2 x 20 % margin


# ---- sdflib::Mesh(path): the PLY / OBJ reader on its own (host only) ---------------------------------------------
@pytest.fixture(scope="module")
def mesh_reader(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("reader") / "mesh_reader")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
                        os.path.join(ROOT, "tests", "cpp", "mesh_reader_main.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def read_back(exe, path):
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    nv, ni = (int(x) for x in lines[0].split())
    v = np.array([[float(x) for x in l.split()] for l in lines[1:1 + nv]], np.float32).reshape(nv, 3)
    i = np.array([int(l) for l in lines[1 + nv:1 + nv + ni]], np.uint32)
    box = np.array([float(x) for x in lines[1 + nv + ni].split()], np.float32)
    return v, i, box


def test_mesh_reader_round_trips_ply_and_obj(mesh_reader, tmp_path):
    v, i = displaced_sphere(2)
    for name, writer in (("a.ply", lambda p: write_ply(p, v, i, False)), ("b.ply", lambda p: write_ply(p, v, i, True)),
                         ("c.obj", lambda p: write_obj(p, v, i))):
        path = str(tmp_path / name)
        writer(path)
        rv, ri, box = read_back(mesh_reader, path)
        assert np.array_equal(rv.view(np.uint32), v.view(np.uint32)), name      # %.9g round-trips float32 exactly
        assert np.array_equal(ri, i), name
        assert np.array_equal(box, np.concatenate([v.min(0), v.max(0)])), name


def test_mesh_reader_polygons_extra_properties_and_relative_indices(mesh_reader, tmp_path):
    quad = np.float32([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]])
    # binary PLY: double coordinates, extra vertex properties (normals, colour), a face property before the index list,
    # a quad and a triangle -> fan triangulation 0 1 2, 0 2 3, 0 1 4
    path = str(tmp_path / "rich.ply")
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 5\nproperty double x\nproperty double y\nproperty double z\n"
                b"property float nx\nproperty uchar red\nelement face 2\nproperty uchar flags\nproperty list uchar uint vertex_index\n"
                b"element edge 1\nproperty int vertex1\nproperty int vertex2\nend_header\n")
        for p in quad:
            f.write(struct.pack("<3dfB", float(p[0]), float(p[1]), float(p[2]), 0.5, 7))
        f.write(struct.pack("<BB4I", 1, 4, 0, 1, 2, 3))
        f.write(struct.pack("<BB3I", 0, 3, 0, 1, 4))
        f.write(struct.pack("<2i", 0, 1))
    rv, ri, _ = read_back(mesh_reader, path)
    assert np.array_equal(rv, quad) and ri.tolist() == [0, 1, 2, 0, 2, 3, 0, 1, 4]
    # OBJ: comments, texture / normal references, a quad, negative (relative) indices
    path = str(tmp_path / "rich.obj")
    with open(path, "w") as f:
        f.write("# comment\nvn 0 0 1\nvt 0 0\n")
        for p in quad:
            f.write("v %g %g %g\n" % tuple(p))
        f.write("f 1/1/1 2/1/1 3/1/1 4/1/1\nf -5//1 -4//1 -1//1\n")
    rv, ri, _ = read_back(mesh_reader, path)
    assert np.array_equal(rv, quad) and ri.tolist() == [0, 1, 2, 0, 2, 3, 0, 1, 4]


def test_mesh_reader_rejects_bad_files(mesh_reader, tmp_path):
    cases = {"empty.obj": "", "bad.ply": "not a ply\n", "big.ply": "ply\nformat binary_big_endian 1.0\nend_header\n",
             "range.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n", "model.stl": "solid\n",
             # damaged headers / lists found by fuzzing: must end in an exception, not in a loop over 2^32 entries
             "wrap.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4294967297\n",
             "shifted.ply": "ply\nformat ascii 1.0\nelement vertex 3\nproperXy float x\nproperty float y\nproperty float z\n"
                            "element face 1\nproperty list uchar int vertex_indices\nend_header\n-0.5 0 0.8\n0.5 0 0.8\n0 1 -1e30\n3 0 1 2\n",
             "neglist.ply": "ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                            "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n-7 0 1 2\n",
             "negindex.ply": "ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                             "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 -1 2\n"}
    for name, text in cases.items():
        path = str(tmp_path / name)
        open(path, "w").write(text)
        r = subprocess.run([mesh_reader, path], capture_output=True, text=True)
        assert r.returncode == 1 and "Mesh:" in r.stderr, (name, r.stderr)
    r = subprocess.run([mesh_reader, str(tmp_path / "missing.ply")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_mesh_reader_skips_property_less_elements(mesh_reader, tmp_path):
    path = str(tmp_path / "spin.ply")
    open(path, "w").write("ply\nformat ascii 1.0\nelement foo 99999999999999\nelement vertex 3\nproperty float x\nproperty float y\n"
                          "property float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    r = subprocess.run([mesh_reader, path], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and r.stdout.startswith("3 3"), r.stderr

"""Host models of the building blocks of octreeQueryTileKernel (octree_query_kernels.cuh): the dense top index
(topIndexKernel) and the quad-cooperative evaluation, plus the kernel SOURCE run on the CPU under a warp emulation.

The CUDA code cannot run here; what can be checked on the CPU is the arithmetic it transcribes: the entry packing
((block - G^3) / 8, steps, leaf flag), the cell chosen from the start cell and the leading path bits, and the finish of
the descent for trees deeper than the index — against the plain descent of octreeQueryKernel (itself bit-exact with
the oracle on the GPU, tests/test_gpu_octree.py), on octrees built by the oracle."""
import numpy as np
import pytest

from conftest import assert_bit_equal

LEAF, MASK = np.uint32(1 << 31), np.uint32(~(3 << 30) & 0xFFFFFFFF)
PATH_BITS = 16


def plain_descent(oct_, G, ix, iy, iz, bx, by, bz):
    node = oct_[(iz * G + iy) * G + ix]
    k = 0
    while not node & LEAF:
        sh = PATH_BITS - 1 - k
        child = ((bx >> sh) & 1) | (((by >> sh) & 1) << 1) | (((bz >> sh) & 1) << 2)
        node = oct_[int(node & MASK) + child]
        k += 1
    return int(node & MASK), k


def build_index(oct_, G, levels):
    N, G3 = G << levels, G ** 3
    index = np.zeros(N ** 3, np.uint32)
    for i in range(N ** 3):
        cx, cy, cz = i % N, (i // N) % N, i // (N * N)
        node = oct_[((cz >> levels) * G + (cy >> levels)) * G + (cx >> levels)]
        k = 0
        while not node & LEAF and k < levels:
            sh = levels - 1 - k
            child = ((cx >> sh) & 1) | (((cy >> sh) & 1) << 1) | (((cz >> sh) & 1) << 2)
            node = oct_[int(node & MASK) + child]
            k += 1
        block = int(node & MASK)
        assert block >= G3 and (block - G3) % 8 == 0      # what the kernel reports through `bad`
        index[i] = ((block - G3) >> 3) | (k << 27) | ((1 << 31) if node & LEAF else 0)
    return index


def indexed_descent(oct_, index, G, levels, ix, iy, iz, bx, by, bz):
    N, G3 = G << levels, G ** 3
    cx = (ix << levels) | (bx >> (PATH_BITS - levels))
    cy = (iy << levels) | (by >> (PATH_BITS - levels))
    cz = (iz << levels) | (bz >> (PATH_BITS - levels))
    e = int(index[(cz * N + cy) * N + cx])
    block = ((e & ((1 << 27) - 1)) << 3) + G3
    k = (e >> 27) & 15
    if e & (1 << 31):
        return block, k
    assert k == levels
    while True:
        sh = PATH_BITS - 1 - k
        child = ((bx >> sh) & 1) | (((by >> sh) & 1) << 1) | (((bz >> sh) & 1) << 2)
        node = oct_[block + child]
        k += 1
        block = int(node & MASK)
        if node & LEAF:
            return block, k


@pytest.mark.parametrize("depth,start,levels", [(5, 2, 3), (5, 2, 1), (5, 2, 0), (6, 2, 2), (5, 3, 2)])
def test_indexed_descent_reaches_the_same_leaf(port, depth, start, levels):
    from sdflib_b200 import meshes
    v, i = meshes.isosphere(2)
    v = (v * np.float32([1.0, 0.8, 0.6]) + np.float32([0.013, -0.007, 0.003])).astype(np.float32)
    box = np.float32([-1.3, -1.3, -1.3, 1.3, 1.3, 1.3])
    sdf = port.build_octree(v, i, box, depth, start, threshold=3e-3, algorithm=1, use_cache=False)
    oct_ = sdf.octree_data()
    G = sdf.header()["start_grid_size"]
    assert G == 1 << start and levels <= depth - start
    index = build_index(oct_, G, levels)
    rng = np.random.default_rng(7)
    cells = rng.integers(0, G, size=(4000, 3))
    bits = rng.integers(0, 1 << PATH_BITS, size=(4000, 3))
    bits[:64] = 0                                  # cell corners
    bits[64:128] = (1 << PATH_BITS) - 1            # just below the next cell
    bits[128:192] = 1 << (PATH_BITS - 1)           # exactly on the first split plane (child = frac >= 0.5)
    depths = set()
    for (ix, iy, iz), (bx, by, bz) in zip(cells.tolist(), bits.tolist()):
        want = plain_descent(oct_, G, ix, iy, iz, bx, by, bz)
        got = indexed_descent(oct_, index, G, levels, ix, iy, iz, bx, by, bz)
        assert got == want
        depths.add(want[1])
    assert len(depths) >= 2                        # leaves above, at and below the index depth were all visited


def test_quad_cooperative_evaluation_is_the_same_polynomial(port):
    """octreeQueryCoopKernel: lane r of a quad takes the vectors m = r + 4 k (coefficients c[0..3][j = r][k]), forms
    y^r * Horner_z, the quad adds its four parts with a butterfly, members finish with the cubic in x (and the three
    derivative forms). Same float32 operations here (without fusing); must agree with the reference's evaluation order
    (oracle) within the FMA kernel's tolerance, and all four lanes must hold identical sums."""
    f32 = np.float32
    rng = np.random.default_rng(3)
    for trial in range(200):
        c = (rng.standard_normal(64) * rng.choice([1e-3, 1.0, 30.0])).astype(f32)
        x, y, z = (f32(t) for t in rng.random(3))
        if trial < 8:
            x, y, z = f32(trial & 1), f32((trial >> 1) & 1), f32((trial >> 2) & 1)      # corners, incl. frac = 0
        vec = c.reshape(16, 4)                                                        # vector m = j + 4 k, components i
        parts = np.zeros((4, 12), f32)
        for r in range(4):
            c0, c1, c2, c3 = vec[r], vec[r + 4], vec[r + 8], vec[r + 12]
            yr = [f32(1), y, f32(y * y), f32(f32(y * y) * y)][r]
            dyr = [f32(0), f32(1), f32(f32(2) * y), f32(f32(3) * f32(y * y))][r]
            h = ((c3 * z + c2) * z + c1) * z + c0
            dh = (f32(3) * c3 * z + f32(2) * c2) * z + c1
            parts[r, 0:4], parts[r, 4:8], parts[r, 8:12] = yr * h, dyr * h, yr * dh
        lanes = parts.copy()
        for m in (1, 2):                                                              # butterfly: lane l += lane l ^ m
            lanes = (lanes + lanes[[l ^ m for l in range(4)]]).astype(f32)
        assert (lanes.view(np.uint32) == lanes[0].view(np.uint32)).all()              # identical bits on all four lanes
        a, ay, az = lanes[0, 0:4], lanes[0, 4:8], lanes[0, 8:12]
        value = ((a[3] * x + a[2]) * x + a[1]) * x + a[0]
        gx = (f32(3) * a[3] * x + f32(2) * a[2]) * x + a[1]
        gy = ((ay[3] * x + ay[2]) * x + ay[1]) * x + ay[0]
        gz = ((az[3] * x + az[2]) * x + az[1]) * x + az[0]
        g = np.array([gx, gy, gz], np.float64)
        want_v, want_g, _ = port.tricubic_eval(c, np.array([[x, y, z]], f32))
        scale = float(np.abs(c).sum())
        assert abs(float(value) - float(want_v[0])) <= 2e-6 * scale
        np.testing.assert_allclose(g, want_g[0].astype(np.float64), atol=6e-6 * scale)   # raw (un-normalised) gradient


def _compile_simt(tmp_path, exact=False):
    import os
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / ("simt_query_exact" if exact else "simt_query_main"))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cuda_inc = "/usr/local/cuda/include"
    cmd = [cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-w", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "sdflib_b200", "csrc"), "-x", "c++", os.path.join(ROOT, "tests", "cpp", "simt_query_main.cpp"),
           "-o", exe, "-lpthread"] + (["-DSDFB_QUERY_EXACT"] if exact else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _query_case(port, tmp_path):
    from sdflib_b200 import meshes
    v, i = meshes.isosphere(2)
    v = (v * np.float32([1.0, 0.8, 0.6]) + np.float32([0.013, -0.007, 0.003])).astype(np.float32)
    box = np.float32([-1.3, -1.3, -1.3, 1.3, 1.3, 1.3])
    depth, start = 5, 2
    sdf = port.build_octree(v, i, box, depth, start, threshold=3e-3, algorithm=1, use_cache=False)
    oct_, hdr, area = sdf.octree_data(), sdf.header(), sdf.sample_area()
    size = float(area[3] - area[0])
    rng = np.random.default_rng(11)
    grid_pts = meshes.cell_centre_grid(area, 32)[: 32 * 32 * 2]                       # two z-slabs of rows: shared (leaf, y, z)
    rand_pts = (area[:3] + rng.random((1500, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)
    out_pts = (area[:3] - 0.3 + rng.random((400, 3), np.float32) * (area[3:] - area[:3] + 0.6)).astype(np.float32)
    pts = np.concatenate([grid_pts, rand_pts, out_pts])[:-7].astype(np.float32)        # n % 32 != 0
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(area.astype(np.float32).tobytes())
        f.write(np.float32(size / hdr["start_grid_size"]).tobytes())
        f.write(np.int32(hdr["start_grid_size"]).tobytes())
        f.write(np.float32(hdr["min_border_value"]).tobytes())
        f.write(np.uint32(depth).tobytes())
        f.write(np.uint64(len(oct_)).tobytes()); f.write(oct_.tobytes())
        f.write(np.uint64(pts.size).tobytes()); f.write(pts.tobytes())
    return sdf, pts, area, size


def test_reference_order_kernels_under_warp_emulation_equal_the_oracle(port, tmp_path):
    """Pins the emulation: the -DSDFB_QUERY_EXACT build of the same header (the kernels behind
    SDFB200_QUERY_EXACT_ORDER, bit-exact on the GPU) must give the oracle's bits on the CPU too — in-box distances and
    gradients."""
    import subprocess
    exe = _compile_simt(tmp_path, exact=True)
    sdf, pts, area, size = _query_case(port, tmp_path)
    n = len(pts)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.stdout, r.stderr)
    raw = np.fromfile(tmp_path / "out.bin", np.float32)
    assert raw.size == 5 * n
    ref_d, ref_g = sdf.query(pts, gradient=True)
    inside = ~((pts < area[:3]) | (pts >= area[3:])).any(1)
    for k in range(1):
        d, dg, g = raw[5 * n * k:5 * n * k + n], raw[5 * n * k + n:5 * n * k + 2 * n], raw[5 * n * k + 2 * n:5 * n * (k + 1)].reshape(n, 3)
        assert_bit_equal(d, sdf.query(pts), "distance")
        assert_bit_equal(dg[inside], ref_d[inside], "distance of the gradient kernel")
        assert_bit_equal(g[inside], ref_g[inside], "gradient")


@pytest.mark.parametrize("index_levels", [None, 1])
def test_kernel_sources_under_warp_emulation(port, tmp_path, index_levels):
    """The CUDA source of the query kernels (octree_query_kernels.cuh), compiled for the host and run under a lock-step
    warp emulation (tests/cpp/simt_query_main.cpp): the tile kernel (top index, division-free cell selection —
    compared with IEEE division inside the harness —, quad-cooperative evaluation, persistent warps looping over tiles)
    must stay within the FMA kernel's tolerance of the reference's evaluation order (oracle), including points outside
    the box, rows of a grid (shared classes), unrelated points (one class per lane) and a batch that ends in the middle
    of a warp."""
    import subprocess
    exe = _compile_simt(tmp_path)
    sdf, pts, area, size = _query_case(port, tmp_path)
    n = len(pts)
    cmd = [exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")] + ([str(index_levels)] if index_levels is not None else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.stdout, r.stderr)
    raw = np.fromfile(tmp_path / "out.bin", np.float32)
    assert raw.size == 2 * (n + 4 * n)
    parts = {}
    off = 0
    for name in ("plain", "tile"):
        parts[name] = dict(d=raw[off:off + n], dg=raw[off + n:off + 2 * n], g=raw[off + 2 * n:off + 5 * n].reshape(n, 3))
        off += 5 * n
    ref_d, ref_g = sdf.query(pts, gradient=True)
    # the plain FMA kernel itself sits at 1.05 x the GPU tests' floor (1e-3 of the box) on one near-surface point of
    # this tree (|d| = 0.0027, coefficients of order 1), so the floor here is 1e-2 of the box for both kernels
    tol = 1e-5 * np.maximum(np.abs(ref_d), 1e-2 * size)
    for name in ("plain", "tile"):
        assert (np.abs(parts[name]["d"] - ref_d) <= tol).all(), name
        assert (np.abs(parts[name]["dg"] - ref_d) <= tol).all(), name
        fin = np.isfinite(ref_g).all(1) & np.isfinite(parts[name]["g"]).all(1)
        assert fin.mean() > 0.9 and np.abs(parts[name]["g"][fin] - ref_g[fin]).max() <= 1e-4, name
    outside = ((pts < area[:3]) | (pts >= area[3:])).any(1)
    assert outside.sum() > 50      # box-distance path: identical code in both kernels
    assert (parts["plain"]["d"][outside].view(np.uint32) == parts["tile"]["d"][outside].view(np.uint32)).all()


def test_refill_sampler_source_under_warp_emulation(sdf, tmp_path):
    """The lane-refill BVH sampler (bvh_sampler.cuh) run from its CUDA source under the warp emulation
    (tests/cpp/simt_sampler_main.cpp): it must terminate and reproduce the plain sampler's bits sample for sample."""
    import os
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "simt_sampler_main")
    lib_dir = os.path.join(ROOT, "sdflib_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-w", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(lib_dir, "csrc"), "-x", "c++", os.path.join(ROOT, "tests", "cpp", "simt_sampler_main.cpp"), "-o", exe,
           "-L" + lib_dir, "-lsdfb200", "-Wl,-rpath," + lib_dir, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # third argument: the leaf-batch threshold of the refill kernel (1 = the default schedule; larger values were measured slower
    # on the GPU but must stay correct: they are reachable through SDFB200_LEAF_BATCH)
    # fourth argument: 1 = the samples are started in a permuted order (the far-first schedule of LevelSampler::run)
    for subdivisions, nodes, leaf_batch, scheduled in ((4, 120, 1, 0), (4, 120, 1, 1), (0, 3, 1, 1), (2, 1, 1, 1), (3, 40, 8, 0), (3, 40, 32, 1)):
        r = subprocess.run([exe, str(subdivisions), str(nodes), str(leaf_batch), str(scheduled)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "identical" in r.stdout, (subdivisions, nodes, leaf_batch, scheduled, r.stdout, r.stderr)

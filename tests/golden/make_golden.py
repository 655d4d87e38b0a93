"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsdfref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures are small on purpose; they pin the CPU restatement (oracle/oracle.cpp) wherever the
reference itself is not present, and give the GPU tests reference outputs that do not depend on any
oracle library being loadable.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def displace(v):   # same closed form as sdflib_b200.meshes.displace; the displaced vertices are STORED in the fixture
    v = np.asarray(v, np.float32)
    d = np.zeros(len(v), np.float32)
    a, f = np.float32(0.15), np.float32(3.0)
    for o in range(5):
        d += a * np.sin(f * v[:, 0] + np.float32(1.3 * o), dtype=np.float32) * np.sin(f * v[:, 1] + np.float32(2.1 * o), dtype=np.float32) \
               * np.sin(f * v[:, 2] + np.float32(0.7 * o), dtype=np.float32)
        a *= np.float32(0.5); f *= np.float32(2.0)
    return (v * (np.float32(1.0) + d)[:, None] + np.float32([0.013, -0.007, 0.003])).astype(np.float32)


def box_of(v):
    mn, mx = v.min(0), v.max(0)
    m = np.float32(0.2) * (mx - mn).max()
    return np.concatenate([mn - m, mx + m]).astype(np.float32)


def main():
    rng = np.random.default_rng(2222)
    # ---- kernel known-answer vectors ------------------------------------------------------------
    # the triangle of the reference's TriangleDistanceTest (src/tools/TriangleDistanceTest/main.cpp:12-64)
    tv = np.float32([[-0.5, -0.5, 0.0], [0.5, -0.5, 0.0], [0.0, 0.5, 0.0], [0.1, 0.0, 0.7]])
    ti = np.uint32([0, 1, 2, 1, 0, 3, 2, 1, 3, 0, 2, 3])   # closed tetrahedron around that triangle
    td = ref.triangle_data(tv, ti)
    pts = rng.uniform(-1, 1, (512, 3)).astype(np.float32)
    w = tv[ti[:3]].reshape(-1)
    sq = ref.sq_dist(td[0], pts)
    s0, _ = ref.signed_dist(td[0], w, pts, 0)
    s1, g1 = ref.signed_dist(td[0], w, pts, 1)
    s2, g2 = ref.signed_dist(td[0], w, pts, 2)
    vals = rng.standard_normal((8, 8)).astype(np.float32)
    vals[:, 4:] = 0
    coeff = ref.tricubic_coefficients(vals, 0.37)
    frac = rng.uniform(0, 1, (64, 3)).astype(np.float32)
    ev, eg, evv = ref.tricubic_eval(coeff, frac, 0.37)
    mid = rng.standard_normal((19, 8)).astype(np.float32)
    err = np.float32([ref.error_estimate(coeff, mid, 1), ref.error_estimate(coeff, mid, 3, 0.1)])
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), tet_vertices=tv, tet_indices=ti, tet_triangle_data=td, points=pts,
                        sq_dist=sq, signed0=s0, signed1=s1, grad1=g1, signed2=s2, grad2=g2, corner_values=vals,
                        node_size=np.float32(0.37), coefficients=coeff, frac=frac, value=ev, gradient=eg, vertex_values=evv,
                        mid_values=mid, error=err)

    # ---- nearest triangle + Frank-Wolfe filter on a small displaced sphere -----------------------
    v, i = ref.isosphere(2)
    vd = displace(v)
    qp = rng.uniform(-1.6, 1.6, (256, 3)).astype(np.float32)
    near = ref.nearest_triangle(vd, i, qp)
    centre = np.float32([0.21, -0.33, 0.4])
    half = np.float32(0.35)
    corners = centre + half * np.float32([[(c & 1) * 2 - 1, ((c >> 1) & 1) * 2 - 1, (c >> 2) * 2 - 1] for c in range(8)])
    corner_tris = ref.nearest_triangle(vd, i, corners.astype(np.float32))
    kept = ref.filter_triangles(vd, i, centre, half, np.arange(i.size // 3, dtype=np.uint32), corner_tris)
    np.savez_compressed(os.path.join(OUT, "mesh_small.npz"), sphere_vertices=v, vertices=vd, indices=i, triangle_data=ref.triangle_data(vd, i),
                        query_points=qp, nearest=near, filter_centre=centre, filter_half=half, filter_corner_tris=corner_tris,
                        filter_kept=kept)

    # ---- config 1 (icosphere 320 tris, OctreeSdf d5 s3 thr 1e-3, single thread) ------------------
    box = box_of(v)
    a = ref.build_octree(v, i, box, 5, 3, 1e-3, 1, 1)
    d = a.octree_data()
    area = a.sample_area()
    q = (area[:3] + rng.uniform(-0.05, 1.05, (512, 3)) * (area[3:] - area[:3])).astype(np.float32)
    dist, grad = a.query(q, True)
    h = a.header()
    np.savez_compressed(os.path.join(OUT, "config1_octree.npz"), box=box, words=np.int64(d.size), sha256=sha(d), start_slots=d[:512],
                        value_range=np.float32(h["value_range"]), min_border_value=np.float32(h["min_border_value"]),
                        query_points=q, distances=dist, gradients=grad)

    # ---- generic-position mesh: OctreeSdf + ExactOctreeSdf, complete .bin files -------------------
    boxd = box_of(vd)
    o = ref.build_octree(vd, i, boxd, 4, 2, 1e-3, 1, 1)
    e = ref.build_exact(vd, i, boxd, 4, 1, 16, 1)
    o.save("/tmp/golden_octree.bin"); e.save("/tmp/golden_exact.bin")
    ob = np.frombuffer(open("/tmp/golden_octree.bin", "rb").read(), np.uint8)
    eb = np.frombuffer(open("/tmp/golden_exact.bin", "rb").read(), np.uint8)
    ar = o.sample_area()
    q2 = (ar[:3] + rng.uniform(-0.05, 1.05, (512, 3)) * (ar[3:] - ar[:3])).astype(np.float32)
    od, og = o.query(q2, True)
    ed, eg2 = e.query(q2, True)
    np.savez_compressed(os.path.join(OUT, "small_structures.npz"), box=boxd, octree_bin=ob, exact_bin=eb, query_points=q2,
                        octree_distances=od, octree_gradients=og, exact_distances=ed, exact_gradients=eg2)
    continuity()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


def continuity():
    """InitAlgorithm::CONTINUITY (src/sdf/OctreeSdfBreadthFirstNoDelay.h) + the Simpson / by-distance rules.
    Own generator so that `make_golden.py continuity` adds this fixture without touching the others."""
    rng = np.random.default_rng(3333)
    v, i = ref.isosphere(2)
    vd = displace(v)
    box = box_of(vd)
    out = {"box": box}
    for name, rule, p1 in (("trapezoid", 1, 0.0), ("simpson", 2, 0.0), ("by_distance", 3, 0.1)):
        a = ref.build_octree(vd, i, box, 5, 3, 1e-3, 2, 1, termination_rule=rule, param1=p1)
        d = a.octree_data()
        out[name + "_words"] = np.int64(d.size)
        out[name + "_sha256"] = sha(d)
        out[name + "_min_border_value"] = np.float32(a.header()["min_border_value"])
        if rule == 1:
            area = a.sample_area()
            q = (area[:3] + rng.uniform(-0.05, 1.05, (512, 3)) * (area[3:] - area[:3])).astype(np.float32)
            dist, grad = a.query(q, True)
            out.update(query_points=q, distances=dist, gradients=grad, start_slots=d[:512])
    vals = rng.standard_normal((8, 8)).astype(np.float32)
    coeff = ref.tricubic_coefficients(vals, 0.41)     # all eight Hermite slots in use (mixed derivatives non-zero)
    mid = rng.standard_normal((19, 8)).astype(np.float32)
    out.update(corner_values=vals, node_size=np.float32(0.41), coefficients=coeff, mid_values=mid,
               error=np.float32([ref.error_estimate(coeff, mid, r, 0.1) for r in (1, 2, 3)]))
    np.savez_compressed(os.path.join(OUT, "continuity_small.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["continuity"]:
        continuity()
    else:
        main()

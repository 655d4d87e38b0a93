"""Generates tests/golden/full_size.npz: parity pins of BASELINE configs 2 and 3 AT THEIR STATED SIZES.

Run in the build container (needs oracle/_ref/libsdfref.so = the unmodified reference and oracle/liboracle.so):
    python tests/golden/make_golden_full.py            # ~10 minutes on 8 cores
Per structure it stores
  * the word count and sha256 of the arrays of the HISTORY-FREE oracle (port, use_cache=0, single thread) — the
    GPU builds must reproduce them bit for bit (tests/test_gpu_full_size.py);
  * the sha256 of the topology words of the REFERENCE's own single-thread build (node words of OctreeSdf; node /
    set / mask arrays of ExactOctreeSdf) — the history cache of the reference only moves tie-broken leaf values;
  * the reference's getDistance (+ gradient) on a strided sample of the 256^3 cell-centre grid (stride 61, coprime
    with the row length so every x column is hit; gradients on every 8th of those points).
Mesh M1 = PrimitivesFactory::getIsosphere(7) + the closed-form displacement of make_golden.py (327 680 triangles).
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import ref, port          # noqa: E402
from make_golden import displace, box_of, sha  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STRIDE, GRAD_EVERY = 61, 8


def topology_mask(words, g3):
    """Vectorised walk of an OctreeSdf array: True where the word is a node word (start slots + child blocks)."""
    words = np.asarray(words)
    topo = np.zeros(words.size, bool)
    topo[:g3] = True
    level = np.arange(g3, dtype=np.int64)
    while level.size:
        w = words[level]
        inner = (w & 0x80000000) == 0
        base = (w[inner] & 0x3FFFFFFF).astype(np.int64)
        level = (base[:, None] + np.arange(8)).reshape(-1)
        topo[level] = True
    return topo


def grid_sample(area, n=256):
    g = (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)
    idx = np.arange(0, n ** 3, STRIDE, dtype=np.int64)
    p = np.stack([g[idx % n], g[(idx // n) % n], g[idx // (n * n)]], -1)
    return (area[:3] + p * (area[3:] - area[:3])).astype(np.float32)


def file_sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def exact_arrays(p):
    import ctypes as C
    ns, nm, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    p.b.fn("exact_sizes")(p.h, C.byref(ns), C.byref(nm), C.byref(nt))
    sets, masks, tris = np.empty(ns.value, np.uint32), np.empty(nm.value, np.uint8), np.empty((nt.value, 37), np.float32)
    p.b.fn("exact_arrays")(p.h, sets.ctypes.data_as(C.c_void_p), masks.ctypes.data_as(C.c_void_p), tris.ctypes.data_as(C.c_void_p))
    return sets, masks, tris


def main():
    v, i = ref.isosphere(7)
    v = displace(v)
    box = box_of(v)
    out = {"box": box, "stride": np.int64(STRIDE), "grad_every": np.int64(GRAD_EVERY), "mesh_sha256": sha(v) + sha(i)}
    threads = os.cpu_count() or 1
    if os.path.exists("/tmp/full_size_partial.npz") and "--resume" in sys.argv:
        out.update({k: v for k, v in np.load("/tmp/full_size_partial.npz").items()})
    for name, algo in (("c2_nocont", 1), ("c2_cont", 2)):
        if name + "_distances" in out:
            continue
        t = time.time()
        p = port.build_octree(v, i, box, 8, 3, 1e-3, algo, 1, use_cache=False)
        d = p.octree_data()
        h = p.header()
        out.update({name + "_words": np.int64(d.size), name + "_sha256": sha(d), name + "_min_border_value": np.float32(h["min_border_value"]),
                    name + "_value_range": np.float32(h["value_range"])})
        print(name, "port", d.size, f"{time.time() - t:.1f}s", flush=True)
        p.save("/tmp/golden_full.bin")
        out[name + "_bin_sha256"] = file_sha("/tmp/golden_full.bin")
        area = p.sample_area()
        q = grid_sample(area)
        pd, pg = p.query(q, True, threads)
        out.update({name + "_port_distances_sha256": sha(pd), name + "_port_gradients_sha256": sha(pg[::GRAD_EVERY])})
        p.close()
        t = time.time()
        r = ref.build_octree(v, i, box, 8, 3, 1e-3, algo, 1)
        rd = r.octree_data()
        topo = topology_mask(rd, 512)
        out.update({name + "_ref_words": np.int64(rd.size), name + "_ref_topology_sha256": sha(rd[topo]), name + "_ref_c0_sha256": sha(rd[~topo].reshape(-1, 64)[:, 0])})
        # The reference's 32^3 vertex cache (TrianglesInfluence.h:934-991) makes tie-broken samples depend on traversal
        # history; where such a sample sits next to the error threshold the subdivision decision can flip. Recorded, not
        # asserted: NO_CONTINUITY keeps the reference's topology at this size, CONTINUITY does not.
        same_topology = d.size == rd.size and np.array_equal(topology_mask(d, 512), topo) and np.array_equal(d[topo], rd[topo])
        out[name + "_topology_equals_reference"] = np.bool_(same_topology)
        if not same_topology:
            tp = topology_mask(d, 512)
            out[name + "_leaves"] = np.int64((~tp).sum() // 64)
            out[name + "_ref_leaves"] = np.int64((~topo).sum() // 64)
            print(name, "topology differs from the reference's single-thread build: leaves", int((~tp).sum() // 64), "vs", int((~topo).sum() // 64), flush=True)
        dist, grad = r.query(q, True, threads)
        out.update({name + "_distances": dist, name + "_gradients": grad[::GRAD_EVERY].copy()})
        print(name, "ref", rd.size, f"{time.time() - t:.1f}s", "values differing from the history-free build:", int((rd != d).sum()) if rd.size == d.size else -1, flush=True)
        r.close()
        out["sample_area"] = area
        np.savez_compressed(os.path.join("/tmp", "full_size_partial.npz"), **out)
    # config 3: ExactOctreeSdf depth 7, start depth 3 (the SdfExporter default start depth of the exact path), minTri 128
    t = time.time()
    p = port.build_exact(v, i, box, 7, 3, 128, 1, use_cache=False)
    nodes = p.octree_data()
    sets, masks, tris = exact_arrays(p)
    out.update(c3_nodes=np.int64(nodes.size // 2), c3_nodes_sha256=sha(nodes), c3_sets_words=np.int64(sets.size), c3_sets_sha256=sha(sets),
               c3_masks_bytes=np.int64(masks.size), c3_masks_sha256=sha(masks), c3_triangle_data_sha256=sha(tris))
    print("c3 port", nodes.size // 2, sets.size, masks.size, f"{time.time() - t:.1f}s", flush=True)
    p.save("/tmp/golden_full.bin")
    out["c3_bin_sha256"] = file_sha("/tmp/golden_full.bin")
    q = grid_sample(p.sample_area())
    pd, pg = p.query(q, True, threads)
    p.close()
    t = time.time()
    r = ref.build_exact(v, i, box, 7, 3, 128, 1)
    rn = r.octree_data()
    r.save("/tmp/golden_full_ref.bin")
    out.update(c3_ref_nodes_sha256=sha(rn), c3_ref_bin_sha256=file_sha("/tmp/golden_full_ref.bin"))
    dist, grad = r.query(q, True, threads)
    # exact distances do not depend on layout or history: the history-free port must give the same bits
    assert np.array_equal(dist.view(np.uint32), pd.view(np.uint32)) and np.array_equal(grad.view(np.uint32), pg.view(np.uint32))
    out.update(c3_distances=dist, c3_gradients=grad[::GRAD_EVERY].copy())
    print("c3 ref", rn.size // 2, f"{time.time() - t:.1f}s", "nodes equal:", np.array_equal(rn, nodes), "bin equal:", out["c3_ref_bin_sha256"] == out["c3_bin_sha256"], flush=True)
    r.close()
    np.savez_compressed(os.path.join(OUT, "full_size.npz"), **out)
    print("full_size.npz", os.path.getsize(os.path.join(OUT, "full_size.npz")))


if __name__ == "__main__":
    main()

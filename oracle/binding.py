"""TEST INFRASTRUCTURE — ctypes loaders for the two CPU oracles.

  ref  : oracle/_ref/libsdfref.so — the UNMODIFIED reference compiled by oracle/Makefile (kind "reference")
  port : oracle/liboracle.so      — our CPU restatement, oracle/oracle.cpp (kind "port")

Both export the same entry points (prefix ``ref_`` / ``orc_``) so a test can run one call against both
and compare bit for bit. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module; the product package (sdflib_b200) never does.
"""
import ctypes as C
import os
import time
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Backend:
    def __init__(self, prefix, path, kind):
        self.prefix, self.path, self.kind = prefix, path, kind
        self._lib = None

    def available(self):
        return os.path.exists(self.path)

    @property
    def lib(self):
        if self._lib is None:
            L = C.CDLL(self.path)
            for name, res in (("build_octree", C.c_void_p), ("build_exact", C.c_void_p), ("load", C.c_void_p),
                              ("octree_data_size", C.c_uint64), ("query", C.c_double),
                              ("error_estimate", C.c_float), ("filter_triangles", C.c_uint32)):
                getattr(L, self.prefix + name).restype = res
            if self.kind == "port":
                L.orc_exact_sizes.restype = C.c_uint64
            self._lib = L
        return self._lib

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- fixtures / kernels ---------------------------------------------------------------
    def isosphere(self, subdiv):
        nv, ni = C.c_uint32(), C.c_uint32()
        self.fn("isosphere")(C.c_uint32(subdiv), None, None, C.byref(nv), C.byref(ni))
        v = np.empty((nv.value, 3), np.float32)
        i = np.empty(ni.value, np.uint32)
        self.fn("isosphere")(C.c_uint32(subdiv), _p(v), _p(i), C.byref(nv), C.byref(ni))
        return v, i

    def triangle_data(self, verts, idx):
        verts, idx = _f(verts), _u(idx)
        out = np.empty((idx.size // 3, 37), np.float32)
        self.fn("triangle_data")(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(out))
        return out

    def sq_dist(self, tri37, pts):
        tri37, pts = _f(tri37), _f(pts)
        out = np.empty(len(pts), np.float32)
        self.fn("sq_dist")(_p(tri37), _p(pts), C.c_uint64(len(pts)), _p(out))
        return out

    def signed_dist(self, tri37, v123, pts, mode):
        tri37, pts, v123 = _f(tri37), _f(pts), _f(v123)
        d = np.empty(len(pts), np.float32)
        g = np.zeros((len(pts), 3), np.float32)
        self.fn("signed_dist")(_p(tri37), _p(v123), _p(pts), C.c_uint64(len(pts)), C.c_int(mode), _p(d), _p(g))
        return d, g

    def tricubic_coefficients(self, values8x8, node_size):
        out = np.empty(64, np.float32)
        self.fn("tricubic_coefficients")(_p(_f(values8x8)), C.c_float(node_size), _p(out))
        return out

    def tricubic_eval(self, coeff64, frac, node_size=1.0):
        c, frac = _f(coeff64), _f(frac)
        n = len(frac)
        val = np.empty(n, np.float32)
        grad = np.empty((n, 3), np.float32)
        vv = np.empty((n, 8), np.float32)
        self.fn("tricubic_eval")(_p(c), _p(frac), C.c_uint64(n), _p(val), _p(grad), _p(vv), C.c_float(node_size))
        return val, grad, vv

    def error_estimate(self, coeff64, mid19x8, rule=1, decay=0.0):
        return float(self.fn("error_estimate")(_p(_f(coeff64)), _p(_f(mid19x8)), C.c_int(rule), C.c_float(decay)))

    def is_near_minimize(self, half, radius8, tri9, thr):
        it = C.c_uint32()
        r = self.fn("is_near_minimize")(C.c_float(half), _p(_f(radius8)), _p(_f(tri9)), C.c_float(thr), C.byref(it))
        return bool(r), it.value

    def filter_triangles(self, verts, idx, center, half, in_tris, corner_tris):
        verts, idx, in_tris = _f(verts), _u(idx), _u(in_tris)
        out = np.empty(max(len(in_tris), 1), np.uint32)
        n = self.fn("filter_triangles")(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size),
                                        _p(_f(center)), C.c_float(half), _p(in_tris), C.c_uint32(len(in_tris)),
                                        _p(_u(corner_tris)), _p(out))
        return out[:n].copy()

    def nearest_triangle(self, verts, idx, pts):
        verts, idx, pts = _f(verts), _u(idx), _f(pts)
        out = np.empty(len(pts), np.uint32)
        self.fn("nearest_triangle")(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(pts),
                                    C.c_uint64(len(pts)), _p(out))
        return out

    def nearest_triangle_visits(self, verts, idx, pts):
        """Port only: nearest triangle plus (inner nodes, leaves) visited per query — traversal models in tests/."""
        verts, idx, pts = _f(verts), _u(idx), _f(pts)
        out = np.empty(len(pts), np.uint32)
        visits = np.empty((len(pts), 2), np.uint32)
        self.fn("nearest_triangle_visits")(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(pts),
                                           C.c_uint64(len(pts)), _p(out), _p(visits))
        return out, visits

    def nearest_triangle_visits_seeded(self, verts, idx, pts, seeds):
        """Port only: as nearest_triangle_visits, the running best of query i starting from seeds[i] (float64)."""
        verts, idx, pts = _f(verts), _u(idx), _f(pts)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.float64)   # None: the box-pruning model (no seed)
        out = np.empty(len(pts), np.uint32)
        visits = np.empty((len(pts), 2), np.uint32)
        self.fn("nearest_triangle_visits_seeded")(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(pts),
                                                  C.c_uint64(len(pts)), None if seeds is None else _p(seeds), _p(out), _p(visits))
        return out, visits

    # ---- structures -----------------------------------------------------------------------
    def build_octree(self, verts, idx, box6, depth, start_depth, threshold=1e-3, algorithm=1, num_threads=1,
                     termination_rule=1, param1=0.0, use_cache=True):
        verts, idx = _f(verts), _u(idx)
        args = [_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(_f(box6)), C.c_uint32(depth),
                C.c_uint32(start_depth), C.c_int(termination_rule), C.c_float(threshold), C.c_float(param1),
                C.c_int(algorithm), C.c_uint32(num_threads)]
        t0 = time.perf_counter()
        if self.kind == "port":
            h = self.fn("build_octree")(*args, C.c_int(1 if use_cache else 0))
            secs = time.perf_counter() - t0
        else:
            s = C.c_double()
            h = self.fn("build_octree")(*args, C.byref(s))
            secs = s.value
        return Sdf(self, h, secs)

    def build_exact(self, verts, idx, box6, max_depth, start_depth=1, min_tris=128, num_threads=1, use_cache=True):
        verts, idx = _f(verts), _u(idx)
        args = [_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(_f(box6)), C.c_uint32(max_depth),
                C.c_uint32(start_depth), C.c_uint32(min_tris), C.c_uint32(num_threads)]
        t0 = time.perf_counter()
        if self.kind == "port":
            h = self.fn("build_exact")(*args, C.c_int(1 if use_cache else 0))
            secs = time.perf_counter() - t0
        else:
            s = C.c_double()
            h = self.fn("build_exact")(*args, C.byref(s))
            secs = s.value
        return Sdf(self, h, secs)

    def load(self, path):
        return Sdf(self, self.fn("load")(path.encode()))

    def max_threads(self):
        return os.cpu_count() or 1


class Sdf:
    """Handle on an oracle-side SdfFunction (OctreeSdf or ExactOctreeSdf)."""

    def __init__(self, backend, handle, build_seconds=None):
        if not handle:
            raise RuntimeError(f"{backend.kind} oracle returned a null SdfFunction")
        self.b = backend
        self.h = C.c_void_p(handle)
        self.build_seconds = build_seconds
        self.last_query_seconds = None

    def save(self, path):
        return bool(self.b.fn("save")(self.h, path.encode()))

    def format(self):
        return self.b.fn("format")(self.h)

    def sample_area(self):
        out = np.empty(6, np.float32)
        self.b.fn("sample_area")(self.h, _p(out))
        return out

    def octree_data(self):
        n = self.b.fn("octree_data_size")(self.h)
        out = np.empty(n * (2 if self.format() == 2 else 1), np.uint32)
        self.b.fn("octree_data")(self.h, _p(out))
        return out

    def header(self):
        sg, md, a, b = C.c_int(), C.c_uint32(), C.c_float(), C.c_float()
        self.b.fn("octree_header")(self.h, C.byref(sg), C.byref(md), C.byref(a), C.byref(b))
        if self.format() == 2:
            return dict(start_grid_size=sg.value, max_depth=md.value, min_tris=int(a.value), max_tris=int(b.value))
        return dict(start_grid_size=sg.value, max_depth=md.value, value_range=a.value, min_border_value=b.value)

    def query(self, pts, gradient=False, num_threads=1):
        pts = _f(pts)
        d = np.empty(len(pts), np.float32)
        g = np.zeros((len(pts), 3), np.float32) if gradient else None
        self.last_query_seconds = self.b.fn("query")(self.h, _p(pts), C.c_uint64(len(pts)), _p(d), _p(g),
                                                     C.c_int(num_threads))
        return (d, g) if gradient else d

    def close(self):
        if self.h:
            self.b.fn("delete")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ref = Backend("ref_", os.path.join(_HERE, "_ref", "libsdfref.so"), "reference")
port = Backend("orc_", os.path.join(_HERE, "liboracle.so"), "port")

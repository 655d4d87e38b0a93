"""TEST INFRASTRUCTURE — ctypes loader for oracle/_ref/libsdfref.so (the unmodified reference).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this.
Every function forwards to oracle/ref_capi.cpp, which in turn forwards to the reference classes.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsdfref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_build_octree.restype = C.c_void_p
        L.ref_build_exact.restype = C.c_void_p
        L.ref_load.restype = C.c_void_p
        L.ref_octree_data_size.restype = C.c_uint64
        L.ref_query.restype = C.c_double
        L.ref_error_estimate.restype = C.c_float
        L.ref_filter_triangles.restype = C.c_uint32
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def isosphere(subdiv):
    nv, ni = C.c_uint32(), C.c_uint32()
    lib().ref_isosphere(C.c_uint32(subdiv), None, None, C.byref(nv), C.byref(ni))
    v = np.empty((nv.value, 3), np.float32)
    i = np.empty(ni.value, np.uint32)
    lib().ref_isosphere(C.c_uint32(subdiv), _p(v), _p(i), C.byref(nv), C.byref(ni))
    return v, i


def triangle_data(verts, idx):
    verts, idx = _f(verts), _u(idx)
    out = np.empty((idx.size // 3, 37), np.float32)
    lib().ref_triangle_data(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(out))
    return out


def sq_dist(tri37, pts):
    tri37, pts = _f(tri37), _f(pts)
    out = np.empty(len(pts), np.float32)
    lib().ref_sq_dist(_p(tri37), _p(pts), C.c_uint64(len(pts)), _p(out))
    return out


def signed_dist(tri37, v123, pts, mode):
    tri37, pts, v123 = _f(tri37), _f(pts), _f(v123)
    d = np.empty(len(pts), np.float32)
    g = np.zeros((len(pts), 3), np.float32)
    lib().ref_signed_dist(_p(tri37), _p(v123), _p(pts), C.c_uint64(len(pts)), C.c_int(mode), _p(d), _p(g))
    return d, g


def tricubic_coefficients(values8x8, node_size):
    v = _f(values8x8)
    out = np.empty(64, np.float32)
    lib().ref_tricubic_coefficients(_p(v), C.c_float(node_size), _p(out))
    return out


def tricubic_eval(coeff64, frac, node_size=1.0):
    c, frac = _f(coeff64), _f(frac)
    n = len(frac)
    val = np.empty(n, np.float32)
    grad = np.empty((n, 3), np.float32)
    vv = np.empty((n, 8), np.float32)
    lib().ref_tricubic_eval(_p(c), _p(frac), C.c_uint64(n), _p(val), _p(grad), _p(vv), C.c_float(node_size))
    return val, grad, vv


def error_estimate(coeff64, mid19x8, rule=1, decay=0.0):
    return float(lib().ref_error_estimate(_p(_f(coeff64)), _p(_f(mid19x8)), C.c_int(rule), C.c_float(decay)))


def is_near_minimize(half, radius8, tri9, thr):
    it = C.c_uint32()
    r = lib().ref_is_near_minimize(C.c_float(half), _p(_f(radius8)), _p(_f(tri9)), C.c_float(thr), C.byref(it))
    return bool(r), it.value


def filter_triangles(verts, idx, center, half, in_tris, corner_tris):
    verts, idx, in_tris = _f(verts), _u(idx), _u(in_tris)
    out = np.empty(len(in_tris), np.uint32)
    n = lib().ref_filter_triangles(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(_f(center)),
                                   C.c_float(half), _p(in_tris), C.c_uint32(len(in_tris)), _p(_u(corner_tris)),
                                   _p(out))
    return out[:n].copy()


def nearest_triangle(verts, idx, pts):
    verts, idx, pts = _f(verts), _u(idx), _f(pts)
    out = np.empty(len(pts), np.uint32)
    lib().ref_nearest_triangle(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(pts),
                               C.c_uint64(len(pts)), _p(out))
    return out


class RefSdf:
    """Handle on a reference SdfFunction (OctreeSdf or ExactOctreeSdf)."""

    def __init__(self, handle, build_seconds=None):
        if not handle:
            raise RuntimeError("reference returned a null SdfFunction")
        self.h = C.c_void_p(handle)
        self.build_seconds = build_seconds

    @staticmethod
    def build_octree(verts, idx, box6, depth, start_depth, threshold=1e-3, algorithm=1, num_threads=1,
                     termination_rule=1, param1=0.0):
        verts, idx = _f(verts), _u(idx)
        secs = C.c_double()
        h = lib().ref_build_octree(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(_f(box6)),
                                   C.c_uint32(depth), C.c_uint32(start_depth), C.c_int(termination_rule),
                                   C.c_float(threshold), C.c_float(param1), C.c_int(algorithm),
                                   C.c_uint32(num_threads), C.byref(secs))
        return RefSdf(h, secs.value)

    @staticmethod
    def build_exact(verts, idx, box6, max_depth, start_depth=1, min_tris=128, num_threads=1):
        verts, idx = _f(verts), _u(idx)
        secs = C.c_double()
        h = lib().ref_build_exact(_p(verts), C.c_uint32(len(verts)), _p(idx), C.c_uint32(idx.size), _p(_f(box6)),
                                  C.c_uint32(max_depth), C.c_uint32(start_depth), C.c_uint32(min_tris),
                                  C.c_uint32(num_threads), C.byref(secs))
        return RefSdf(h, secs.value)

    @staticmethod
    def load(path):
        return RefSdf(lib().ref_load(path.encode()))

    def save(self, path):
        return bool(lib().ref_save(self.h, path.encode()))

    def format(self):
        return lib().ref_format(self.h)

    def sample_area(self):
        out = np.empty(6, np.float32)
        lib().ref_sample_area(self.h, _p(out))
        return out

    def octree_data(self):
        n = lib().ref_octree_data_size(self.h)
        width = 2 if self.format() == 2 else 1
        out = np.empty(n * width, np.uint32)
        lib().ref_octree_data(self.h, _p(out))
        return out

    def header(self):
        sg, md, a, b = C.c_int(), C.c_uint32(), C.c_float(), C.c_float()
        lib().ref_octree_header(self.h, C.byref(sg), C.byref(md), C.byref(a), C.byref(b))
        return dict(start_grid_size=sg.value, max_depth=md.value, value_range=a.value, min_border_value=b.value)

    def query(self, pts, gradient=False, num_threads=1):
        pts = _f(pts)
        d = np.empty(len(pts), np.float32)
        g = np.zeros((len(pts), 3), np.float32) if gradient else None
        secs = lib().ref_query(self.h, _p(pts), C.c_uint64(len(pts)), _p(d), _p(g) if gradient else None,
                               C.c_int(num_threads))
        self.last_query_seconds = secs
        return (d, g) if gradient else d

    def close(self):
        if self.h:
            lib().ref_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def max_threads():
    return lib().ref_max_threads()

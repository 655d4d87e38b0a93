"""Synthetic meshes of the benchmark configs: the real Armadillo / Dragon / Lucy scans
are not available offline, so the configs use the reference's own icosphere generator with a closed-form
multi-octave displacement that puts the surface in generic position w.r.t. the octree lattice."""
import ctypes as C

import numpy as np

from . import _capi


def isosphere(subdivisions):
    """PrimitivesFactory::getIsosphere(subdivisions): 20 * 4^s triangles, same vertex/triangle order."""
    L = _capi.lib()
    nv, ni = C.c_uint32(), C.c_uint32()
    _capi.check(L.sdfb200_make_isosphere(C.c_uint32(subdivisions), None, None, C.byref(nv), C.byref(ni)))
    v = np.empty((nv.value, 3), np.float32)
    i = np.empty(ni.value, np.uint32)
    _capi.check(L.sdfb200_make_isosphere(C.c_uint32(subdivisions), _capi.ptr(v), _capi.ptr(i), C.byref(nv), C.byref(ni)))
    return v, i


def displace(vertices, amplitude=0.15, frequency=3.0, octaves=5, offset=(0.013, -0.007, 0.003)):
    """v <- v * (1 + sum_o a 0.5^o sin(f 2^o x + 1.3 o) sin(f 2^o y + 2.1 o) sin(f 2^o z + 0.7 o)) + offset (float32)."""
    v = np.asarray(vertices, np.float32)
    d = np.zeros(len(v), np.float32)
    a, f = np.float32(amplitude), np.float32(frequency)
    for o in range(octaves):
        d += a * np.sin(f * v[:, 0] + np.float32(1.3 * o), dtype=np.float32) \
               * np.sin(f * v[:, 1] + np.float32(2.1 * o), dtype=np.float32) \
               * np.sin(f * v[:, 2] + np.float32(0.7 * o), dtype=np.float32)
        a *= np.float32(0.5)
        f *= np.float32(2.0)
    return (v * (np.float32(1.0) + d)[:, None] + np.asarray(offset, np.float32)).astype(np.float32)


def bounding_box_with_margin(vertices, margin_fraction=0.2):
    """Mesh bbox + margin_fraction * max extent on every side (reference README.md:85-87)."""
    v = np.asarray(vertices, np.float32)
    mn, mx = v.min(0), v.max(0)
    m = np.float32(margin_fraction) * (mx - mn).max()
    return np.concatenate([mn - m, mx + m]).astype(np.float32)


def config_mesh(name):
    """M0 (config 1), M1 Armadillo-class, M2 Dragon-class, M3 Lucy-class."""
    sub = {"M0": 2, "M1": 7, "M2": 9, "M3": 10}[name]
    v, i = isosphere(sub)
    if name != "M0":
        v = displace(v)
    return v, i


def cell_centre_grid(box6, n):
    """n^3 cell centres of the (cubified) octree box, x fastest — the query set of the configs."""
    box6 = np.asarray(box6, np.float32)
    g = (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    p = np.stack([x, y, z], -1).reshape(-1, 3)
    return (box6[:3] + p * (box6[3:] - box6[:3])).astype(np.float32)

"""Python mirror of the reference's public classes for the two hot paths — same constructor arguments,
same getters, same .bin files — on top of the C-ABI (include/sdfb200.h).

  reference                                         here
  sdflib::BoundingBox   (utils/Mesh.h:16-70)        BoundingBox
  sdflib::Mesh          (utils/Mesh.h:72-106)       Mesh(vertices, indices)
  sdflib::SdfFunction   (SdfFunction.h:12-58)       SdfFunction (getDistance / saveToFile / loadFromFile)
  sdflib::OctreeSdf     (OctreeSdf.h:20-292)        OctreeSdf
  sdflib::ExactOctreeSdf(ExactOctreeSdf.h:17-214)   ExactOctreeSdf

getDistance accepts one point or an (n, 3) array — the bulk form is the hot path — as numpy arrays
(host pointers, copies included in the call) or CUDA torch tensors (device pointers, zero copy).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import SdfB200Error  # noqa: F401


class BoundingBox:
    def __init__(self, min_point, max_point):
        self.min = np.asarray(min_point, np.float32).copy()
        self.max = np.asarray(max_point, np.float32).copy()

    def getSize(self):
        return self.max - self.min

    def getCenter(self):
        return self.min + np.float32(0.5) * self.getSize()

    def addMargin(self, margin):
        self.min -= np.float32(margin)
        self.max += np.float32(margin)

    def as_array(self):
        return np.concatenate([self.min, self.max]).astype(np.float32)


class Mesh:
    """Mesh(glm::vec3* vertices, n, uint32_t* indices, n) — the raw-array constructor (src/utils/Mesh.cpp:34-42)."""

    def __init__(self, vertices, indices):
        self.vertices = _capi.f32(vertices).reshape(-1, 3)
        self.indices = _capi.u32(indices).reshape(-1)
        self._bbox = None

    def getVertices(self):
        return self.vertices

    def getIndices(self):
        return self.indices

    def computeBoundingBox(self):
        self._bbox = BoundingBox(self.vertices.min(0), self.vertices.max(0))

    def getBoundingBox(self):
        if self._bbox is None:
            self.computeBoundingBox()
        return self._bbox


# BvhNode of the traversal kernels (include/sdfb200.h, sdfb200_mesh_bvh): two float64 child spheres, links, leaf flag
BVH_NODE = np.dtype([("left_sphere", "<f8", 4), ("right_sphere", "<f8", 4), ("left", "<i4"), ("right", "<i4"), ("leaf", "<i4"), ("pad", "<i4")])


def bvh_host(vertices, indices):
    """sdfb200_bvh_host: the host builder (libstdc++'s std::sort routines), for the parity tests of the device build."""
    v, i = _capi.f32(vertices).reshape(-1, 3), _capi.u32(indices).reshape(-1)
    out = np.zeros(2 * (i.size // 3) - 1, BVH_NODE)
    _capi.check(_capi.lib().sdfb200_bvh_host(_capi.ptr(v), C.c_uint32(len(v)), _capi.ptr(i), C.c_uint32(i.size), _capi.ptr(out), C.c_uint64(out.size)))
    return out


class PreparedMesh:
    """sdfb200_mesh: a mesh ingested on the current device (TriangleData on the GPU; with bvh=True the nearest-triangle BVH
    of the OctreeSdf builders, with exact=True the side arrays of ExactOctreeSdf). One prepared mesh serves any number of
    builds; export_blob()/from_blob() replicate it on another rank or device without repeating the host work."""
    BVH, EXACT, ALL_HOST_THREADS = 1, 2, 4

    def __init__(self, mesh=None, bvh=True, exact=True, all_host_threads=False, _handle=None):
        if _handle is not None:
            self._h = C.c_void_p(_handle)
            return
        parts = (self.BVH if bvh else 0) | (self.EXACT if exact else 0) | (self.ALL_HOST_THREADS if all_host_threads else 0)
        h = C.c_void_p()
        _capi.check(_capi.lib().sdfb200_mesh_create(_capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices),
                                                    C.c_uint32(mesh.indices.size), C.c_int(parts), C.byref(h)))
        self._h = h

    def blob_bytes(self):
        n = C.c_uint64()
        _capi.check(_capi.lib().sdfb200_mesh_blob_bytes(self._h, C.byref(n)))
        return n.value

    def export_blob(self, device_ptr, capacity):
        _capi.check(_capi.lib().sdfb200_mesh_export(self._h, C.c_void_p(device_ptr), C.c_uint64(capacity)))

    @staticmethod
    def from_blob(device_ptr, nbytes):
        h = C.c_void_p()
        _capi.check(_capi.lib().sdfb200_mesh_import(C.c_void_p(device_ptr), C.c_uint64(nbytes), C.byref(h)))
        return PreparedMesh(_handle=h.value)

    def stats(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _capi.check(_capi.lib().sdfb200_mesh_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"triangle_data_ms": a.value, "bvh_ms": b.value, "upload_ms": c.value}

    def bvh_nodes(self, num_triangles):
        """The nearest-triangle BVH this mesh holds on its device (sdfb200_mesh_bvh), as a BVH_NODE record array."""
        out = np.zeros(2 * num_triangles - 1, BVH_NODE)
        _capi.check(_capi.lib().sdfb200_mesh_bvh(self._h, _capi.ptr(out), C.c_uint64(out.size)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().sdfb200_mesh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _adopt(cls, handle):
    obj = cls.__new__(cls)
    SdfFunction.__init__(obj, handle)
    return obj


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


class SdfFunction:
    GRID, OCTREE, EXACT_OCTREE, NONE = 0, 1, 2, 3

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self._info = None

    # ---- reference interface ---------------------------------------------------------------
    def getDistance(self, sample, gradient=False, exact_order=False, out=None, out_gradient=None, stream=None):
        """Signed distance (and unit gradient if gradient=True) at one point or at an (n, 3) batch."""
        L = _capi.lib()
        flags = _capi.QUERY_EXACT_ORDER if exact_order else 0
        if _is_torch_cuda(sample):
            import torch
            pts = sample.reshape(-1, 3)
            if pts.dtype != torch.float32 or not pts.is_contiguous():
                pts = pts.float().contiguous()
            n = pts.shape[0]
            for name, buf, need in (("out", out, n), ("out_gradient", out_gradient if gradient else None, 3 * n)):
                if buf is not None and not (_is_torch_cuda(buf) and buf.dtype == torch.float32 and buf.is_contiguous()
                                            and buf.numel() >= need and buf.device == pts.device):
                    raise ValueError(f"{name} must be a contiguous float32 CUDA tensor on {pts.device} with at least {need} elements")
            dist = out if out is not None else torch.empty(n, dtype=torch.float32, device=pts.device)
            grad = None
            if gradient:
                grad = out_gradient if out_gradient is not None else torch.empty((n, 3), dtype=torch.float32, device=pts.device)
            st = stream if stream is not None else torch.cuda.current_stream(pts.device).cuda_stream
            _capi.check(L.sdfb200_query(self._h, C.c_void_p(pts.data_ptr()), C.c_uint64(n), C.c_void_p(dist.data_ptr()),
                                        C.c_void_p(grad.data_ptr()) if gradient else None,
                                        C.c_int(flags | _capi.QUERY_DEVICE_POINTERS), C.c_void_p(st)))
            return (dist, grad) if gradient else dist
        pts = _capi.f32(sample)
        single = pts.ndim == 1
        pts = pts.reshape(-1, 3)
        n = len(pts)
        for name, buf, need in (("out", out, n), ("out_gradient", out_gradient if gradient else None, 3 * n)):
            if buf is not None and not (isinstance(buf, np.ndarray) and buf.dtype == np.float32 and buf.flags.c_contiguous and buf.size >= need):
                raise ValueError(f"{name} must be a C-contiguous float32 numpy array with at least {need} elements")
        dist = out if out is not None else np.empty(n, np.float32)
        grad = (out_gradient if out_gradient is not None else np.zeros((n, 3), np.float32)) if gradient else None
        _capi.check(L.sdfb200_query(self._h, _capi.ptr(pts), C.c_uint64(n), _capi.ptr(dist), _capi.ptr(grad), C.c_int(flags), None))
        if single:
            return (float(dist[0]), grad[0]) if gradient else float(dist[0])
        return (dist, grad) if gradient else dist

    def getSampleArea(self):
        i = self.info()
        return BoundingBox(np.array(i.box_min[:], np.float32), np.array(i.box_max[:], np.float32))

    def getFormat(self):
        return self.info().format

    def saveToFile(self, path):
        code = _capi.lib().sdfb200_save(self._h, str(path).encode())
        return code == _capi.OK   # reference returns false on file errors (SdfFunction.cpp:12-16)

    @staticmethod
    def loadFromFile(path):
        """Returns an OctreeSdf / ExactOctreeSdf, or None when the file cannot be loaded (SdfFunction.cpp:47-78)."""
        h = C.c_void_p()
        code = _capi.lib().sdfb200_load(str(path).encode(), C.byref(h))
        if code == _capi.ERR_IO:
            return None
        _capi.check(code)
        tmp = SdfFunction(h.value)
        cls = OctreeSdf if tmp.info().format == _capi.FORMAT_OCTREE else ExactOctreeSdf
        obj = cls.__new__(cls)
        SdfFunction.__init__(obj, h.value)
        tmp._h = None
        return obj

    # ---- extras ------------------------------------------------------------------------------
    def info(self):
        if self._info is None:
            i = _capi.Info()
            _capi.check(_capi.lib().sdfb200_get_info(self._h, C.byref(i)))
            self._info = i
        return self._info

    def build_stats(self):
        s = _capi.BuildStats()
        _capi.check(_capi.lib().sdfb200_get_build_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def getGridBoundingBox(self):
        return self.getSampleArea()

    def getStartGridSize(self):
        g = self.info().start_grid_size
        return (g, g, g)

    def getOctreeMaxDepth(self):
        return self.info().max_depth

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().sdfb200_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OctreeSdf(SdfFunction):
    UNIFORM, NO_CONTINUITY, CONTINUITY = 0, 1, 2                      # InitAlgorithm
    NONE_RULE, TRAPEZOIDAL_RULE, SIMPSONS_RULE, BY_DISTANCE_RULE = 0, 1, 2, 3   # TerminationRule

    def __init__(self, mesh, box, depth, startDepth, maxError=1e-3, initAlgorithm=NO_CONTINUITY, numThreads=1,
                 terminationRule=TRAPEZOIDAL_RULE, terminationRuleParams=None):
        params = list(terminationRuleParams) if terminationRuleParams is not None else [maxError]
        params += [0.0] * (2 - len(params))
        h = C.c_void_p()
        if isinstance(mesh, PreparedMesh):
            _capi.check(_capi.lib().sdfb200_build_octree_from_mesh(
                mesh._h, _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(depth), C.c_uint32(startDepth), C.c_int(terminationRule),
                C.c_float(params[0]), C.c_float(params[1]), C.c_int(initAlgorithm), C.c_uint32(numThreads), C.c_uint32(0), C.c_uint32(1),
                C.byref(h)))
        else:
            _capi.check(_capi.lib().sdfb200_build_octree(
                _capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices), C.c_uint32(mesh.indices.size),
                _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(depth), C.c_uint32(startDepth), C.c_int(terminationRule),
                C.c_float(params[0]), C.c_float(params[1]), C.c_int(initAlgorithm), C.c_uint32(numThreads), C.byref(h)))
        super().__init__(h.value)

    @staticmethod
    def build_on_devices(mesh, box, depth, startDepth, devices, maxError=1e-3, initAlgorithm=1, numThreads=2,
                         terminationRule=1, terminationRuleParams=None):
        """sdfb200_build_octree_multi: ONE process, the listed devices of this box; returns one complete OctreeSdf per device."""
        params = list(terminationRuleParams) if terminationRuleParams is not None else [maxError]
        params += [0.0] * (2 - len(params))
        dev = (C.c_int * len(devices))(*devices)
        handles = (C.c_void_p * len(devices))()
        _capi.check(_capi.lib().sdfb200_build_octree_multi(
            _capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices), C.c_uint32(mesh.indices.size),
            _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(depth), C.c_uint32(startDepth), C.c_int(terminationRule),
            C.c_float(params[0]), C.c_float(params[1]), C.c_int(initAlgorithm), C.c_uint32(numThreads), dev, C.c_uint32(len(devices)), handles))
        return [_adopt(OctreeSdf, h) for h in handles]

    def getOctreeData(self):
        out = np.empty(self.info().octree_words, np.uint32)
        _capi.check(_capi.lib().sdfb200_get_octree_data(self._h, _capi.ptr(out), C.c_uint64(out.size)))
        return out

    def sphereTrace(self, origins, directions, far_distance, epsilon=1e-5, max_iterations=1024, exact_order=False):
        """Sphere tracing (the reference viewer's raycast loop over getDistance) of n rays: returns (hit positions (n, 3),
        travelled distance (n,), -1 where no surface was reached, iterations (n,)). numpy arrays or CUDA torch tensors."""
        L = _capi.lib()
        flags = _capi.QUERY_EXACT_ORDER if exact_order else 0
        if _is_torch_cuda(origins):
            import torch
            o, d = origins.reshape(-1, 3).float().contiguous(), directions.reshape(-1, 3).float().contiguous()
            n = o.shape[0]
            hit = torch.empty((n, 3), dtype=torch.float32, device=o.device)
            trav = torch.empty(n, dtype=torch.float32, device=o.device)
            its = torch.empty(n, dtype=torch.int32, device=o.device)
            _capi.check(L.sdfb200_sphere_trace(self._h, C.c_void_p(o.data_ptr()), C.c_void_p(d.data_ptr()), C.c_uint64(n), C.c_float(epsilon),
                                               C.c_float(far_distance), C.c_uint32(max_iterations), C.c_void_p(hit.data_ptr()),
                                               C.c_void_p(trav.data_ptr()), C.c_void_p(its.data_ptr()),
                                               C.c_int(flags | _capi.QUERY_DEVICE_POINTERS), C.c_void_p(torch.cuda.current_stream(o.device).cuda_stream)))
            return hit, trav, its
        o, d = _capi.f32(origins).reshape(-1, 3), _capi.f32(directions).reshape(-1, 3)
        n = len(o)
        hit, trav, its = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty(n, np.uint32)
        _capi.check(L.sdfb200_sphere_trace(self._h, _capi.ptr(o), _capi.ptr(d), C.c_uint64(n), C.c_float(epsilon), C.c_float(far_distance),
                                           C.c_uint32(max_iterations), _capi.ptr(hit), _capi.ptr(trav), _capi.ptr(its), C.c_int(flags), None))
        return hit, trav, its

    def getOctreeValueRange(self):
        return self.info().value_range

    def getOctreeMinBorderValue(self):
        return self.info().min_border_value


class ExactOctreeSdf(SdfFunction):
    def __init__(self, mesh, box, maxDepth, startDepth=1, minTrianglesPerNode=128, numThreads=1):
        h = C.c_void_p()
        if isinstance(mesh, PreparedMesh):
            _capi.check(_capi.lib().sdfb200_build_exact_from_mesh(
                mesh._h, _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(maxDepth), C.c_uint32(startDepth), C.c_uint32(minTrianglesPerNode),
                C.c_uint32(numThreads), C.c_uint32(0), C.c_uint32(1), C.byref(h)))
        else:
            _capi.check(_capi.lib().sdfb200_build_exact(
                _capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices), C.c_uint32(mesh.indices.size),
                _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(maxDepth), C.c_uint32(startDepth),
                C.c_uint32(minTrianglesPerNode), C.c_uint32(numThreads), C.byref(h)))
        super().__init__(h.value)

    @staticmethod
    def build_on_devices(mesh, box, maxDepth, startDepth, devices, minTrianglesPerNode=128, numThreads=2):
        """sdfb200_build_exact_multi: ONE process, the listed devices of this box; returns one complete ExactOctreeSdf per device."""
        dev = (C.c_int * len(devices))(*devices)
        handles = (C.c_void_p * len(devices))()
        _capi.check(_capi.lib().sdfb200_build_exact_multi(
            _capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices), C.c_uint32(mesh.indices.size),
            _capi.ptr(_capi.f32(box.as_array())), C.c_uint32(maxDepth), C.c_uint32(startDepth), C.c_uint32(minTrianglesPerNode),
            C.c_uint32(numThreads), dev, C.c_uint32(len(devices)), handles))
        return [_adopt(ExactOctreeSdf, h) for h in handles]

    def getOctreeData(self):
        out = np.empty(2 * self.info().octree_words, np.uint32)
        _capi.check(_capi.lib().sdfb200_get_octree_data(self._h, _capi.ptr(out), C.c_uint64(out.size)))
        return out.reshape(-1, 2)

    def getTrianglesSets(self):
        i = self.info()
        sets = np.empty(i.triangle_sets_words, np.uint32)
        _capi.check(_capi.lib().sdfb200_get_exact_arrays(self._h, _capi.ptr(sets), None, None))
        return sets

    def getTrianglesMasks(self):
        i = self.info()
        masks = np.empty(i.triangle_masks_bytes, np.uint8)
        _capi.check(_capi.lib().sdfb200_get_exact_arrays(self._h, None, _capi.ptr(masks), None))
        return masks

    def getTrianglesData(self):
        i = self.info()
        tris = np.empty((i.num_triangles, 37), np.float32)
        _capi.check(_capi.lib().sdfb200_get_exact_arrays(self._h, None, None, _capi.ptr(tris)))
        return tris

    def getMaxTrianglesInLeafs(self):
        return self.info().max_triangles_in_leafs

    def getMinTrianglesInLeafs(self):
        return self.info().min_triangles_in_leafs

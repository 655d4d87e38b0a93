"""sdflib_b200 — B200-native implementation of SdfLib's two data-parallel hot paths (octree
construction and bulk getDistance) behind the reference's own API. See DESIGN.md."""
from ._capi import SdfB200Error, lib, LIB_PATH  # noqa: F401
from .sdf import BoundingBox, Mesh, PreparedMesh, SdfFunction, OctreeSdf, ExactOctreeSdf, BVH_NODE, bvh_host  # noqa: F401
from . import meshes  # noqa: F401


def device_count():
    return lib().sdfb200_device_count()

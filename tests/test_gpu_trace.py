"""Sphere tracing over an OctreeSdf (sdfb200_sphere_trace; SURVEY.md 8 row f-4, reference: raycast() of
src/render_engine/shaders/sdfOctreeRender.comp:392-410) against a CPU loop over the oracle's getDistance: with the
reference-order kernel every ray must stop at the same position after the same number of steps, bit for bit; the FMA
kernel must find the same surface within the query tolerance."""
import numpy as np
import pytest

from conftest import assert_bit_equal, displaced_sphere

pytestmark = pytest.mark.gpu


def cpu_march(oracle_sdf, origins, directions, far, eps=1e-5, max_it=1024):
    n = len(origins)
    pos, hit = origins.astype(np.float32).copy(), origins.astype(np.float32).copy()
    acc, last, it = np.zeros(n, np.float32), np.full(n, 1e8, np.float32), np.zeros(n, np.uint32)
    while True:
        act = (last > np.float32(eps)) & (acc < np.float32(far)) & (it < max_it)
        if not act.any():
            break
        idx = np.nonzero(act)[0]
        hit[idx] = pos[idx]
        d = oracle_sdf.query(pos[idx])
        last[idx] = d
        step = np.maximum(d, np.float32(0.0))
        acc[idx] = acc[idx] + step
        pos[idx] = pos[idx] + directions[idx] * step[:, None]
        it[idx] += 1
    return hit, np.where(last < np.float32(eps), acc, np.float32(-1.0)).astype(np.float32), it


def outward_sphere(subdiv):
    """The reference icosphere winds its triangles so that the field is positive inside; a march needs it positive
    at the eye, so the winding is flipped here."""
    v, i = displaced_sphere(subdiv)
    return v, np.ascontiguousarray(i.reshape(-1, 3)[:, [0, 2, 1]]).reshape(i.shape)


def camera_rays(box, n_side, seed):
    rng = np.random.default_rng(seed)
    centre, size = 0.5 * (box[:3] + box[3:]), float((box[3:] - box[:3]).max())
    # the eye sits inside the box near a corner (outside the box getDistance is the distance to the box, so a march
    # that starts there converges onto the box face and never reaches the octree); the fan is wide enough that the
    # outer rays miss the mesh, leave the box and run out to `far`
    eye = (centre + np.float32([0.42, 0.43, 0.44]) * size).astype(np.float32)
    u = np.linspace(-1.5, 1.5, n_side, dtype=np.float32)
    px, py = np.meshgrid(u, u)
    target = centre + np.stack([px.ravel() * size * 0.5, py.ravel() * size * 0.5, np.zeros(px.size, np.float32)], -1)
    target = target + rng.normal(0, 1e-3, target.shape)
    d = (target - eye).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return np.repeat(eye[None], len(d), 0).astype(np.float32), d


@pytest.mark.parametrize("algorithm", [1, 2])
def test_trace_bit_exact_with_cpu_loop(sdf, port, algorithm):
    v, i = outward_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    g = sdf.OctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 6, 3, 1e-3, algorithm, 1)
    p = port.build_octree(v, i, box, 6, 3, 1e-3, algorithm, 1, use_cache=False)
    o, d = camera_rays(box, 48, 5)
    far = 4.0 * float((box[3:] - box[:3]).max())
    want_hit, want_t, want_it = cpu_march(p, o, d, far, max_it=256)
    hit, t, it = g.sphereTrace(o, d, far, max_iterations=256, exact_order=True)
    assert np.array_equal(it, want_it)
    assert_bit_equal(hit, want_hit, "last evaluated position")
    assert_bit_equal(t, want_t, "travelled distance")
    assert (t >= 0).mean() > 0.3 and (t < 0).mean() > 0.05        # rays that reach the surface and rays that miss
    # the FMA kernel: the same surface (hit positions within the query tolerance of the box, a step count that differs by a few)
    fhit, ft, fit = g.sphereTrace(o, d, far, max_iterations=256)
    both = (t >= 0) & (ft >= 0)
    assert both.sum() >= 0.99 * (t >= 0).sum()
    assert np.abs(fhit[both] - hit[both]).max() < 1e-4 * float((box[3:] - box[:3]).max())


def test_trace_device_pointers_and_errors(sdf):
    import torch
    v, i = outward_sphere(2)
    box = sdf.meshes.bounding_box_with_margin(v)
    bb = sdf.BoundingBox(box[:3], box[3:])
    g = sdf.OctreeSdf(sdf.Mesh(v, i), bb, 5, 3)
    o, d = camera_rays(box, 33, 6)          # 1089 rays: not a multiple of the CTA
    far = 4.0 * float((box[3:] - box[:3]).max())
    host = g.sphereTrace(o, d, far, exact_order=True)
    dev = g.sphereTrace(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), far, exact_order=True)
    for a, b in zip(host, dev):
        assert np.array_equal(np.asarray(a).view(np.uint32), b.cpu().numpy().view(np.uint32))
    e = sdf.ExactOctreeSdf(sdf.Mesh(v, i), bb, 4, 1, 16)
    with pytest.raises(sdf.SdfB200Error):
        sdf.OctreeSdf.sphereTrace(e, o, d, far)

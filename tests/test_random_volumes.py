"""Randomised small volumes: sparse voxel sets at random offsets (including far from the origin and across
octant boundaries), baked or edited, against the reference (CPU) and through the CUDA path (GPU)."""
import numpy as np
import pytest

from conftest import assert_hits_identical, reference_hits


def random_volume(ref, rng):
    v = ref.volume()
    size = int(rng.choice([4, 9, 16, 33]))
    # Coordinates stay below 2^23 so that every voxel plane is an exact float: beyond that the planes of
    # neighbouring voxels collapse onto one float, the reference's stack-write guard (raytracing.cpp:285)
    # misfires and it dereferences an uninitialised stack slot (it segfaults at 2^30; we read 0 there).
    centre = rng.choice([0, 1, -1, 64, -4096, 100000, -(1 << 20), (1 << 22)], size=3).astype(np.int64)
    n = int(rng.integers(1, 400))
    xyz = centre + rng.integers(-size, size + 1, size=(n, 3))
    mats = rng.integers(1, 200, size=(n, 1))
    v.set_voxels(np.concatenate([xyz, mats], axis=1).astype(np.int32))
    style = rng.integers(0, 3)
    if style > 0:
        v.bake()
    if style == 2:                         # edited after baking: unshared tail, possibly all-empty nodes
        v.checkpoint()
        c = centre + rng.integers(-size, size + 1, size=3)
        v.fill_sphere(float(c[0]), float(c[1]), float(c[2]), float(rng.integers(1, size)), int(rng.choice([0, 7])))
    return v, centre.astype(np.float64), float(size)


def rays_around(rng, centre, size, n):
    scale = max(size * 3.0, 8.0)
    o = centre + rng.normal(0, 1, (n, 3)) * scale
    t = centre + rng.uniform(-size, size, (n, 3))
    d = t - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)
    inside = rng.random(n) < 0.3                    # some rays start inside the occupied box
    o[inside] = (centre + rng.uniform(-size, size, (n, 3)))[inside]
    rays = np.zeros(n, dtype=[("o", "<f4", 3), ("d", "<f4", 3)])
    rays["o"], rays["d"] = o.astype(np.float32), d.astype(np.float32)
    return rays


@pytest.mark.parametrize("seed", range(12))
def test_port_and_device_core_vs_reference(port, ref, hostcore, seed):
    rng = np.random.default_rng(1000 + seed)
    v, centre, size = random_volume(ref, rng)
    rays = rays_around(rng, centre, size, 3000)
    sd = port.find_subdags(v.nodes(), v.root())
    for surf, mf in [(True, -1.0), (True, 0.0035), (False, 0.2)]:
        got, want, mask = reference_hits(v, port, rays, surf, mf)
        assert_hits_identical(got[mask], want, "seed %d" % seed)
        assert_hits_identical(hostcore(v.nodes(), sd, rays, surf, mf), got, "device core, seed %d" % seed)


@pytest.mark.gpu
def test_gpu_on_random_volumes(gpu, port, ref):
    total_hits = 0
    for seed in range(40):
        rng = np.random.default_rng(5000 + seed)
        v, centre, size = random_volume(ref, rng)
        nodes, root = v.nodes(), v.root()
        gpu.upload(nodes, root)
        rays = rays_around(rng, centre, size, 4000)
        sd = port.find_subdags(nodes, root)
        for surf, mf in [(True, -1.0), (True, 0.0035)]:
            want, _, _ = port.trace(nodes, sd, rays, surf, mf)
            assert_hits_identical(gpu.intersect_volume(rays, surf, mf), want, "seed %d" % seed)
        total_hits += int(want["hit"].sum())
    assert total_hits > 10000

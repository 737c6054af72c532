"""Pins the oracle port (oracle/cbq_oracle.c) to the UNMODIFIED reference (oracle/_ref).

Method = the reference's own testRaytracingBehaviour (commands/test/test_rendering.cpp:23-82), made
stricter: instead of hit counts and a 1e-3 distance tolerance we demand bit equality of every field.
"""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays, reference_hits


def striped_sphere(ref, n=48, holes=0.05, seed=1):
    v = ref.volume()
    g = np.arange(-n // 2, n // 2)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    rng = np.random.default_rng(seed)
    solid = (X * X + Y * Y + Z * Z < (0.44 * n) ** 2) & (rng.random(X.shape) > holes)
    mat = (1 + ((Z + n) // 3) % 5).astype(np.int32)
    v.set_voxels(np.stack([X[solid], Y[solid], Z[solid], mat[solid]], axis=1))
    v.bake()
    return v


def test_hash_known_answers(port, ref):
    # reference commands/test/test_base.cpp:14-35
    assert port.lib.cbqo_bit_mix64(0x0123456789abcdef) == 0x960cbea3c15f985a
    assert ref.lib.ref_bit_mix(0x0123456789abcdef) == 0x960cbea3c15f985a
    assert port.lib.cbqo_fnv1a(b"hello", 5) == 0xa430d84680aabd0b
    assert ref.lib.ref_fnv1a(b"hello", 5) == 0xa430d84680aabd0b


def test_subdags_match_reference(port, ref):
    v = striped_sphere(ref)
    a, b = v.subdags(), port.find_subdags(v.nodes(), v.root())
    for f in ("lower", "height", "node"):   # the pad words are uninitialised in the reference
        assert (a[f] == b[f]).all(), f


@pytest.mark.parametrize("surface", [True, False])
@pytest.mark.parametrize("max_footprint", [-1.0, 0.0035, 0.05])
def test_intersect_volume_bit_exact(port, ref, surface, max_footprint):
    v = striped_sphere(ref)
    _, lo, hi = v.bounds()
    rays = mixed_rays(lo, hi, 60000, seed=7)
    got, want, mask = reference_hits(v, port, rays, surface, max_footprint)
    assert mask.mean() > 0.999
    assert (want["distance"] < 0).sum() > 0, "ray set should exercise quirk Q1 (negative-distance hits)"
    assert_hits_identical(got[mask], want, "port vs reference")


def test_off_centre_single_octant_volume(port, ref):
    # All voxels in the +++ octant: seven empty sub-DAGs (node 0, height 31), one deep chain.
    v = ref.volume()
    g = np.arange(100, 132)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    solid = ((X + Y + Z) % 7) < 3
    v.set_voxels(np.stack([X[solid], Y[solid], Z[solid], np.full(solid.sum(), 9)], axis=1))
    v.bake()
    nodes, root = v.nodes(), v.root()
    sd = port.find_subdags(nodes, root)
    assert (sd["node"] > 0).sum() == 1
    rays = mixed_rays([90, 90, 90], [140, 140, 140], 20000, seed=3)
    got, want, mask = reference_hits(v, port, rays, True, -1.0)
    assert_hits_identical(got[mask], want, "single octant")


def test_unbaked_edited_volume(port, ref):
    # After checkpoint + fillBrush the array has an unshared tail and duplicate nodes; still must agree.
    v = striped_sphere(ref, n=32, holes=0.0)
    v.checkpoint()
    v.fill_sphere(6, -3, 4, 7, 0)
    v.fill_sphere(-8, 2, -5, 4, 3)
    nodes, root = v.nodes(), v.root()
    assert v.shared_end() < len(nodes)
    rays = mixed_rays([-16, -16, -16], [16, 16, 16], 20000, seed=5)
    for mf in (-1.0, 0.0035):
        got, want, mask = reference_hits(v, port, rays, True, mf)
        # Quirk Q6: fillBrush leaves all-empty internal nodes behind; rays whose hit lands on one make
        # the reference spin in findNearestMaterial. The port abandons exactly those.
        assert 0 < (~mask).sum() < 0.01 * len(rays)
        assert_hits_identical(got[mask], want, "edited volume")


def test_brute_force_checker_agrees_where_defined(port, ref):
    # The reference's own differential test: ESVO vs traceRayRef on in-bounds rays.
    from cubiquity_b200 import rays as R
    v = striped_sphere(ref, n=32, holes=0.02)
    nodes, root = v.nodes(), v.root()
    _, lo, hi = v.bounds()
    rays = R.in_bounds_rays(1000, lo, hi, seed=0)
    got, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, True, -1.0)
    brute = v.trace_ref(rays)
    both = (got["hit"] == 1) & (brute["hit"] == 1) & (got["distance"] >= 0)
    assert both.sum() > 300
    assert np.abs(got["distance"][both] - brute["distance"][both]).max() < 1e-3   # test_rendering.cpp:78
    assert (got["material"][both] == brute["material"][both]).all()


def test_degenerate_rays_terminate(port, ref):
    # Quirk Q5: the reference never returns for these, so only the port is exercised.
    v = striped_sphere(ref, n=32, holes=0.0)
    nodes, root = v.nodes(), v.root()
    rays = np.zeros(64, dtype=port.trace.__globals__["RAY_DTYPE"])
    rays["o"] = np.random.default_rng(0).integers(-20, 20, (64, 3)) + 0.5
    rays["d"] = [0, 0, 1]
    got, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, True, -1.0)
    assert (got["pad"] == 1).any()
    assert (got["hit"][got["pad"] == 1] == 0).all()


def extreme_volume(ref):
    """Voxels at the corners of the 2^32 grid and at the origin: sub-DAG roots at height 31, INT_MIN lower
    bounds, wrap-around in `lowerBound * sign - signBit * size` (raytracing.cpp:454-456)."""
    v = ref.volume()
    lo, hi = -(1 << 31), (1 << 31) - 1
    pts = [(lo, lo, lo, 3), (hi, hi, hi, 4), (lo, hi, 0, 5), (0, 0, 0, 6), (-1, -1, -1, 7), (hi, lo, hi, 8), (1000, -2000, 3000, 9)]
    v.set_voxels(np.array(pts, dtype=np.int32))
    v.bake()
    return v


def extreme_rays(n, seed):
    rng = np.random.default_rng(seed)
    targets = np.array([(-2.0**31, -2.0**31, -2.0**31), (2.0**31, 2.0**31, 2.0**31), (-2.0**31, 2.0**31, 0), (0, 0, 0),
                        (-1, -1, -1), (2.0**31, -2.0**31, 2.0**31), (1000, -2000, 3000)])
    t = targets[rng.integers(0, len(targets), n)]
    o = t + rng.normal(0, 1, (n, 3)) * np.where(np.abs(t) > 1e6, 3e5, 50.0)
    d = (t + rng.normal(0, 0.3, (n, 3)) * np.where(np.abs(t) > 1e6, 200.0, 0.4)) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, dtype=[("o", "<f4", 3), ("d", "<f4", 3)])
    rays["o"], rays["d"] = o.astype(np.float32), d.astype(np.float32)
    return rays


def test_full_height_volume_with_extreme_coordinates(port, ref, hostcore):
    v = extreme_volume(ref)
    sd = port.find_subdags(v.nodes(), v.root())
    # four occupied octants; the one holding a single voxel chains all the way down to a MATERIAL node at
    # height 0 (the "FIXME" case of raytracing.cpp:448-449, quirk Q2), the others branch at height 31
    assert sd["height"].max() == 31 and (sd["node"] > 0).sum() == 4 and ((sd["node"] > 0) & (sd["node"] < 256)).sum() == 1
    rays = extreme_rays(40000, seed=12)
    for surf, mf in [(True, -1.0), (True, 0.0035), (False, 0.05)]:
        got, want, mask = reference_hits(v, port, rays, surf, mf)
        assert mask.mean() > 0.99
        assert_hits_identical(got[mask], want, "extreme coordinates")
        assert_hits_identical(hostcore(v.nodes(), sd, rays, surf, mf), got, "extreme coordinates, device core on host")
    assert want["hit"].sum() > 100

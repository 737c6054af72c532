"""The DEVICE traversal core (csrc/traverse.cuh) compiled for the host and checked against the oracle.

This is how the per-ray state machine is validated on a machine without a GPU. It is test-only: the
product library contains no host path.
"""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays


@pytest.mark.parametrize("kind,size_log2", [("sphere_noise", 7), ("terrain", 8), ("soup", 7), ("city", 10)])
@pytest.mark.parametrize("surface,max_footprint", [(True, -1.0), (False, -1.0), (True, 0.0035), (True, 0.05)])
def test_core_matches_oracle(port, hostcore, scenes, kind, size_log2, surface, max_footprint):
    sc = scenes(kind, size_log2)
    sd = port.find_subdags(sc.nodes, sc.root)
    rays = mixed_rays(sc.lower, sc.upper, 40000, seed=11)
    want, _, _ = port.trace(sc.nodes, sd, rays, surface, max_footprint)
    got = hostcore(sc.nodes, sd, rays, surface, max_footprint)
    assert want["hit"].sum() > 1000
    assert_hits_identical(got, want, "%s %d" % (kind, size_log2))


def test_core_abandons_the_same_rays(port, hostcore, scenes):
    sc = scenes("sphere_noise", 6)
    sd = port.find_subdags(sc.nodes, sc.root)
    rays = np.zeros(256, dtype=[("o", "<f4", 3), ("d", "<f4", 3)])
    rng = np.random.default_rng(4)
    rays["o"] = rng.integers(-30, 30, (256, 3)) + 0.5       # on cell boundaries ...
    rays["d"] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 256)]   # ... with zero components
    want, _, _ = port.trace(sc.nodes, sd, rays, True, -1.0)
    assert want["pad"].sum() > 0
    assert_hits_identical(hostcore(sc.nodes, sd, rays, True, -1.0), want, "degenerate")

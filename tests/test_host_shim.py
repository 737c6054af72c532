"""cbq_expand_hits (csrc/host_shim.cpp, host only): 8-byte compact results back to the 40-byte records. Checked
against the oracle's own hit records: compact them in numpy, expand them with the library, expect the same bytes
(position = origin + dir * distance must come out un-fused, raytracing.cpp:463-466)."""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays


def compact_of(hits, api):
    """numpy statement of the cbq_hit_compact code (include/cubiquity_b200.h)."""
    out = np.zeros(len(hits), dtype=api.COMPACT_DTYPE)
    out["distance"] = hits["distance"]
    code = hits["material"].astype(np.uint32) & 0xff
    nbits = hits["normal"].view(np.uint32).reshape(-1, 3)
    for a in range(3):
        code |= (((nbits[:, a] << 1) != 0).astype(np.uint32) | ((nbits[:, a] >> 31) << 1)) << (8 + 2 * a)
    code |= (hits["hit"] != 0).astype(np.uint32) << 14
    code |= (hits[hits.dtype.names[-1]] != 0).astype(np.uint32) << 15
    out["code"] = code
    return out


@pytest.mark.parametrize("surface,mf", [(True, -1.0), (False, -1.0), (True, 0.05)])
def test_expand_restores_the_oracles_records(api, port, scenes, surface, mf):
    sc = scenes("terrain", 8)
    sd = port.find_subdags(sc.nodes, sc.root)
    rays = mixed_rays(sc.lower, sc.upper, 150000, seed=12)
    rng = np.random.default_rng(5)
    rays["o"][:128] = rng.integers(-60, 60, (128, 3)) + 0.5          # degenerate: some are abandoned
    rays["d"][:128] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 128)]
    want, _, _ = port.trace(sc.nodes, sd, rays, surface, mf)
    assert want["hit"].sum() > 10000 and want["pad"].sum() > 0
    compact = compact_of(want, api)
    assert_hits_identical(api.expand_hits(rays, compact), want.view(api.HIT_DTYPE), "all threads")
    assert_hits_identical(api.expand_hits(rays, compact, threads=1), want.view(api.HIT_DTYPE), "one thread")
    assert_hits_identical(api.expand_hits(rays[:1000], compact[:1000], threads=3), want[:1000].view(api.HIT_DTYPE), "small batch")


def test_expand_of_nothing_and_bad_arguments(api):
    assert len(api.expand_hits(np.zeros(0, dtype=api.RAY_DTYPE), np.zeros(0, dtype=api.COMPACT_DTYPE))) == 0
    L = api.load_library()
    assert L.cbq_expand_hits(None, None, 5, None, 1) == api.ERROR_INVALID_ARGUMENT

"""The procedural scene builder (scenes/scene_builder.cpp, the host-only input generator) against the reference's own volume code."""
import os
import tempfile

import numpy as np
import pytest

from cubiquity_b200 import dagfile


@pytest.mark.parametrize("kind,size_log2", [("sphere_noise", 6), ("terrain", 6), ("soup", 6)])
def test_builder_equals_setvoxel_plus_bake(api, ref, kind, size_log2):
    """Same voxels => same canonical DAG: node COUNT must equal what setVoxel + bake produce
    (reference storage.cpp:396-438, 208-290), and Volume::voxel must agree everywhere."""
    sc = api.Scene(kind, size_log2, seed=3)
    n = 1 << size_log2
    g = np.arange(-n // 2, n // 2, dtype=np.int32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    mats = sc.voxels(xyz)
    assert mats.any() and not mats.all()

    built = ref.volume().load_arrays(sc.nodes, sc.root)          # through the reference's .dag reader
    assert (built.voxels(xyz) == mats).all()
    # a shell of voxels just outside the cube must be empty
    out = xyz[::97].copy()
    out[:, 0] += n
    assert not built.voxels(out).any()

    direct = ref.volume()
    solid = mats > 0
    direct.set_voxels(np.concatenate([xyz[solid], mats[solid, None].astype(np.int32)], axis=1))
    direct.bake()
    assert len(direct.nodes()) == len(sc.nodes)
    sd_a, sd_b = built.subdags(), direct.subdags()
    for f in ("lower", "height"):
        assert (sd_a[f] == sd_b[f]).all()


def test_builder_is_deterministic_and_thread_count_independent(api):
    a = api.Scene("terrain", 8, seed=5)
    b = api.Scene("terrain", 8, seed=5)
    assert a.root == b.root and np.array_equal(a.nodes, b.nodes)
    c = api.Scene("terrain", 8, seed=6)
    assert not (len(c.nodes) == len(a.nodes) and np.array_equal(a.nodes, c.nodes))


def test_city_shares_buildings(api):
    small = api.Scene("city", 10, seed=2)
    big = api.Scene("city", 13, seed=2)
    # 64x the ground area, far fewer than 64x the nodes: whole lots are shared.
    assert len(big.nodes) < 8 * len(small.nodes)
    probe = np.array([[5, 5, -10], [5, 5, 3000], [100, 130, 1]], dtype=np.int32)
    assert big.voxels(probe)[0] == 22 and big.voxels(probe)[1] == 0


def test_dag_file_round_trip(api, ref):
    sc = api.Scene("sphere_noise", 6, seed=1)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.dag")
        dagfile.write_dag(path, sc.nodes, sc.root)
        nodes, root = dagfile.read_dag(path)
        assert root == sc.root and np.array_equal(nodes, sc.nodes)
        v = ref.volume().load(path)
        assert v.root() == sc.root and np.array_equal(v.nodes(), sc.nodes)
        # and the reference's writer (which bakes first) is readable by ours
        out = os.path.join(d, "b.dag")
        v.save(out)
        nodes2, root2 = dagfile.read_dag(out)
        assert len(nodes2) == len(sc.nodes)
        assert np.array_equal(nodes2[:256], dagfile.material_nodes())
        assert root2 == v.root()

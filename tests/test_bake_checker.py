"""The order-independent DAG signature used to judge cbq_bake (tests/test_gpu_bake.py) is itself pinned here against
the reference's Volume::bake (storage.cpp:388-395): an unbaked array and the reference's merge of it must have the
same signature, and the number of distinct nodes must be the reference's merged node count."""
import numpy as np

from oracle import pyoracle


def unbaked_volume(ref, seed, n=4000, side=40):
    rng = np.random.default_rng(seed)
    v = ref.volume()
    xyz = rng.integers(-side, side, size=(n, 3))
    # blocks of identical material so that there is something to merge and something to collapse
    m = 1 + ((xyz[:, 0] // 8 + xyz[:, 1] // 8 + xyz[:, 2] // 8) % 5)
    v.set_voxels(np.column_stack([xyz, m]).astype(np.int32))
    return v


def test_signature_matches_reference_bake(ref):
    for seed in range(3):
        v = unbaked_volume(ref, seed)
        raw, root = v.nodes().copy(), v.root()
        s0, d0 = pyoracle.dag_signature(raw, root)
        v.bake()
        baked, broot = v.nodes(), v.root()
        s1, d1 = pyoracle.dag_signature(baked, broot)
        assert s0 == s1
        assert d0 == d1 == len(baked) - 256      # the reference's merge leaves exactly the distinct nodes
        assert len(raw) > len(baked)


def test_signature_sees_a_changed_voxel(ref):
    v = unbaked_volume(ref, 5)
    s0, _ = pyoracle.dag_signature(v.nodes(), v.root())
    v.set_voxels(np.array([[1, 2, 3, 7]], dtype=np.int32))
    s1, _ = pyoracle.dag_signature(v.nodes(), v.root())
    assert s0 != s1


def test_uniform_cube_collapses_like_the_reference(ref):
    """Eight equal material children become the material (isMaterialNode(const Node&), storage.cpp:69-75)."""
    v = ref.volume()
    cube = np.array([[x, y, z, 4] for x in range(8, 16) for y in range(8, 16) for z in range(8, 16)], dtype=np.int32)
    v.set_voxels(cube)
    s0, d0 = pyoracle.dag_signature(v.nodes(), v.root())
    v.bake()
    s1, d1 = pyoracle.dag_signature(v.nodes(), v.root())
    assert (s0, d0) == (s1, d1)
    assert d1 == len(v.nodes()) - 256

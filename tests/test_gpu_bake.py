"""cbq_bake: Volume::bake (reference storage.cpp:388-395, merge / merge_node :208-290) on the device copy.

Parity = the same canonical DAG as the reference's merge of the same array: equal order-independent signature and
equal node count (oracle.pyoracle.dag_signature, pinned against the reference in tests/test_bake_checker.py), and
bit-identical ray hits before and after."""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays
from oracle import pyoracle
from test_bake_checker import unbaked_volume

pytestmark = pytest.mark.gpu


def check_against_reference_bake(gpu, port, ref, nodes, root, lower, upper, what):
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    rays = mixed_rays(lower, upper, 60000, seed=11)
    gpu.upload(nodes, root)
    before = gpu.intersect_volume(rays, True, -1.0)
    count, new_root = gpu.bake()
    baked = gpu.download_nodes()
    assert len(baked) == count
    v = ref.volume().load_arrays(nodes, root)
    v.bake()
    assert count == len(v.nodes()), what                                   # the canonical DAG has one size
    assert pyoracle.dag_signature(baked, new_root) == pyoracle.dag_signature(v.nodes(), v.root()), what
    assert np.array_equal(baked[:256], np.repeat(np.arange(256, dtype=np.uint32)[:, None], 8, axis=1))
    assert int(baked.max()) < count
    # the device keeps tracing the merged copy: same hits as before, and as the oracle on the downloaded array
    after = gpu.intersect_volume(rays, True, -1.0)
    # Merging changes which cubes are internal nodes (eight equal children collapse), and quirk Q1 reports internal
    # nodes BEHIND the origin as hits (and Q6 gives up on all-empty ones), so only forward hits are comparable.
    forward = (before["status"] == 0) & (after["status"] == 0) & ~((before["hit"] != 0) & (before["distance"] < 0)) \
        & ~((after["hit"] != 0) & (after["distance"] < 0))
    assert forward.mean() > 0.5
    assert_hits_identical(after[forward], before[forward], what + ": before/after")
    sub = port.find_subdags(baked, new_root)
    assert gpu.subdags().tobytes() == sub.tobytes()
    want, _, _ = port.trace(baked, sub, rays, True, -1.0, threads=8)
    assert_hits_identical(after, want, what + ": oracle on the baked array")
    return baked, new_root


def test_bake_of_unbaked_reference_volumes(gpu, port, ref):
    for seed in range(3):
        v = unbaked_volume(ref, seed)
        nodes, root = v.nodes().copy(), v.root()
        baked, _ = check_against_reference_bake(gpu, port, ref, nodes, root, np.array([-40] * 3), np.array([40] * 3), "unbaked %d" % seed)
        assert len(baked) < len(nodes)


def test_bake_after_runtime_edits_drops_history_and_garbage(gpu, port, ref, scenes):
    sc = scenes("sphere_noise", 7)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    rng = np.random.default_rng(3)
    for step in range(5):
        v.checkpoint()
        c = rng.uniform(-40, 40, 3)
        v.fill_sphere(c[0], c[1], c[2], 14.0, 0 if step % 2 == 0 else 3)
    v.undo()
    nodes, root = v.nodes().copy(), v.root()
    baked, _ = check_against_reference_bake(gpu, port, ref, nodes, root, sc.lower, sc.upper, "edited")
    assert len(baked) < len(nodes)


@pytest.mark.parametrize("kind,log2", [("sphere_noise", 7), ("terrain", 9), ("soup", 9), ("city", 16)])
def test_bake_of_canonical_scenes_keeps_them(gpu, port, ref, scenes, kind, log2):
    sc = scenes(kind, log2)
    baked, root = check_against_reference_bake(gpu, port, ref, sc.nodes, sc.root, sc.lower, sc.upper, kind)
    assert len(baked) == len(sc.nodes)                 # the scene builder hash-conses: nothing to merge


def test_bake_is_deterministic_and_idempotent(gpu, ref):
    v = unbaked_volume(ref, 9, n=20000, side=64)
    nodes, root = v.nodes().copy(), v.root()
    outs = []
    for _ in range(3):
        gpu.upload(nodes, root)
        count, r = gpu.bake()
        outs.append((gpu.download_nodes(), r))
    for o, r in outs[1:]:
        assert r == outs[0][1] and np.array_equal(o, outs[0][0])           # CAS races do not show in the result
    count2, r2 = gpu.bake()                                                # bake of a baked array: identity
    assert r2 == outs[0][1] and np.array_equal(gpu.download_nodes(), outs[0][0])
    assert gpu.counter("bake_reachable") == count2 - 256


def test_bake_of_uniform_and_empty_volumes(gpu, port, ref):
    """The whole volume collapses into a material root, exactly like merge_node (storage.cpp:261-263)."""
    v = ref.volume()
    v.set_voxels(np.array([[1, 2, 3, 5]], dtype=np.int32))
    v.set_voxels(np.array([[1, 2, 3, 0]], dtype=np.int32))                 # empty again, but the nodes are still there
    nodes, root = v.nodes().copy(), v.root()
    assert root >= 256
    gpu.upload(nodes, root)
    count, r = gpu.bake()
    v.bake()
    assert (count, r) == (len(v.nodes()), v.root()) == (256, 0)
    rays = mixed_rays(np.array([-10] * 3), np.array([10] * 3), 1000, seed=1)
    assert not gpu.intersect_volume(rays, True, -1.0)["hit"].any()


def test_bake_rejects_cycles_and_keeps_the_volume(gpu, api):
    nodes = np.repeat(np.arange(258, dtype=np.uint32)[:, None], 8, axis=1)
    nodes[256] = [257, 257, 0, 0, 0, 0, 0, 0]
    nodes[257] = [256, 1, 0, 0, 0, 0, 0, 0]
    gpu.upload(nodes, 256)
    with pytest.raises(api.CubiquityError) as e:
        gpu.bake()
    assert e.value.code == api.ERROR_CORRUPT_VOLUME
    assert gpu.node_count() == 258 and np.array_equal(gpu.download_nodes(), nodes)


def test_bake_large_scene_matches_reference_count(gpu, ref, scenes, api):
    """4096^3 terrain (BASELINE configs[1]) with 40 brush edits: millions of nodes, several scan levels."""
    sc = scenes("terrain", 12)
    ed = api.Editable(sc.nodes, sc.root)
    rng = np.random.default_rng(0)
    centre = (sc.lower + sc.upper) / 2.0
    for k in range(40):
        ed.checkpoint()
        c = centre + rng.uniform(-600, 600, 3)
        ed.fill_sphere(c[0], c[1], c[2], 30.0, 0 if k % 2 else 2)
    nodes, root = ed.nodes().copy(), ed.root()
    gpu.upload(nodes, root)
    count, new_root = gpu.bake()
    v = ref.volume().load_arrays(nodes, root)
    v.bake()
    assert count == len(v.nodes())
    assert count < len(nodes)

"""cbq_fill_sphere / cbq_set_root: the reference's runtime edit (checkpoint + fillBrush(SphereBrush), viewer.cpp:165-168,
voxelization.cpp:825-915) applied to the device copy. Parity: the unfolded tree below the new root is exactly the
reference's (oracle.pyoracle.dag_signature with collapse=False), hence every ray hit is bit-identical, quirks included."""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def same_tree(gpu, v, root):
    mine = gpu.download_nodes()
    assert pyoracle.dag_signature(mine, root, collapse=False)[0] == pyoracle.dag_signature(v.nodes(), v.root(), collapse=False)[0]
    return mine


def test_device_edits_match_reference_edits(gpu, port, ref, scenes):
    sc = scenes("sphere_noise", 7)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    gpu.upload(sc.nodes, sc.root)
    rays = mixed_rays(sc.lower, sc.upper, 60000, seed=5)
    rng = np.random.default_rng(7)
    roots = [sc.root]
    shipped = 0
    for step in range(6):
        c = rng.uniform(-45, 45, 3).astype(np.float32)
        r = np.float32(rng.uniform(3, 16))
        m = 0 if step % 2 == 0 else 4
        v.checkpoint()
        v.fill_sphere(c[0], c[1], c[2], r, m)
        h2d = gpu.counter("bytes_h2d")
        root, count = gpu.fill_sphere(c[0], c[1], c[2], r, m)
        shipped += gpu.counter("bytes_h2d") - h2d
        roots.append(root)
        assert count == gpu.node_count() and root < count
        mine = same_tree(gpu, v, root)
        assert np.array_equal(mine[:len(sc.nodes)], sc.nodes)                # nothing that existed was written
        sub = port.find_subdags(v.nodes(), v.root())
        want, _, _ = port.trace(v.nodes(), sub, rays, True, -1.0, threads=8)
        assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "device edit %d" % step)
        got_sub, ref_sub = gpu.subdags(), sub
        assert got_sub["height"].tolist() == ref_sub["height"].tolist() and got_sub["lower"].tolist() == ref_sub["lower"].tolist()
    assert shipped <= 6 * 512                                                # header + sub-DAGs per edit: no node crossed PCIe
    # undo / redo = an earlier / later root over the same array (storage.cpp:373-385)
    for back in (1, 2, 3):
        v.undo()
        gpu.set_root(roots[-1 - back])
        want, _, _ = port.trace(v.nodes(), port.find_subdags(v.nodes(), v.root()), rays, True, -1.0, threads=8)
        assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "undo %d" % back)
    v.redo()
    gpu.set_root(roots[-3])
    same_tree(gpu, v, roots[-3])
    # bake what the edits left: same canonical DAG as the reference's bake of its own edited volume
    count, root = gpu.bake()
    v.bake()
    assert count == len(v.nodes())
    assert pyoracle.dag_signature(gpu.download_nodes(), root) == pyoracle.dag_signature(v.nodes(), v.root())


def test_large_brush_grows_the_array_and_the_work_list(gpu, port, ref, scenes):
    sc = scenes("terrain", 10)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    gpu.upload(sc.nodes, sc.root)
    before = gpu.node_count()
    v.checkpoint()
    v.fill_sphere(10.0, -20.0, -40.0, 220.0, 0)
    root, count = gpu.fill_sphere(10.0, -20.0, -40.0, 220.0, 0)
    assert count - before > (1 << 16)                                        # more than the head-room of the upload
    mine = same_tree(gpu, v, root)
    assert np.array_equal(mine[:len(sc.nodes)], sc.nodes)                    # the attempts that ran out of room wrote nothing that existed
    rays = mixed_rays(sc.lower, sc.upper, 40000, seed=6)
    want, _, _ = port.trace(v.nodes(), port.find_subdags(v.nodes(), v.root()), rays, True, -1.0, threads=8)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "large brush")


def test_brush_that_touches_nothing(gpu, scenes):
    sc = scenes("sphere_noise", 6)
    gpu.upload(sc.nodes, sc.root)
    root, count = gpu.fill_sphere(5000.0, 5000.0, 5000.0, 3.0, 0)             # empty space stays empty
    # copies made on the way down were all put back: below the new root is the old tree, node for node
    mine = gpu.download_nodes()
    assert np.array_equal(mine[root], sc.nodes[sc.root]) and root == len(sc.nodes)
    assert pyoracle.dag_signature(mine, root, collapse=False) == pyoracle.dag_signature(sc.nodes, sc.root, collapse=False)


def test_set_root_argument_checks(gpu, api, scenes):
    sc = scenes("sphere_noise", 6)
    gpu.upload(sc.nodes, sc.root)
    with pytest.raises(api.CubiquityError):
        gpu.set_root(len(sc.nodes) + 10)
    gpu.set_root(sc.root)


@pytest.mark.parametrize("radius", [0.0, 0.6, 1.5, 2.5])
def test_tiny_brushes_take_the_one_block_path(gpu, port, ref, scenes, radius):
    """Radii this small are walked entirely by the one-block kernels (no per-level launches)."""
    sc = scenes("sphere_noise", 6)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    gpu.upload(sc.nodes, sc.root)
    for k, (x, y, z) in enumerate([(0.0, 0.0, 0.0), (10.0, -3.0, 7.0), (-1.0, -1.0, -1.0), (15.5, 15.5, 15.5)]):
        m = 0 if k % 2 == 0 else 6
        v.checkpoint()
        v.fill_sphere(x, y, z, radius, m)
        root, count = gpu.fill_sphere(x, y, z, radius, m)
        same_tree(gpu, v, root)


def test_host_delta_after_a_device_edit_is_refused(gpu, api, scenes):
    sc = scenes("sphere_noise", 6)
    gpu.upload(sc.nodes, sc.root)
    gpu.update(sc.nodes, len(sc.nodes), sc.root)                            # fine: nothing happened on the device
    gpu.fill_sphere(0.0, 0.0, 0.0, 4.0, 0)
    with pytest.raises(api.CubiquityError):
        gpu.update(sc.nodes, len(sc.nodes), sc.root)                        # the device copy has nodes the host never saw
    gpu.upload(sc.nodes, sc.root)
    gpu.update(sc.nodes, len(sc.nodes), sc.root)

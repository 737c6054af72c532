"""cbq_build_dense: a dense grid of material ids -> the canonical DAG, on the device.

Parity = what the reference makes of the same voxels with Volume::setVoxel per voxel followed by Volume::bake
(storage.cpp:388-438): equal order-independent signature and node count (oracle.pyoracle.dag_signature), equal
voxels read back through the reference, and bit-identical ray hits on the downloaded array."""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def blobs(side, seed):
    """A few solid boxes and balls of different materials plus sparse noise."""
    rng = np.random.default_rng(seed)
    z, y, x = np.mgrid[0:side, 0:side, 0:side]
    g = np.zeros((side, side, side), dtype=np.uint8)
    for m in range(1, 6):
        c = rng.integers(side // 4, 3 * side // 4, 3)
        r = rng.integers(side // 8, side // 3)
        g[(x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2 < r * r] = m
    lo = rng.integers(0, side // 2, 3)
    g[lo[2]:lo[2] + side // 4, lo[1]:lo[1] + side // 4, lo[0]:lo[0] + side // 4] = 7
    noise = rng.random((side, side, side)) < 0.002
    g[noise] = 9
    return g


def reference_volume(ref, grid, origin):
    z, y, x = np.nonzero(grid)
    v = ref.volume()
    xyzm = np.column_stack([x + origin[0], y + origin[1], z + origin[2], grid[z, y, x]]).astype(np.int32)
    v.set_voxels(xyzm)
    v.bake()
    return v


@pytest.mark.parametrize("side,origin", [(32, (0, 0, 0)), (32, (-16, -16, -16)), (64, (-64, 0, 64)), (16, (-8, 8, 2 ** 31 - 16)), (4, (-2, -2, -2)),
                                         (64, (-2 ** 31, -2 ** 31, -2 ** 31))])
def test_build_matches_setvoxel_plus_bake(gpu, port, ref, side, origin):
    grid = blobs(side, seed=side + origin[0] % 7)
    count, root = gpu.build_dense(grid, origin)
    nodes = gpu.download_nodes()
    v = reference_volume(ref, grid, origin)
    assert count == len(nodes) == len(v.nodes())
    assert pyoracle.dag_signature(nodes, root) == pyoracle.dag_signature(v.nodes(), v.root())
    # voxels, read back by the reference from OUR array (inside the grid and around it)
    rng = np.random.default_rng(1)
    q = rng.integers(-side // 2, side + side // 2, size=(20000, 3)) + np.asarray(origin, dtype=np.int64)
    q = np.clip(q, -2 ** 31, 2 ** 31 - 1).astype(np.int32)
    mine = ref.volume().load_arrays(nodes, root).voxels(q)
    local = q.astype(np.int64) - np.asarray(origin, dtype=np.int64)
    inside = ((local >= 0) & (local < side)).all(axis=1)
    want = np.zeros(len(q), dtype=mine.dtype)
    want[inside] = grid[local[inside, 2], local[inside, 1], local[inside, 0]]
    assert np.array_equal(mine, want)
    # and the device traces what it built
    if abs(origin[2]) < 2 ** 22 and abs(origin[0]) < 2 ** 22:
        lower = np.asarray(origin, dtype=np.float64)
        rays = mixed_rays(lower, lower + side, 30000, seed=3)
        sub = port.find_subdags(nodes, root)
        assert gpu.subdags().tobytes() == sub.tobytes()
        want_hits, _, _ = port.trace(nodes, sub, rays, True, -1.0, threads=8)
        assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want_hits, "built volume")


def test_build_reproduces_a_scene_from_its_voxels(gpu, scenes):
    """The host scene builder and the device builder agree on the canonical DAG of the same voxels."""
    sc = scenes("sphere_noise", 7)
    side = 128
    z, y, x = np.mgrid[0:side, 0:side, 0:side]
    xyz = np.column_stack([x.ravel() - 64, y.ravel() - 64, z.ravel() - 64]).astype(np.int32)
    grid = sc.voxels(xyz).reshape(side, side, side).astype(np.uint8)
    count, root = gpu.build_dense(grid, (-64, -64, -64), colours=sc.colours)
    assert count == len(sc.nodes)
    assert pyoracle.dag_signature(gpu.download_nodes(), root) == pyoracle.dag_signature(sc.nodes, sc.root)


def test_build_empty_and_full_grids(gpu):
    count, root = gpu.build_dense(np.zeros((8, 8, 8), dtype=np.uint8), (0, 0, 0))
    assert (count, root) == (256, 0)                       # everything collapses into the empty material
    count, root = gpu.build_dense(np.full((8, 8, 8), 3, dtype=np.uint8), (8, 8, 8))
    nodes = gpu.download_nodes()
    assert root >= 256 and count == 256 + 32 - 3           # one chain from the root down to the solid 8^3 cube
    assert sorted(set(nodes[256:].ravel().tolist()) - set(range(256, count))) == [0, 3]


def test_build_argument_checks(gpu, api):
    g = np.zeros((8, 8, 8), dtype=np.uint8)
    with pytest.raises(api.CubiquityError):
        gpu.build_dense(g, (1, 0, 0))                      # not a multiple of half the side
    with pytest.raises(api.CubiquityError):
        gpu.build_dense(g, (0, 0, 2 ** 31 - 4))            # leaves the volume
    with pytest.raises(ValueError):
        gpu.build_dense(np.zeros((8, 8, 4), dtype=np.uint8), (0, 0, 0))
    with pytest.raises(api.CubiquityError):
        gpu.build_dense(np.zeros((2, 2, 2), dtype=np.uint8), (0, 0, 0))


@pytest.mark.parametrize("side,origin,brick_log2", [(64, (-64, 0, 64), 4), (64, (-32, -32, -32), 5), (128, (0, 0, 0), 5), (32, (-2 ** 31, 0, 2 ** 31 - 32), 3)])
def test_bricked_build_equals_setvoxel_plus_bake(gpu, ref, side, origin, brick_log2):
    """The path cbq_build_dense takes past 1024^3 -- bricks built and merged one by one, a host-made top, one last merge --
    forced onto small grids: the same canonical DAG as the reference's setVoxel + bake, empty and uniform bricks included."""
    grid = blobs(side, seed=side + brick_log2)
    grid[: side // 2, : side // 4, :] = 0                   # whole bricks of nothing ...
    grid[side // 2:, side // 2:, side // 2:] = 5            # ... and of one material
    gpu.set_option("dense_brick_log2", brick_log2)
    try:
        count, root = gpu.build_dense(grid, origin)
    finally:
        gpu.set_option("dense_brick_log2", 0)
    nodes = gpu.download_nodes()
    v = reference_volume(ref, grid, origin)
    assert count == len(nodes) == len(v.nodes())
    assert pyoracle.dag_signature(nodes, root) == pyoracle.dag_signature(v.nodes(), v.root())
    one_piece, root1 = gpu.build_dense(grid, origin)
    assert one_piece == count
    assert pyoracle.dag_signature(gpu.download_nodes(), root1) == pyoracle.dag_signature(nodes, root)


def test_build_2048_cubed(gpu, port):
    """Past the one-piece limit: a 2048^3 grid (8.6 GB of voxels, 64 bricks of 512^3) that holds a 1024^3 model in one octant
    is the same volume as the model built in one piece at the same place -- equal node count, equal signature, identical hits."""
    import torch
    dev = torch.device("cuda", 0)
    small = torch.from_numpy(blobs(256, seed=11)).to(dev)
    model = small.repeat_interleave(4, 0).repeat_interleave(4, 1).repeat_interleave(4, 2).contiguous()      # 1024^3
    origin = (0, 0, -2048)
    count1, root1 = gpu.build_dense(None, (0, 0, -1024), device_ptr=model.data_ptr(), size_log2=10)
    nodes1 = gpu.download_nodes()
    big = torch.zeros((2048, 2048, 2048), dtype=torch.uint8, device=dev)
    big[1024:, :1024, :1024] = model                        # [z, y, x]: x, y in [0, 1024), z in [-1024, 0)
    del model
    count2, root2 = gpu.build_dense(None, origin, device_ptr=big.data_ptr(), size_log2=11)
    del big
    nodes2 = gpu.download_nodes()
    assert count2 == count1
    assert pyoracle.dag_signature(nodes2, root2) == pyoracle.dag_signature(nodes1, root1)
    rays = mixed_rays(np.array([0.0, 0.0, -1024.0]), np.array([1024.0, 1024.0, 0.0]), 30000, seed=5)
    sub = port.find_subdags(nodes2, root2)
    assert gpu.subdags().tobytes() == sub.tobytes()
    want_hits, _, _ = port.trace(nodes2, sub, rays, True, -1.0, threads=8)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want_hits, "2048^3 build")


def test_baseline_terrain_built_from_its_voxels(gpu, scenes):
    """BASELINE configs[1]'s scene family at 512^3: every voxel evaluated on the device (tests/terrain_voxels.py, checked against the
    scene library on the CPU), cbq_build_dense in bricks of 128^3 -> the DAG the host scene builder makes, node for node.
    scripts/build_baseline_scene.py does the same at 4096^3."""
    import torch
    import terrain_voxels
    k = 9
    sc = scenes("terrain", k, 1)
    n = 1 << k
    grid = terrain_voxels.fill(torch.empty((n, n, n), dtype=torch.uint8, device="cuda"), k, 1)
    gpu.set_option("dense_brick_log2", 7)
    try:
        count, root = gpu.build_dense(None, (-n // 2,) * 3, device_ptr=grid.data_ptr(), size_log2=k, colours=sc.colours)
    finally:
        gpu.set_option("dense_brick_log2", 0)
    assert count == len(sc.nodes)
    assert pyoracle.dag_signature(gpu.download_nodes(), root) == pyoracle.dag_signature(sc.nodes, sc.root)

"""cbq_voxelize against the reference's voxelize() (src/library/voxelization.cpp:692-744) run on the box:
every voxel of the grid must carry the same material."""
import numpy as np
import pytest

import meshes

pytestmark = pytest.mark.gpu


def grid_points(origin, size):
    z, y, x = np.mgrid[0:size, 0:size, 0:size]
    return np.stack([x.ravel() + origin[0], y.ravel() + origin[1], z.ravel() + origin[2]], axis=1).astype(np.int32)


def reference_grid(ref, tris, mats, fill, origin, size, thin=False):
    v = ref.volume()
    closed, inside_out, secs = v.voxelize(tris, mats, fill, 0, thin)
    return v.voxels(grid_points(origin, size)).reshape(size, size, size), closed, inside_out, v


def device_grid(gpu, api, scene_voxels, tris, mats, fill, log2, origin, thin=False):
    count, root, info = gpu.voxelize(tris, mats, fill, log2, origin, thin=thin)
    nodes = gpu.download_nodes()
    return nodes, root, info


def voxels_of(port_or_ref, nodes, root, origin, size, ref):
    v = ref.volume().load_arrays(nodes, root)
    return v.voxels(grid_points(origin, size)).reshape(size, size, size)


@pytest.mark.parametrize("origin", [(0, 0, 0), (-64, -64, -64)])
def test_closed_soup_voxel_for_voxel(gpu, api, ref, origin):
    """Two icospheres, a long box (faces split by drawLargeTriangle) and a box on integer coordinates, filled; once in a
    grid aligned to its own size and once centred on the origin (straddling the octree's octant planes)."""
    size, log2 = 128, 7
    tris, mats = meshes.soup(offset=origin)
    want, closed, inside_out, _ = reference_grid(ref, tris, mats, 11, origin, size)
    assert closed and not inside_out
    nodes, root, info = device_grid(gpu, api, None, tris, mats, 11, log2, origin)
    assert info.is_closed == 1 and info.is_inside_out == 0
    got = voxels_of(None, nodes, root, origin, size, ref)
    assert (want == 11).sum() > 50000 and len(np.unique(want)) >= 6          # interior + the four surface materials + empty
    differing = np.argwhere(got != want)
    assert len(differing) == 0, "%d voxels differ, first %s: got %d want %d" % (len(differing), differing[0], got[tuple(differing[0])], want[tuple(differing[0])])
    assert gpu.counter("voxelize_leaves") > 10000 and gpu.counter("voxelize_pieces") > len(tris)
    # and nothing outside the grid
    outside = ref.volume().load_arrays(nodes, root).voxels(np.array([[origin[0] - 1, origin[1], origin[2]], [origin[0] + size, origin[1] + 5, origin[2] + 5]], dtype=np.int32))
    assert not outside.any()


def test_open_mesh_gives_the_shell_in_triangle_order(gpu, api, ref):
    size, log2, origin = 64, 6, (0, 0, 0)
    a, am = meshes.icosphere([30.2, 31.7, 29.4], 18.6, 2, 4)
    b, bm = meshes.box([10.5, 12.25, 9.75], [50.1, 40.6, 30.3], 9)
    tris, mats = np.concatenate([a[150:], b]), np.concatenate([am[150:], bm])        # half a sphere: not closed
    want, closed, _, _ = reference_grid(ref, tris, mats, 11, origin, size)
    assert not closed
    count, root, info = gpu.voxelize(tris, mats, 11, log2, origin)
    assert info.is_closed == 0
    got = voxels_of(None, gpu.download_nodes(), root, origin, size, ref)
    assert (want > 0).sum() > 3000 and not (want == 11).any()
    assert np.array_equal(got, want)


def test_thin_mesh_and_argument_checks(gpu, api, ref):
    size, log2, origin = 64, 6, (0, 0, 0)
    tris, mats = meshes.icosphere([31.4, 30.9, 32.2], 20.3, 2, 6)
    want, closed, _, _ = reference_grid(ref, tris, mats, 2, origin, size, thin=True)
    count, root, info = gpu.voxelize(tris, mats, 2, log2, origin, thin=True)
    got = voxels_of(None, gpu.download_nodes(), root, origin, size, ref)
    assert closed and np.array_equal(got, want) and (want == 6).sum() > (want == 2).sum() * 0.05
    with pytest.raises(api.CubiquityError):                      # does not fit
        gpu.voxelize(tris + np.float32(40), mats, 2, log2, origin)
    with pytest.raises(api.CubiquityError):                      # inside out
        gpu.voxelize(meshes.flipped(tris), mats, 2, log2, origin)
    with pytest.raises(api.CubiquityError):                      # background other than 0
        gpu.voxelize(tris, mats, 2, log2, origin, background=1)
    with pytest.raises(api.CubiquityError):                      # origin not a multiple of S / 2
        gpu.voxelize(tris, mats, 2, log2, (8, 0, 0))


def test_2048_grid_gives_the_voxels_of_the_1024_grid(gpu):
    """Voxels are decided one by one from the mesh alone, so a mesh that fits [0, 1024)^3 voxelises to the same volume in a
    2048^3 grid (43 GB of work space, dense build in bricks) as in a 1024^3 grid at the same origin: the canonical DAGs are equal."""
    from oracle import pyoracle
    k = 8.0
    parts = [meshes.icosphere(np.array([40.3, 44.1, 50.7]) * k, 21.4 * k, 3, 3), meshes.icosphere(np.array([78.2, 70.9, 60.2]) * k, 17.8 * k, 3, 7),
             meshes.box(np.array([20.25, 80.5, 30.75]) * k, np.array([100.6, 95.1, 41.2]) * k, 5)]
    tris, mats = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    count1, root1, _ = gpu.voxelize(tris, mats, 11, 10, (0, 0, 0))
    nodes1 = gpu.download_nodes()
    count2, root2, _ = gpu.voxelize(tris, mats, 11, 11, (0, 0, 0))
    nodes2 = gpu.download_nodes()
    assert count2 == count1 > 10000
    assert pyoracle.dag_signature(nodes2, root2) == pyoracle.dag_signature(nodes1, root1)

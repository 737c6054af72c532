"""tests/terrain_voxels.py (torch) against the scene library's own Terrain::voxel, voxel for voxel, on the CPU."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import terrain_voxels


@pytest.mark.parametrize("size_log2,seed", [(6, 1), (7, 5)])
def test_torch_terrain_equals_the_scene_library(scenes, size_log2, seed):
    n = 1 << size_log2
    sc = scenes("terrain", size_log2, seed)
    grid = terrain_voxels.fill(torch.empty((n, n, n), dtype=torch.uint8), size_log2, seed).numpy()
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    xyz = np.column_stack([x.ravel() - n // 2, y.ravel() - n // 2, z.ravel() - n // 2]).astype(np.int32)
    want = sc.voxels(xyz).reshape(n, n, n)
    assert len(np.unique(want)) > 5
    assert np.array_equal(grid, want)

"""The voxels of the "terrain" scene (scenes/scene_builder.cpp: Terrain::voxel), evaluated with torch integer arithmetic so that a
dense grid of BASELINE configs[1] (4096^3, 68.7 GB) can be produced where cbq_build_dense wants it: in device memory. Same integer
lattice noise, same hash, same material rules; checked voxel for voxel against the scene library on the CPU (tests/test_terrain_voxels.py)."""
import torch

_M64 = (1 << 64)


def _s64(c):
    return c - _M64 if c >= (1 << 63) else c


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix64(x):
    x = x ^ _lsr(x, 30)
    x = x * _s64(0xbf58476d1ce4e5b9)
    x = x ^ _lsr(x, 27)
    x = x * _s64(0x94d049bb133111eb)
    return x ^ _lsr(x, 31)


def hash3(x, y, z, seed):
    """scene_builder.cpp:hash3 on int64 tensors (x, y) and a Python int z: upper 32 bits of three chained mix64."""
    h = _mix64((x * _s64(0x9e3779b97f4a7c15)) ^ _s64(seed))
    h = _mix64(h ^ (y * _s64(0xc2b2ae3d27d4eb4f)))
    zc = (z * 0x165667b19e3779f9) % _M64
    h = _mix64(h ^ _s64(zc))
    return _lsr(h, 32)


def _octave(x, y, cell_log2, o, seed):
    cx, cy = x >> cell_log2, y >> cell_log2
    mask = (1 << cell_log2) - 1
    fx, fy = ((x & mask) << 16) >> cell_log2, ((y & mask) << 16) >> cell_log2
    fx = ((fx * fx >> 16) * (3 * 65536 - 2 * fx)) >> 16
    fy = ((fy * fy >> 16) * (3 * 65536 - 2 * fy)) >> 16
    v00, v10 = hash3(cx, cy, o, seed) & 0xffff, hash3(cx + 1, cy, o, seed) & 0xffff
    v01, v11 = hash3(cx, cy + 1, o, seed) & 0xffff, hash3(cx + 1, cy + 1, o, seed) & 0xffff
    a = v00 + (((v10 - v00) * fx) >> 16)
    b = v01 + (((v11 - v01) * fx) >> 16)
    return a + (((b - a) * fy) >> 16)


def columns(size_log2, seed, device):
    """Per column (y, x): surface height, rock-strata warp, biome -- int64 tensors [N, N]."""
    n = 1 << size_log2
    half = n // 2
    ax = torch.arange(-half, half, dtype=torch.int64, device=device)
    y, x = torch.meshgrid(ax, ax, indexing="ij")
    h = torch.zeros_like(x)
    amp, o, cell = n // 5, 0, size_log2 - 2
    while cell >= 3 and amp > 0:
        h = h + (((_octave(x, y, cell, o, seed) - 32768) * amp) >> 15)
        o, cell, amp = o + 1, cell - 1, amp * 7 // 16
    surface = h - n // 16
    warp = hash3(x >> 7, y >> 7, 77, seed) & 31
    biome = hash3(x >> 8, y >> 8, 99, seed) & 3
    return surface, warp, biome


def fill(grid, size_log2, seed=1, layers=16):
    """grid: uint8 tensor [N, N, N] indexed [z, y, x]; voxel (x, y, z) of the scene sits at [z + N/2, y + N/2, x + N/2]."""
    n = 1 << size_log2
    half = n // 2
    surface, warp, biome = columns(size_log2, seed, grid.device)
    s = surface.to(torch.int32)[None]
    w = warp.to(torch.int32)[None]
    turf = (1 + biome).to(torch.uint8)[None]
    soil = (5 + (biome & 1)).to(torch.uint8)[None]
    snow, sand = (s > n // 10), (s < -(n // 7))
    for z0 in range(-half, half, layers):
        z = torch.arange(z0, min(z0 + layers, half), dtype=torch.int32, device=grid.device)[:, None, None]
        depth = s - z
        rock = (8 + (((z + w) >> 5) & 7)).to(torch.uint8)
        m = torch.where(depth == 1, turf, soil)
        m = torch.where(sand, torch.where(depth <= 3, 18, 19).to(torch.uint8), m)
        m = torch.where(snow, torch.where(depth <= 2, 16, 17).to(torch.uint8), m)
        m = torch.where(depth > 4, rock, m)
        grid[z0 + half:z0 + half + z.shape[0]] = torch.where(depth >= 1, m, torch.zeros_like(m))
    return grid

"""Synthetic ray sets for the BASELINE.json configurations (host side, numpy, seeded)."""
import numpy as np

from .api import RAY_DTYPE


def _unit_ball_dirs(rng, n):
    """Directions = normalised points drawn uniformly in the unit ball by rejection."""
    out = np.empty((n, 3), dtype=np.float64)
    filled = 0
    while filled < n:
        need = n - filled
        cand = rng.uniform(-1.0, 1.0, size=(int(need * 2.2) + 16, 3))
        r2 = (cand * cand).sum(axis=1)
        cand = cand[(r2 < 1.0) & (r2 > 1e-12)]
        take = min(len(cand), need)
        out[filled:filled + take] = cand[:take]
        filled += take
    out /= np.sqrt((out * out).sum(axis=1, keepdims=True))
    return out.astype(np.float32)


def random_rays(n, lower, upper, seed=0, dilate=0.1):
    """C3: collision-query style rays, origin ~ U(bounds dilated by `dilate`), random direction."""
    rng = np.random.Generator(np.random.Philox(seed))
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    ext = (upper - lower) * dilate
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = rng.uniform(lower - ext, upper + ext, size=(n, 3)).astype(np.float32)
    rays["d"] = _unit_ball_dirs(rng, n)
    return rays


def in_bounds_rays(n, lower, upper, seed=0):
    """The reference's own test distribution (commands/test/test_rendering.cpp:29-49): origin and target
    both uniform inside the occupied bounds, direction = normalize(target - origin)."""
    rng = np.random.Generator(np.random.Philox(seed))
    lower = np.asarray(lower, dtype=np.float32)
    upper = np.asarray(upper, dtype=np.float32)
    o = rng.uniform(lower, upper, size=(n, 3)).astype(np.float32)
    t = rng.uniform(lower, upper, size=(n, 3)).astype(np.float32)
    d = (t - o).astype(np.float32)
    length = np.sqrt((d * d).sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    length[length == 0] = 1
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = o
    rays["d"] = (d / length).astype(np.float32)
    return rays


def hit_voxels(hits):
    """The integer cell a hit lies in, derived identically for oracle and GPU results:
    v = floor(position - 0.5 * normal + 0.5). Only meaningful where hits['hit'] != 0."""
    p = hits["position"].astype(np.float64) - 0.5 * hits["normal"].astype(np.float64) + 0.5
    return np.floor(p).astype(np.int64)


def _counter_uniform(seed, i, k):
    """csrc/trace_kernels.cu:counterUniform in numpy (uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = np.uint64(seed) + (np.uint64(8) * i + np.uint64(k + 1)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def counter_rays(indices, n, lower, upper, seed=0):
    """Host twin of cbq_random_rays_device: rays `indices` of the n-ray batch (float32 arithmetic, no FMA)."""
    i = np.asarray(indices, dtype=np.uint64)
    lo = np.asarray(lower, dtype=np.float32)
    ext = (np.asarray(upper, dtype=np.float32) - lo).astype(np.float32)
    rays = np.zeros(len(i), dtype=RAY_DTYPE)
    for a in range(3):
        rays["o"][:, a] = lo[a] + _counter_uniform(seed, i, a) * ext[a]
    d = np.zeros((len(i), 3), dtype=np.float32)
    d[:, 2] = 1.0
    done = np.zeros(len(i), dtype=bool)
    two, one = np.float32(2.0), np.float32(1.0)
    for t in range(5):
        j = i + np.uint64(t) * np.uint64(n)
        x = _counter_uniform(seed, j, 3) * two - one
        y = _counter_uniform(seed, j, 4) * two - one
        z = _counter_uniform(seed, j, 5) * two - one
        r2 = (x * x + y * y) + z * z
        ok = (~done) & (r2 < one) & (r2 > np.float32(1e-12))
        length = np.sqrt(np.where(ok, r2, one)).astype(np.float32)
        d[ok, 0] = (x / length)[ok]; d[ok, 1] = (y / length)[ok]; d[ok, 2] = (z / length)[ok]
        done |= ok
    rays["d"] = d
    return rays

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

"""Generates the golden vectors in this directory from the UNMODIFIED reference.

    python tests/golden/make_golden.py        (needs /root/reference; run in the build container)

Everything stored here was produced by the reference's own code through oracle/ref_shim.cpp:
  sphere32.dag        written by Cubiquity::Volume::save (storage.cpp:530-542) after setVoxel + bake
  raycast.npz         rays + Cubiquity::intersectVolume results (raytracing.cpp:397-478) for several
                      (computeSurfaceProperties, maxFootprint) pairs, findSubDAGs output, and the same
                      for an edited, un-baked copy of the volume (checkpoint + fillBrush)
The GPU box has no /root/reference; these files are how the pin travels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import mixed_rays  # noqa: E402
from oracle import pyoracle  # noqa: E402

CASES = [(1, -1.0), (0, -1.0), (1, 0.0035), (1, 0.05)]


def main():
    ref, port = pyoracle.Ref(), pyoracle.Port()
    v = ref.volume()
    n = 32
    g = np.arange(-n // 2, n // 2)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    rng = np.random.default_rng(42)
    solid = (X * X + Y * Y + Z * Z < 14 * 14) & (rng.random(X.shape) > 0.04)
    mat = (1 + ((X + 2 * Z + 3 * n) // 5) % 7).astype(np.int32)
    v.set_voxels(np.stack([X[solid], Y[solid], Z[solid], mat[solid]], axis=1))
    path = os.path.join(HERE, "sphere32.dag")
    v.save(path)                                   # bakes, then writes
    v = ref.volume().load(path)
    nodes, root = v.nodes(), v.root()

    rays = mixed_rays([-16] * 3, [16] * 3, 4000, seed=99)
    # keep only rays the reference can finish (quirks Q5/Q6 never return); the port tells us which
    keep = np.ones(len(rays), dtype=bool)
    for surf, mf in CASES:
        got, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, surf, mf)
        keep &= got["pad"] == 0
    rays = rays[keep]
    out = {"rays": rays, "root": np.uint32(root), "subdags": v.subdags()}
    for i, (surf, mf) in enumerate(CASES):
        out["hits_%d" % i] = v.intersect(rays, surf, mf)[0]

    # edited, un-baked volume
    v.checkpoint()
    v.fill_sphere(5, -2, 3, 6, 0)
    v.fill_sphere(-7, 3, -4, 3, 2)
    enodes, eroot = v.nodes(), v.root()
    ekeep = np.ones(len(rays), dtype=bool)
    for surf, mf in CASES:
        got, _, _ = port.trace(enodes, port.find_subdags(enodes, eroot), rays, surf, mf)
        ekeep &= got["pad"] == 0
    erays = rays[ekeep]
    out.update({"edited_nodes": enodes, "edited_root": np.uint32(eroot), "edited_shared_end": np.uint32(v.shared_end()),
                "edited_rays": erays, "edited_subdags": v.subdags()})
    for i, (surf, mf) in enumerate(CASES):
        out["edited_hits_%d" % i] = v.intersect(erays, surf, mf)[0]
    out["cases"] = np.array(CASES, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "raycast.npz"), **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "raycast.npz", os.path.getsize(os.path.join(HERE, "raycast.npz")), "bytes;",
          len(rays), "rays,", int(out["hits_0"]["hit"].sum()), "hits,", len(erays), "edited rays")


if __name__ == "__main__":
    main()

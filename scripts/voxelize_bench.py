#!/usr/bin/env python
"""cbq_voxelize against the reference's voxelize() on the same mesh (two icospheres + two boxes, scaled to the grid).
    python scripts/voxelize_bench.py [--log2 8] [--subdivisions 3]
Prints one JSON line per size: GPU milliseconds (whole call: upload, kernels, dense build + bake), reference seconds on the host,
voxels that differ."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cubiquity_b200 import api  # noqa: E402
from oracle import pyoracle  # noqa: E402
import meshes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, nargs="+", default=[7, 8])
ap.add_argument("--subdivisions", type=int, default=3)
ap.add_argument("--no-reference", action="store_true")
args = ap.parse_args()
ref = pyoracle.Ref()
ctx = api.Context(0)
for log2 in args.log2:
    size = 1 << log2
    k = size / 128.0
    parts = [meshes.icosphere(np.array([40.3, 44.1, 50.7]) * k, 21.4 * k, args.subdivisions, 3), meshes.icosphere(np.array([78.2, 70.9, 60.2]) * k, 17.8 * k, args.subdivisions, 7),
             meshes.box(np.array([20.25, 80.5, 30.75]) * k, np.array([100.6, 95.1, 41.2]) * k, 5), meshes.box(np.array([60, 20, 20]) * k, np.array([76, 36, 44]) * k, 9)]
    tris, mats = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    ctx.voxelize(tris, mats, 11, log2, (0, 0, 0))          # warm-up (pool allocations)
    t0 = time.perf_counter()
    count, root, info = ctx.voxelize(tris, mats, 11, log2, (0, 0, 0))
    gpu_ms = 1e3 * (time.perf_counter() - t0)
    line = {"grid": "%d^3" % size, "triangles": int(len(tris)), "pieces": ctx.counter("voxelize_pieces"), "leaves_classified": ctx.counter("voxelize_leaves"),
            "gpu_ms": gpu_ms, "nodes": count}
    if not args.no_reference:
        v = ref.volume()
        closed, inside_out, secs = v.voxelize(tris, mats, 11)
        line["reference_s"] = secs
        line["speedup"] = secs / (gpu_ms * 1e-3)
        z, y, x = np.mgrid[0:size:1, 0:size:1, 0:size:1] if size <= 256 else np.mgrid[0:size:2, 0:size:2, 0:size:2]
        pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int32)
        want = v.voxels(pts)
        got = ref.volume().load_arrays(ctx.download_nodes(), root).voxels(pts)
        line["voxels_compared"] = int(len(pts))
        line["voxels_differing"] = int((want != got).sum())
    print(json.dumps(line), flush=True)

#!/usr/bin/env python
"""A few launches of the bench frame (primary rays, 8x4-tile order, LOD off, buffer order) for ncu to capture.

    ncu --set full --import-source on --clock-control none -k regex:tracePersistent --launch-skip 2 -c 1 \
        -o gpurun_out/trace python scripts/one_frame.py [--lod 0.0035] [--random N] [--opt key=value]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lod", type=float, default=-1.0)
ap.add_argument("--random", type=int, default=0)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--frames", type=int, default=1, help="orbit poses in the batch (bench.py's job is 64)")
ap.add_argument("--compact", action="store_true", help="8-byte results (cbq_trace_compact_device), as bench.py's job")
ap.add_argument("--opt", action="append", default=[])
args = ap.parse_args()
W, H = 1920, 1080
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
scene = api.Scene("terrain", 12, 1)
ctx = api.Context(0)
ctx.upload(scene.nodes, scene.root, scene.colours)
ctx.set_option("adaptive_order", 0)
ctx.set_option("refill_threshold", 8 if args.random else 32)
for kv in args.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
lower = np.asarray(scene.lower, dtype=np.float64); upper = np.asarray(scene.upper, dtype=np.float64)
centre = (lower + upper) * 0.5
hd = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
PI_F = float(np.float32(3.14159265358979))
n = args.random or W * H * args.frames
rays = torch.empty(n * 6, dtype=torch.float32, device=dev)
if args.random:
    ext = (upper - lower) * 0.1
    ctx.random_rays_device(100, (lower - ext).astype(np.float32), (upper + ext).astype(np.float32), n, rays.data_ptr(), stream)
else:
    for f in range(args.frames):
        yaw = f * (2.0 * np.pi / max(args.frames, 1))
        cam = api.camera_from_pose([centre[0] - hd * np.sin(yaw), centre[1] - hd * np.cos(yaw), centre[2] + hd], -(PI_F / 4.0), yaw)
        ctx.primary_rays_tiled_device(cam, W, H, rays.data_ptr() + f * W * H * 24, None, stream)
hits = torch.zeros(n * (2 if args.compact else 10), dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for _ in range(args.launches):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if args.compact:
        ctx.trace_compact_device(rays.data_ptr(), n, hits.data_ptr(), True, args.lod, stream)
    else:
        ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, args.lod, stream)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
print("done: %d rays, ms per launch %s, last %.3f Grays/s" % (n, [round(t, 3) for t in times], n / times[-1] / 1e6))

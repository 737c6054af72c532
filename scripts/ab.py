#!/usr/bin/env python
"""A/B probe for kernel variants: the same measurements for whichever library CBQ_LIBRARY names.

    CBQ_LIBRARY=cubiquity_b200/lib/<variant>.so python scripts/ab.py [--tag name] [--opt key=value ...]

Prints one JSON line: cold primary rays (adaptive_order 0, L2 flushed), replayed order, LOD 0.0035, 8 M random rays,
1080p path tracing. CUDA-event timed on the launching stream; never run under a profiler for numbers."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default=os.path.basename(os.environ.get("CBQ_LIBRARY", "default")))
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--scene-log2", type=int, default=12)
ap.add_argument("--frames", type=int, default=1, help="camera poses traced by one launch (rays concatenated)")
ap.add_argument("--skip", default="", help="comma list of: replay,lod,random,pt")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
skip = set(args.skip.split(","))

W, H = 1920, 1080
PI_F = float(np.float32(3.14159265358979))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream
scene = api.Scene("terrain", args.scene_log2, 1)
ctx = api.Context(0)
ctx.upload(scene.nodes, scene.root, scene.colours)
for kv in args.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def pose(i):
    lower = np.asarray(scene.lower, dtype=np.float64); upper = np.asarray(scene.upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    hd = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    yaw = i * (2.0 * np.pi / 8.0)
    return api.camera_from_pose([centre[0] - hd * np.sin(yaw), centre[1] - hd * np.cos(yaw), centre[2] + hd], -(PI_F / 4.0), yaw)


def timed(fn, steps=args.steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms)), float(np.min(ms))


n = W * H * args.frames
rays = torch.empty(n * 6, dtype=torch.float32, device=dev)
for f in range(args.frames):
    ctx.primary_rays_tiled_device(pose(f), W, H, rays.data_ptr() + f * W * H * 24, None, stream)
hits = torch.zeros(n * 10, dtype=torch.int32, device=dev)
out = {"tag": args.tag, "opts": args.opt, "frames": args.frames}

thr0 = [o for o in args.opt if o.startswith("refill_threshold=")]
ctx.set_option("refill_threshold", int(thr0[0].split("=")[1]) if thr0 else 32)
ctx.set_option("adaptive_order", 0)
ms, best = timed(lambda: ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, -1.0, stream))
out["primary_cold_grays"] = n / ms / 1e6
out["primary_cold_ms"] = ms
out["hit_checksum"] = int(hits.view(torch.int32).to(torch.int64).sum().item())
if "replay" not in skip:
    ctx.set_option("adaptive_order", 1)
    ms, best = timed(lambda: ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, -1.0, stream), warm=6)
    out["primary_replay_grays"] = n / ms / 1e6
    ctx.set_option("adaptive_order", 0)
if "lod" not in skip:
    ms, best = timed(lambda: ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, 0.0035, stream))
    out["primary_lod_cold_grays"] = n / ms / 1e6
if "random" not in skip:
    m = 8_000_000
    rr = torch.empty(m * 6, dtype=torch.float32, device=dev)
    rh = torch.zeros(m * 10, dtype=torch.int32, device=dev)
    ext = (np.asarray(scene.upper, dtype=np.float64) - np.asarray(scene.lower, dtype=np.float64)) * 0.1
    ctx.random_rays_device(100, (scene.lower - ext).astype(np.float32), (scene.upper + ext).astype(np.float32), m, rr.data_ptr(), stream)
    thr = [o for o in args.opt if o.startswith("refill_threshold=")]
    ctx.set_option("refill_threshold", int(thr[0].split("=")[1]) if thr else 8)
    ms, best = timed(lambda: ctx.trace_device(rr.data_ptr(), m, rh.data_ptr(), True, -1.0, stream), steps=5)
    out["random_grays"] = m / ms / 1e6
    out["random_checksum"] = int(rh.view(torch.int32).to(torch.int64).sum().item())
    del rr, rh
if "pt" not in skip:
    thr = [o for o in args.opt if o.startswith("refill_threshold=")]
    ctx.set_option("refill_threshold", int(thr[0].split("=")[1]) if thr else 8)
    acc = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    p = api.pt_params(W, H, spp=8, bounces=4, variant=api.VARIANT_RECURSIVE)
    ms, best = timed(lambda: ctx.render_device(pose(0), p, acc.data_ptr(), stream), steps=3, warm=1)
    out["pt_mspp"] = W * H * 8 / ms / 1e3
    out["pt_mean"] = float(acc.mean().item())
print(json.dumps(out), flush=True)

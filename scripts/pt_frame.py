#!/usr/bin/env python
"""The bench's path-traced 1080p frame on its own: timing for A/B of options, or a target for ncu.

    python scripts/pt_frame.py [--spp 16] [--bounces 4] [--opt sample_group=16] [--renders 4]
    ncu --set full --import-source on --clock-control none -k regex:'shadeKernel|lightKernel|accumulateKernel' --launch-skip 11 -c 11 \
        -o gpurun_out/shade python scripts/pt_frame.py --spp 8 --renders 2"""
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
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--bounces", type=int, default=4)
ap.add_argument("--renders", type=int, default=4)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--opt", action="append", default=[])
args = ap.parse_args()
W, H = args.width, args.height
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
scene = api.Scene("terrain", 12, 1)
ctx = api.Context(0)
ctx.upload(scene.nodes, scene.root, scene.colours)
for kv in args.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
import bench  # noqa: E402
cam, _, _ = bench.orbit_camera(api, scene, 0)
p = api.pt_params(W, H, spp=args.spp, bounces=args.bounces, variant=api.VARIANT_RECURSIVE, frame_id=0)
acc = torch.zeros(W * H * 3, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for i in range(args.renders):
    acc.zero_(); flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.render_device(cam, p, acc.data_ptr(), stream); b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
ms = float(np.median(times[1:])) if len(times) > 1 else times[0]
print(json.dumps({"spp": args.spp, "bounces": args.bounces, "options": args.opt, "ms": round(ms, 3), "spp_per_s_M": round(W * H * args.spp / ms / 1e3, 1),
                  "mean_radiance": float(acc.mean().item()) / args.spp, "checksum": int(acc.view(torch.int32).to(torch.int64).sum().item())}))

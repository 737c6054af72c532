#!/usr/bin/env python
"""Kernel-time sweep over launch options on the bench workload (1080p primary rays, 4096^3 terrain).
Prints one line per configuration: cold-L2 and warm-L2 ms per frame and Grays/s. Development aid."""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402
from cubiquity_b200 import rays as R  # noqa: E402

W, H = 1920, 1080
PI_F = float(np.float32(3.14159265358979))


def time_config(ctx, d_rays, n, d_hits, flush, stream, steps=10, mf=-1.0, surface=True):
    def step():
        ctx.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), surface, mf, stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    cold = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        cold.append(a.elapsed_time(b))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    return float(np.mean(cold)), a.elapsed_time(b) / steps


def render_sweep(ctx, scene, dev, stream):
    cam = api.default_camera(scene.lower, scene.upper)
    acc = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    for mode, variant, bounces, spp, thr, grp in [(0, 1, 4, 16, 8, 1), (0, 1, 4, 16, 8, 2), (0, 1, 4, 16, 8, 4), (0, 1, 4, 16, 8, 8), (0, 1, 4, 16, 8, 16),
                                                  (0, 1, 4, 16, 4, 8), (0, 1, 4, 16, 2, 8), (0, 1, 1, 16, 8, 8), (0, 0, 1, 16, 8, 8)]:
        ctx.set_option("sample_group", grp)
        ctx.set_option("refill_threshold", thr)
        p = api.pt_params(W, H, spp=spp, bounces=bounces, variant=variant)
        ctx.render_device(cam, api.pt_params(W, H, spp=1, bounces=bounces, variant=variant), acc.data_ptr(), stream)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            acc.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ctx.render_device(cam, p, acc.data_ptr(), stream); b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = float(np.median(times))
        print(json.dumps({"workload": "render", "lib": os.environ.get("CBQ_LIBRARY", "default"), "mode": "wavefront" if mode == 0 else "megakernel", "variant": variant, "bounces": bounces, "spp": spp,
                          "refill_threshold": thr, "sample_group": grp, "ms": round(ms, 3), "mspp_per_s": round(W * H * spp / ms / 1e3, 1),
                          "mean": round(float(acc.mean().item()) / spp, 5)}), flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "primary"
    scene = api.Scene(os.environ.get("CBQ_SWEEP_SCENE", "terrain"), int(os.environ.get("CBQ_SWEEP_LOG2", "12")), 1)
    ctx = api.Context(0)
    ctx.upload(scene.nodes, scene.root, scene.colours)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    if what == "render":
        render_sweep(ctx, scene, dev, stream)
        return
    if what == "tiled_refill":
        cam = api.default_camera(scene.lower, scene.upper)
        n = W * H
        d_tiled = torch.empty(n * 6, dtype=torch.float32, device=dev)
        ctx.primary_rays_tiled_device(cam, W, H, d_tiled.data_ptr(), None, stream)
        d_hits = torch.zeros(n * 10, dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for thr, q in ((32, 1), (16, 16), (24, 16), (16, 8), (8, 8), (24, 8), (32, 32)):
            ctx.set_option("refill_threshold", thr)
            ctx.set_option("refill_quantum", q)
            for mf in (-1.0, 0.0035):
                cold, warm = time_config(ctx, d_tiled, n, d_hits, flush, stream, mf=mf)
                print(json.dumps({"workload": "tiled_refill", "refill_threshold": thr, "refill_quantum": q, "max_footprint": mf, "cold_ms": round(cold, 4), "cold_grays": round(n / cold / 1e6, 3)}), flush=True)
        return
    if what == "frame":
        # fused camera path (8x4-pixel tiles per warp) vs the row-major ray buffer, and the L2 window on/off
        cam = api.default_camera(scene.lower, scene.upper)
        n = W * H
        d_rays = torch.empty(n * 6, dtype=torch.float32, device=dev)
        ctx.primary_rays_device(cam, W, H, d_rays.data_ptr(), stream)
        d_tiled = torch.empty(n * 6, dtype=torch.float32, device=dev)
        ctx.primary_rays_tiled_device(cam, W, H, d_tiled.data_ptr(), None, stream)
        d_hits = torch.zeros(n * 10, dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ctx.set_option("refill_threshold", 32)
        for l2 in (1, 0):
            ctx.set_option("l2_persist", l2)
            for mf in (-1.0, 0.0035):
                for name, fn in (("buffer_row_major", lambda: ctx.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), True, mf, stream)),
                                 ("buffer_8x4_tiles", lambda: ctx.trace_device(d_tiled.data_ptr(), n, d_hits.data_ptr(), True, mf, stream)),
                                 ("fused_camera_tiles", lambda: ctx.raycast_frame_device(cam, W, H, d_hits.data_ptr(), True, mf, stream))):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(10):
                        flush.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(); fn(); b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                    print(json.dumps({"workload": "frame", "path": name, "l2_persist": l2, "max_footprint": mf, "cold_ms": round(float(np.mean(ts)), 4),
                                      "cold_grays": round(n / float(np.mean(ts)) / 1e6, 3)}), flush=True)
        return
    if what == "random":
        n = 8_000_000
        host = R.random_rays(n, scene.lower, scene.upper, seed=100)
        d_rays = torch.from_numpy(host.view(np.float32).reshape(-1)).to(dev)
    else:
        n = W * H
        cam = api.default_camera(scene.lower, scene.upper)
        d_rays = torch.empty(n * 6, dtype=torch.float32, device=dev)
        ctx.primary_rays_device(cam, W, H, d_rays.data_ptr(), stream)
    d_hits = torch.zeros(n * 10, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    grid = {
        "refill_threshold": [1, 4, 8, 16, 32],
        "blocks_per_sm": [4],
        "block_threads": [256, 128],
        "l2_persist": [1],
    }
    if len(sys.argv) > 2:
        grid = json.loads(sys.argv[2])
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        for k, v in zip(keys, combo):
            ctx.set_option(k, v)
        for mf in (-1.0, 0.0035):
            cold, warm = time_config(ctx, d_rays, n, d_hits, flush, stream, mf=mf)
            print(json.dumps({"workload": what, **dict(zip(keys, combo)), "max_footprint": mf, "cold_ms": round(cold, 4), "warm_ms": round(warm, 4),
                              "cold_grays": round(n / cold / 1e6, 3), "warm_grays": round(n / warm / 1e6, 3)}), flush=True)


if __name__ == "__main__":
    main()

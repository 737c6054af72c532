"""Is sorting secondary rays worth it? (BASELINE north star: "sorted secondary rays for warp coherence".)
Diffuse bounce rays of a 1080p frame (origin = primary hit + 0.01 n, direction = normalize(n + point in the unit
ball), the path tracer's rule) are traced, with the path tracer's LOD, in four orders: pixel order (what the wavefront
tracer does: ballot compaction keeps it), stably sorted by direction octant, sorted by octant then 8x4 tile, and shuffled.
Also the fully incoherent random batch of BASELINE configs[2], as is and sorted by octant.

    python scripts/sort_probe.py [--out gpurun_out/sort.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

W, H = 1920, 1080


def timed(ctx, rays, n, hits, flush, stream, mf, steps=8):
    for _ in range(2):
        ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, mf, stream)
    torch.cuda.synchronize()
    t = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), True, mf, stream); b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return float(np.median(t))


def octant(rays):
    d = rays[:, 3:6]
    return ((d[:, 0] < 0).to(torch.int64) | ((d[:, 1] < 0).to(torch.int64) << 1) | ((d[:, 2] < 0).to(torch.int64) << 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    sc = api.Scene("terrain", 12, 1)
    ctx = api.Context(0)
    ctx.upload(sc.nodes, sc.root, sc.colours)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cam = api.default_camera(sc.lower, sc.upper)
    n = W * H
    prim = torch.empty(n * 6, dtype=torch.float32, device=dev)
    hits = torch.zeros(n * 10, dtype=torch.int32, device=dev)
    ctx.primary_rays_device(cam, W, H, prim.data_ptr(), stream)
    ctx.trace_device(prim.data_ptr(), n, hits.data_ptr(), True, 0.0035, stream)
    torch.cuda.synchronize()
    h = hits.view(n, 10)
    hit = h[:, 0] != 0
    pos = h[:, 3:6].view(torch.float32)
    nrm = h[:, 6:9].view(torch.float32)
    g = torch.Generator(device=dev); g.manual_seed(1)
    ball = torch.randn(n, 3, device=dev, generator=g)
    ball = ball / ball.norm(dim=1, keepdim=True) * torch.rand(n, 1, device=dev, generator=g).pow(1.0 / 3.0)
    d = nrm + ball
    d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-12)
    bounce = torch.cat([pos + nrm * 0.01, d], dim=1)[hit].contiguous()
    m = bounce.shape[0]
    pix = torch.arange(n, device=dev)[hit]
    tile = (pix // W // 4) * (W // 8) + (pix % W) // 8
    out = torch.zeros(m * 10, dtype=torch.int32, device=dev)
    results = {"bounce_rays": int(m)}
    oc = octant(bounce)
    orders = {
        "pixel order": torch.arange(m, device=dev),
        "by octant (stable)": torch.sort(oc, stable=True).indices,
        "by octant, then 8x4 tile": torch.sort(oc * (1 << 40) + tile, stable=True).indices,
        "shuffled": torch.randperm(m, device=dev, generator=g),
    }
    for thr in (8, 32):
        ctx.set_option("refill_threshold", thr)
        for name, perm in orders.items():
            r = bounce[perm].contiguous()
            ms = timed(ctx, r.view(-1), m, out, flush, stream, 0.0035)
            results["bounce, %s, refill %d" % (name, thr)] = {"ms": round(ms, 4), "grays_per_s": round(m / ms / 1e6, 3)}
    # fully incoherent batch
    k = 8_000_000
    lo, hi = sc.lower.astype(np.float32), sc.upper.astype(np.float32)
    ext = (hi - lo) * 0.1
    rnd = torch.empty(k * 6, dtype=torch.float32, device=dev)
    ctx.random_rays_device(7, lo - ext, hi + ext, k, rnd.data_ptr(), stream)
    torch.cuda.synchronize()
    rr = rnd.view(k, 6)
    out2 = torch.zeros(k * 10, dtype=torch.int32, device=dev)
    ctx.set_option("refill_threshold", 8)
    for name, perm in {"as generated": torch.arange(k, device=dev), "by octant": torch.sort(octant(rr), stable=True).indices}.items():
        r = rr[perm].contiguous()
        ms = timed(ctx, r.view(-1), k, out2, flush, stream, -1.0)
        results["random 8M, %s, refill 8" % name] = {"ms": round(ms, 4), "grays_per_s": round(k / ms / 1e6, 3)}
    print(json.dumps(results, indent=1))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(results) + "\n")


if __name__ == "__main__":
    main()

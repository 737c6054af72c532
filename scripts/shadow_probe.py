"""How should the path tracer's shadow rays be batched?
At depth 0 the sun ray of a pixel is the same for every sample of the wave (same primary hit, fixed sun direction); deeper, sun rays
share a direction but start from scattered bounce hits; sky rays are diffuse everywhere. Cases, 1080p terrain, LOD 0.0035, 8 samples:
  depth 0: [sun, sky] interleaved per path (what the wavefront tracer does) | sky only | sun once per pixel (row-major / 8x4 tiles)
  depth 1: [sun, sky] interleaved | all sun then all sky | each alone
    python scripts/shadow_probe.py [--out gpurun_out/shadow_probe.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

W, H, S = 1920, 1080, 8
MF = 0.0035


def timed(ctx, rays, hits, flush, stream, steps=6):
    rays = rays.contiguous()
    n = rays.shape[0]
    for _ in range(2):
        ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), False, MF, stream)
    torch.cuda.synchronize()
    t = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.trace_device(rays.data_ptr(), n, hits.data_ptr(), False, MF, stream); b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    ms = float(np.median(t))
    return {"rays": int(n), "ms": round(ms, 4), "grays_per_s": round(n / ms / 1e6, 3)}


def diffuse(nrm, g):
    n = nrm.shape[0]
    ball = torch.randn(n, 3, device=nrm.device, generator=g)
    ball = ball / ball.norm(dim=1, keepdim=True) * torch.rand(n, 1, device=nrm.device, generator=g).pow(1.0 / 3.0)
    d = nrm + ball
    return d / d.norm(dim=1, keepdim=True).clamp_min(1e-12)


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
    ctx.trace_device(prim.data_ptr(), n, hits.data_ptr(), True, MF, stream)
    torch.cuda.synchronize()
    h = hits.view(n, 10)
    hit = h[:, 0] != 0
    pos = h[:, 3:6].view(torch.float32)[hit]
    nrm = h[:, 6:9].view(torch.float32)[hit]
    pix = torch.arange(n, device=dev)[hit]
    L = pos.shape[0]
    g = torch.Generator(device=dev); g.manual_seed(1)
    sun = torch.tensor([1.0, -2.0, 10.0], device=dev); sun = sun / sun.norm()
    out = torch.zeros(2 * S * L * 10, dtype=torch.int32, device=dev)
    res = {"live_pixels": int(L), "samples": S}

    def shadow_sets(pos, nrm, reps):
        o = (pos + nrm * 0.001).repeat(reps, 1)
        sun_rays = torch.cat([o, sun.expand(o.shape[0], 3)], dim=1)
        sky_rays = torch.cat([o, diffuse(nrm.repeat(reps, 1), g)], dim=1)
        return sun_rays, sky_rays

    for thr in (8, 16, 32):
        ctx.set_option("refill_threshold", thr)
        sun0, sky0 = shadow_sets(pos, nrm, S)
        if thr == 8:
            res["d0 interleaved [sun, sky] x %d samples, refill 8" % S] = timed(ctx, torch.stack([sun0, sky0], dim=1).view(-1, 6), out, flush, stream)
            res["d0 sky only x %d samples, refill 8" % S] = timed(ctx, sky0, out, flush, stream)
            res["d0 sun x %d samples (duplicates), refill 8" % S] = timed(ctx, sun0, out, flush, stream)
        res["d0 sun once per pixel, row-major, refill %d" % thr] = timed(ctx, sun0[:L], out, flush, stream)
        tile = (pix // W // 4) * (W // 8) * 32 + ((pix % W) // 8) * 32 + (pix // W % 4) * 8 + pix % 8
        res["d0 sun once per pixel, 8x4 tiles, refill %d" % thr] = timed(ctx, sun0[:L][torch.sort(tile).indices], out, flush, stream)

    # depth 1: bounce rays of all S samples, their hits, then the shadow rays from those
    ctx.set_option("refill_threshold", 8)
    bounce = torch.cat([(pos + nrm * 0.01).repeat(S, 1), diffuse(nrm.repeat(S, 1), g)], dim=1).contiguous()
    m = bounce.shape[0]
    bh = torch.zeros(m * 10, dtype=torch.int32, device=dev)
    ctx.trace_device(bounce.data_ptr(), m, bh.data_ptr(), True, MF, stream)
    torch.cuda.synchronize()
    b = bh.view(m, 10)
    bhit = b[:, 0] != 0
    pos1 = b[:, 3:6].view(torch.float32)[bhit]; nrm1 = b[:, 6:9].view(torch.float32)[bhit]
    res["d1 live paths"] = int(pos1.shape[0])
    sun1, sky1 = shadow_sets(pos1, nrm1, 1)
    for thr in (8, 16, 32):
        ctx.set_option("refill_threshold", thr)
        if thr == 8:
            res["d1 interleaved [sun, sky], refill 8"] = timed(ctx, torch.stack([sun1, sky1], dim=1).view(-1, 6), out, flush, stream)
            res["d1 all sun then all sky, refill 8"] = timed(ctx, torch.cat([sun1, sky1], dim=0), out, flush, stream)
            res["d1 sky only, refill 8"] = timed(ctx, sky1, out, flush, stream)
        res["d1 sun only, refill %d" % thr] = timed(ctx, sun1, out, flush, stream)
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()

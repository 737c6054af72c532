#!/usr/bin/env python
"""ONE-hop parity at full size, counted: the CUDA path against the UNMODIFIED reference (oracle/_ref, Cubiquity::intersectVolume)
on the four BASELINE scene families at their BASELINE sizes -- a 1080p frame of primary rays, random rays through the dilated
bounds, and the tests' mixed set (in-bounds origins, axis-aligned and exact-diagonal corner cases) -- with LOD off and with the
path tracer's LOD. Every field of every record is compared bit for bit; rays the port abandons (the reference itself would not
return for them) are left out of the reference run and must be reported abandoned by the CUDA path.
    python scripts/parity_census.py [--out gpurun_out/parity_census.json] [--rays 2000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from conftest import mixed_rays  # noqa: E402
from cubiquity_b200 import api  # noqa: E402
from cubiquity_b200 import rays as R  # noqa: E402
from oracle import pyoracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--rays", type=int, default=2_000_000)
ap.add_argument("--scenes", nargs="+", default=["terrain:12", "sphere_noise:10", "soup:14", "city:16"])
args = ap.parse_args()
threads = os.cpu_count() or 1
port, ref = pyoracle.Port(), pyoracle.Ref()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
W, H = 1920, 1080
total = {"rays": 0, "differing_records": 0, "abandoned_by_both": 0, "abandoned_mismatch": 0}
lines = []
for spec in args.scenes:
    kind, log2 = spec.split(":")
    t0 = time.perf_counter()
    sc = api.Scene(kind, int(log2), 1)
    ctx = api.Context(0)
    ctx.upload(sc.nodes, sc.root, sc.colours)
    sd = port.find_subdags(sc.nodes, sc.root)
    vol = ref.volume().load_arrays(sc.nodes, sc.root)
    cam, _, _ = bench.orbit_camera(api, sc, 0)
    d = torch.empty(W * H * 6, dtype=torch.float32, device=dev)
    ctx.primary_rays_device(cam, W, H, d.data_ptr(), stream)
    torch.cuda.synchronize()
    primary = d.cpu().numpy().view(pyoracle.RAY_DTYPE).reshape(-1)
    sets = {"1080p primary rays": primary,
            "random rays, bounds dilated by 0.1": R.random_rays(args.rays, sc.lower, sc.upper, 11),
            "mixed: in-bounds origins, axis-aligned, exact diagonals": mixed_rays(sc.lower, sc.upper, args.rays // 2, 12)}
    for what, rays in sets.items():
        for mf in (-1.0, 0.0035):
            got = ctx.intersect_volume(rays, True, mf)
            mine, _, _ = port.trace(sc.nodes, sd, rays, True, mf, threads=threads)
            keep = mine["pad"] == 0
            want, secs = vol.intersect(np.ascontiguousarray(rays[keep]), True, mf, threads=threads)
            a = np.ascontiguousarray(got[keep]).view(np.uint32).reshape(-1, 10)
            b = want.view(np.uint32).reshape(-1, 10)
            differing = int((a != b).any(axis=1).sum())
            mismatch = int((got["status"][~keep] != 1).sum())
            line = {"scene": "%s 2^%s" % (kind, log2), "nodes": int(len(sc.nodes)), "rays": what, "max_footprint": mf, "count": int(len(rays)),
                    "hits": int((got["hit"] != 0).sum()), "differing_records": differing, "abandoned_by_both": int((~keep).sum()) - mismatch,
                    "abandoned_mismatch": mismatch, "reference_seconds": round(secs, 3)}
            lines.append(line)
            print(json.dumps(line), flush=True)
            total["rays"] += line["count"]; total["differing_records"] += differing
            total["abandoned_by_both"] += line["abandoned_by_both"]; total["abandoned_mismatch"] += mismatch
    del ctx, vol
print(json.dumps({"total": total}), flush=True)
if args.out:
    with open(args.out, "w") as f:
        json.dump({"total": total, "cases": lines}, f, indent=1)

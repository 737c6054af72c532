"""Time cbq_build_dense_device (dense grid -> canonical DAG on the GPU) and, at a size the reference can do in
seconds, the reference's own route (Volume::setVoxel per voxel + Volume::bake, oracle/_ref, one thread).

    python scripts/build_bench.py [--out gpurun_out/build.jsonl]

The grid is generated on the device with torch (a ball with a rough surface and material strata), so the timed
region is the build alone: octree levels + merge + sub-DAGs, input already in HBM."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def grid_on_device(torch, k, dev):
    """Written slab by slab: at 4096^3 the grid itself is 68.7 GB, there is no room for float coordinates of all of it."""
    side = 1 << k
    g = torch.empty((side, side, side), dtype=torch.uint8, device=dev)
    ax = torch.arange(side, device=dev, dtype=torch.float32) - side / 2 + 0.5
    layers = max(1, min(side, (1 << 25) // (side * side)))
    for z0 in range(0, side, layers):
        z, y, x = torch.meshgrid(ax[z0:z0 + layers], ax, ax, indexing="ij")
        r = torch.sqrt(x * x + y * y + z * z)
        bump = torch.sin(x * (37.0 / side)) * torch.sin(y * (53.0 / side)) * torch.sin(z * (29.0 / side)) * (0.04 * side)
        mat = (1 + ((z + side / 2) / (side / 8)).to(torch.int32) % 6).to(torch.uint8)
        g[z0:z0 + layers] = torch.where(r < 0.4 * side + bump, mat, torch.zeros_like(mat))
        del x, y, z, r, bump, mat
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=10)
    ap.add_argument("--reference-log2", type=int, default=7)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from cubiquity_b200 import api
    from oracle import pyoracle
    dev = torch.device("cuda", 0)
    ctx = api.Context(0)
    for k in range(6, args.max_log2 + 1):
        g = grid_on_device(torch, k, dev)
        torch.cuda.synchronize()
        side = 1 << k
        origin = (-side // 2,) * 3
        times = []
        for rep in range(4 if k <= 10 else 2):
            t0 = time.perf_counter()
            count, root = ctx.build_dense(None, origin, device_ptr=g.data_ptr(), size_log2=k)
            times.append(time.perf_counter() - t0)
        s = float(np.median(times[1:]))
        line = {"grid": "%d^3" % side, "voxels": side ** 3, "nodes_out": count, "gpu_build_ms": s * 1e3, "gpu_build_ms_all": [round(t * 1e3, 3) for t in times],
                "gpu_Gvoxels_per_s": side ** 3 / s / 1e9,
                "how": "one piece" if k <= 10 else "bricks of 512^3, merged one by one, one last merge"}
        if k == args.reference_log2:
            host = g.cpu().numpy()
            z, y, x = np.nonzero(host)
            xyzm = np.column_stack([x + origin[0], y + origin[1], z + origin[2], host[z, y, x]]).astype(np.int32)
            ref = pyoracle.Ref()
            v = ref.volume()
            t0 = time.perf_counter()
            v.set_voxels(xyzm)
            t1 = time.perf_counter()
            v.bake()
            t2 = time.perf_counter()
            line.update({"reference_setvoxel_ms": (t1 - t0) * 1e3, "reference_bake_ms": (t2 - t1) * 1e3, "reference_nodes_out": int(len(v.nodes())),
                         "same_node_count": bool(len(v.nodes()) == count), "speedup": (t2 - t0) / s})
        del g
        torch.cuda.empty_cache()
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()

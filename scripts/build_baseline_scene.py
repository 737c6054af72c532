#!/usr/bin/env python
"""BASELINE configs[1]'s volume (4096^3 procedural terrain) built ON THE DEVICE from its 68.7 G voxels: the voxels are evaluated with
torch integer arithmetic (tests/terrain_voxels.py, voxel for voxel the scene library's Terrain::voxel), cbq_build_dense_device merges
them brick by brick, and the result is compared with what the host scene builder makes: node count, and every field of every hit of
a 1080p frame of primary rays traced against both.
    python scripts/build_baseline_scene.py [--log2 12] [--out gpurun_out/baseline_scene.json]"""
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
import terrain_voxels  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=12)
ap.add_argument("--out", default="")
args = ap.parse_args()
k, n = args.log2, 1 << args.log2
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
t0 = time.perf_counter()
host_scene = api.Scene("terrain", k, 1)
t_host = time.perf_counter() - t0
grid = torch.empty((n, n, n), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
terrain_voxels.fill(grid, k, 1)
torch.cuda.synchronize()
t_voxels = time.perf_counter() - t0
built = api.Context(0)
times = []
for _ in range(2):
    t0 = time.perf_counter()
    count, root = built.build_dense(None, (-n // 2,) * 3, device_ptr=grid.data_ptr(), size_log2=k, colours=host_scene.colours)
    times.append(time.perf_counter() - t0)
del grid
torch.cuda.empty_cache()
ref = api.Context(0)
ref.upload(host_scene.nodes, host_scene.root, host_scene.colours)
W, H = 1920, 1080
cam, _, _ = bench.orbit_camera(api, host_scene, 0)
rays = torch.empty(W * H * 6, dtype=torch.float32, device=dev)
built.primary_rays_tiled_device(cam, W, H, rays.data_ptr(), None, stream)
a = torch.zeros(W * H * 10, dtype=torch.int32, device=dev)
b = torch.zeros(W * H * 10, dtype=torch.int32, device=dev)
built.trace_device(rays.data_ptr(), W * H, a.data_ptr(), True, -1.0, stream)
ref.trace_device(rays.data_ptr(), W * H, b.data_ptr(), True, -1.0, stream)
torch.cuda.synchronize()
line = {"scene": "terrain 2^%d seed 1" % k, "voxels": n ** 3, "host_scene_builder_s": t_host, "voxels_on_device_s": t_voxels, "build_dense_s": min(times),
        "build_dense_s_all": times, "Gvoxels_per_s": n ** 3 / min(times) / 1e9, "nodes_device_build": count, "nodes_host_builder": int(len(host_scene.nodes)),
        "same_node_count": bool(count == len(host_scene.nodes)), "rays_compared": W * H, "hits": int((a.view(-1, 10)[:, 0] != 0).sum().item()),
        "records_differing": int((a.view(-1, 10) != b.view(-1, 10)).any(dim=1).sum().item())}
print(json.dumps(line))
if args.out:
    with open(args.out, "w") as f:
        f.write(json.dumps(line) + "\n")

#!/usr/bin/env python
"""What the host links allow when N processes copy at once: every rank moves `--mb` of pinned host memory to its GPU and 1/3 of
that back (the e2e job's 24 : 8 byte ratio), both directions at the same time, all ranks together. One JSON line from rank 0.
    torchrun --nproc-per-node N scripts/host_link_probe.py [--mb 400]"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=400)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_in, n_out = args.mb << 20, (args.mb << 20) // 3
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev); d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def once():
    with torch.cuda.stream(s_in):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s_out):
        h_out.copy_(d_out, non_blocking=True)
    s_in.synchronize(); s_out.synchronize()


once()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.reps):
    once()
secs = (time.perf_counter() - t0) / args.reps
if world > 1:
    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
if rank == 0:
    print(json.dumps({"ranks": world, "mb_in_per_rank": args.mb, "mb_out_per_rank": args.mb // 3, "seconds": secs,
                      "aggregate_h2d_gbs": world * n_in / secs / 1e9, "aggregate_d2h_gbs": world * n_out / secs / 1e9,
                      "equivalent_grays": world * n_in / 24 / secs / 1e9}), flush=True)
if world > 1:
    dist.destroy_process_group()

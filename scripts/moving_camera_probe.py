"""Does the cost-feedback ticket order survive a moving camera? 120 frames of an orbit around the 4096^3 terrain
(cbq_raycast_frame_device, 1080p, L2 flushed between frames), the camera turning by `--step` degrees per frame, with
adaptive_order on and off. With the option on, frame k is dealt in the order learnt from frames k-1 ... k-4."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from cubiquity_b200 import api  # noqa: E402

W, H = 1920, 1080
PI_F = float(np.float32(3.14159265358979))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    sc = api.Scene("terrain", 12, 1)
    ctx = api.Context(0)
    ctx.upload(sc.nodes, sc.root, sc.colours)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    hits = torch.zeros(W * H * 10, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lower, upper = np.asarray(sc.lower, dtype=np.float64), np.asarray(sc.upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    half = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    results = {}
    for step_deg in (0.0, 0.25, 1.0, 4.0):
        for ao, refresh in ((1, 4), (0, 4)):
            ctx.set_option("adaptive_order", ao)
            ctx.set_option("order_refresh", refresh)
            t = []
            for k in range(-8, 120):
                yaw = np.radians(step_deg) * k
                pos = [centre[0] - half * np.sin(yaw), centre[1] - half * np.cos(yaw), centre[2] + half]
                cam = api.camera_from_pose(pos, -(PI_F / 4.0), yaw)
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ctx.raycast_frame_device(cam, W, H, hits.data_ptr(), True, -1.0, stream); b.record()
                torch.cuda.synchronize()
                if k >= 0:
                    t.append(a.elapsed_time(b))
            results["%g deg/frame, adaptive_order=%d, refresh every %d" % (step_deg, ao, refresh)] = {"ms": round(float(np.mean(t)), 4), "grays_per_s": round(W * H / float(np.mean(t)) / 1e6, 3)}
    print(json.dumps(results, indent=1))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(results) + "\n")


if __name__ == "__main__":
    main()

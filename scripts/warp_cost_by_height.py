"""Which tree levels do the WARP INSTRUCTIONS of the primary-ray kernel go to? (no GPU) The per-level event counts of
scripts/event_heights.py weigh every event equally, but a warp pays per section executed, not per lane: at the top of
the tree the 32 rays of a tile walk the same cells together. Here the kernel's schedule (one event per live lane per
step, 8x4-pixel tiles, no mid-flight refill) is replayed on event strings that carry heights, and each executed
section's price is shared out over the heights of the lanes in it."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from warp_sim import COST  # noqa: E402

KINDS = "DAPHO"


def main():
    from cubiquity_b200 import api
    from oracle import pyoracle
    import bench
    port = pyoracle.Port()
    scene = api.Scene("terrain", 12, 1)
    class B: pass
    b = B(); b.lower, b.upper = scene.lower, scene.upper
    cam, pos, yaw = bench.orbit_camera(api, b, 0)
    W, H = 1920, 1080
    rays = port.camera_rays(cam, W, H).reshape(H, W)
    t = rays.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(len(t), size=2048, replace=False))
    sel = np.ascontiguousarray(t[pick].reshape(-1))
    sub = port.find_subdags(scene.nodes, scene.root)
    cap = 512
    ev = np.zeros((sel.size, cap), dtype=np.uint8)
    counts = np.zeros(sel.size, dtype=np.uint32)
    port.lib.cbqo_trace_events_with_heights.restype = None
    port.lib.cbqo_trace_events_with_heights(pyoracle._ptr(scene.nodes), pyoracle._ptr(sub), pyoracle._ptr(sel), ctypes.c_uint64(sel.size), 1,
                                            ctypes.c_float(-1.0), pyoracle._ptr(ev), ctypes.c_uint32(cap), pyoracle._ptr(counts))
    by_height = np.zeros(34)
    overhead = 0.0
    lanes_by_height = np.zeros((34, 2))
    for w in range(len(pick)):
        strings = [ev[w * 32 + l, :counts[w * 32 + l]] for l in range(32)]
        steps = max(len(s) for s in strings)
        overhead += COST["refill"]
        for k in range(steps):
            now = [s[k] for s in strings if k < len(s)]
            kinds = {}
            for e in now:
                kinds.setdefault(int(e) & 7, []).append(int(e) >> 3)
            overhead += COST["loop"]
            shared = 0.0
            if any(kd in kinds for kd in (0, 1, 2, 3)):
                shared += COST["common"]
            if 0 in kinds or 3 in kinds:
                shared += COST["occupied"]
            esvo = [h for kd, hs in kinds.items() if kd != 4 for h in hs]
            for h in esvo:
                by_height[h] += shared / len(esvo)
            for kd, hs in kinds.items():
                price = COST[KINDS[kd]]
                for h in hs:
                    by_height[h] += price / len(hs)
                    lanes_by_height[h][0] += 1.0 / len(hs)     # section executions, shared out
                    lanes_by_height[h][1] += 1.0               # lane events
    n = sel.size
    total = by_height.sum() + overhead
    print("modelled warp instructions per ray %.1f (loop control and refill %.1f)" % (total / n, overhead / n))
    cum = 0.0
    for h in range(13, 0, -1):
        if by_height[h] == 0:
            continue
        cum += by_height[h]
        print("  node height %2d (children %4d voxels wide): %5.1f warp instr / ray  %4.1f %%  cumulative from the top %5.1f %%   lanes per section %.1f" % (
            h, 1 << (h - 1), by_height[h] / n, 100 * by_height[h] / total, 100 * cum / total, lanes_by_height[h][1] / max(lanes_by_height[h][0], 1e-9)))


if __name__ == "__main__":
    main()

"""Time one radius-30 sphere-brush edit three ways on the same volume:
  device   cbq_fill_sphere (edit_kernels.cu), nothing crosses PCIe
  host     cbq_editable (csrc/edit.cpp) + cbq_update of the dirty tail
  ref      the reference's checkpoint() + fillBrush() (oracle/_ref), edit only -- it has no device to sync

    python scripts/edit_bench.py [--scene city --log2 16] [--out gpurun_out/edit.jsonl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="city")
    ap.add_argument("--log2", type=int, default=16)
    ap.add_argument("--edits", type=int, default=30)
    ap.add_argument("--radius", type=float, default=30.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from cubiquity_b200 import api
    from oracle import pyoracle
    sc = api.Scene(args.scene, args.log2, 1)
    rng = np.random.default_rng(0)
    if args.scene == "city":
        centres = np.column_stack([rng.uniform(-1200, 1200, args.edits), rng.uniform(-1200, 1200, args.edits), rng.uniform(0, 400, args.edits)])
    else:
        mid = (sc.lower + sc.upper) / 2.0
        centres = mid + rng.uniform(-0.25, 0.25, (args.edits, 3)) * (sc.upper - sc.lower)
    dev_ctx, host_ctx = api.Context(0), api.Context(0)
    dev_ctx.upload(sc.nodes, sc.root)
    ed = api.Editable(sc.nodes, sc.root)
    ed.sync(host_ctx, True) if hasattr(ed, "sync") else host_ctx.upload(sc.nodes, sc.root)
    v = pyoracle.Ref().volume().load_arrays(sc.nodes, sc.root)
    t_dev, t_host_edit, t_host_sync, t_ref, tail = [], [], [], [], []
    synced = len(sc.nodes)
    for c in centres:
        t0 = time.perf_counter()
        dev_ctx.fill_sphere(c[0], c[1], c[2], args.radius, 0)
        t_dev.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ed.checkpoint(); ed.fill_sphere(c[0], c[1], c[2], args.radius, 0)
        t1 = time.perf_counter()
        nodes = ed.nodes()
        host_ctx.update(nodes, synced, ed.root())
        t_host_sync.append(time.perf_counter() - t1)
        t_host_edit.append(t1 - t0)
        tail.append((len(nodes) - synced) * 32)
        synced = ed.shared_end()
        t0 = time.perf_counter()
        v.checkpoint(); v.fill_sphere(c[0], c[1], c[2], args.radius, 0)
        t_ref.append(time.perf_counter() - t0)
    med = lambda a: round(float(np.median(a[1:])) * 1e3, 4)
    line = {"scene": "%s 2^%d" % (args.scene, args.log2), "radius": args.radius, "edits": args.edits, "device_fill_ms": med(t_dev),
            "host_edit_ms": med(t_host_edit), "host_delta_upload_ms": med(t_host_sync), "reference_edit_ms": med(t_ref),
            "dirty_tail_bytes_median": int(np.median(tail)), "device_nodes_after": dev_ctx.node_count(), "host_nodes_after": int(len(ed.nodes()))}
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()

"""Time cbq_bake (device hash-cons merge) against the reference's Volume::bake on the same array.

    python scripts/bake_bench.py [--scene terrain --log2 12 --edits 200] [--out gpurun_out/bake.jsonl]

The array is the scene builder's output plus `edits` sphere-brush edits (checkpoint + fill, the reference's runtime
edit path), i.e. a canonical DAG with an un-merged copy-on-write tail and undo history -- what a viewer session
hands to bake. The reference leg runs oracle/_ref (the unmodified reference, one thread: it is single-threaded)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="terrain")
    ap.add_argument("--log2", type=int, default=12)
    ap.add_argument("--edits", type=int, default=200)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from cubiquity_b200 import api
    from oracle import pyoracle

    sc = api.Scene(args.scene, args.log2, 1)
    ed = api.Editable(sc.nodes, sc.root)
    rng = np.random.default_rng(0)
    centre = (sc.lower + sc.upper) / 2.0
    span = (sc.upper - sc.lower) * 0.25
    for k in range(args.edits):
        ed.checkpoint()
        c = centre + rng.uniform(-1, 1, 3) * span
        ed.fill_sphere(c[0], c[1], c[2], 30.0, 0 if k % 2 else 2)
    nodes, root = ed.nodes().copy(), ed.root()
    ctx = api.Context(0)
    times = []
    for r in range(args.reps + 1):
        ctx.upload(nodes, root)
        ctx.synchronize()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        count, new_root = ctx.bake()
        times.append(time.perf_counter() - t0)
    reachable = ctx.counter("bake_reachable")
    gpu_s = float(np.median(times[1:] if len(times) > 1 else times))
    line = {"scene": "%s 2^%d + %d brush edits" % (args.scene, args.log2, args.edits), "nodes_in": int(len(nodes)), "reachable": int(reachable),
            "nodes_out": int(count), "gpu_bake_ms": gpu_s * 1e3, "gpu_bake_ms_all": [round(t * 1e3, 3) for t in times],
            "gpu_Mnodes_per_s": len(nodes) / gpu_s / 1e6}
    if not args.no_reference:
        ref = pyoracle.Ref()
        v = ref.volume().load_arrays(nodes, root)
        t0 = time.perf_counter()
        v.bake()
        ref_s = time.perf_counter() - t0
        line.update({"reference_bake_ms": ref_s * 1e3, "reference_nodes_out": int(len(v.nodes())), "speedup": ref_s / gpu_s,
                     "same_node_count": bool(len(v.nodes()) == count)})
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()

"""Offline model of the ray-cast kernel on SECONDARY rays (no GPU needed): why do they run at 9-11 active lanes per
instruction, and which refill policy would change that?

Diffuse bounce rays of a band of the bench frame (origin = primary hit + 0.01 n, direction = normalize(n + point in the
unit ball), LOD 0.0035, pixel order -- what the wavefront path tracer feeds the kernel) are turned into event strings by
the instrumented oracle (O sub-DAG entry, D descend, A advance, P advance+pop, H hit) and fed to a model of one warp of
the persistent kernel: a queue of rays, one event per live lane per step, a refill vote every 8 steps. Section prices
are the SASS counts of profiles/r01_analysis.md.

    python scripts/warp_sim_bounce.py [--rows 96]"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

COST = dict(loop=11, common=17, occupied=9, D=49, A=37, P=48, H=30, O=56, refill=65)


def bounce_events(args):
    from cubiquity_b200 import api
    from oracle import pyoracle
    import bench
    port = pyoracle.Port()
    scene = api.Scene("terrain", args.log2, 1)
    class B: pass
    b = B(); b.lower, b.upper = scene.lower, scene.upper
    cam, pos, yaw = bench.orbit_camera(api, b, 0)
    W, H = 1920, 1080
    rays = port.camera_rays(cam, W, H).reshape(H, W)
    y0 = H // 2 - args.rows // 2
    band = np.ascontiguousarray(rays[y0:y0 + args.rows].reshape(-1))
    sub = port.find_subdags(scene.nodes, scene.root)
    hits, _, _ = port.trace(scene.nodes, sub, band, True, 0.0035, threads=os.cpu_count() or 1)
    hit = hits["hit"] != 0
    rng = np.random.default_rng(1)
    n = int(hit.sum())
    ball = rng.normal(size=(n, 3))
    ball = ball / np.linalg.norm(ball, axis=1, keepdims=True) * rng.random((n, 1)) ** (1.0 / 3.0)
    nrm = hits["normal"][hit].astype(np.float64)
    d = nrm + ball
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    out = np.zeros(n, dtype=band.dtype)
    out["o"] = (hits["position"][hit].astype(np.float64) + nrm * 0.01).astype(np.float32)
    out["d"] = d.astype(np.float32)
    cap = 1024
    ev = np.zeros((n, cap), dtype=np.uint8)
    counts = np.zeros(n, dtype=np.uint32)
    port.lib.cbqo_trace_events.restype = None
    port.lib.cbqo_trace_events(pyoracle._ptr(scene.nodes), pyoracle._ptr(sub), pyoracle._ptr(out), ctypes.c_uint64(n), 1, ctypes.c_float(0.0035),
                               pyoracle._ptr(ev), ctypes.c_uint32(cap), pyoracle._ptr(counts))
    return [bytes(ev[i, :counts[i]]).decode() for i in range(n)]


def step_cost(kinds):
    c = COST["loop"]
    if kinds & set("DAPH"):
        c += COST["common"]
    if kinds & set("DH"):
        c += COST["occupied"]
    return c + sum(COST[k] for k in kinds)


def simulate(rays, threshold, start_together=False, vote_every=8):
    """One warp draining `rays` in order. Returns (warp instructions, {section: [executions, lane-events]})."""
    queue = list(rays)
    lane = [None] * 32         # (string, position)
    total = 0
    use = {k: [0, 0] for k in "DAPHO"}
    fresh = set()
    steps_since_vote = vote_every
    while True:
        idle = [l for l in range(32) if lane[l] is None]
        if steps_since_vote >= vote_every:
            steps_since_vote = 0
            if queue and idle and (len(idle) >= threshold or len(idle) == 32):
                total += COST["refill"]
                for l in idle:
                    if not queue:
                        break
                    lane[l] = [queue.pop(0), 0]
                    fresh.add(l)
        live = [l for l in range(32) if lane[l] is not None]
        if not live:
            if not queue:
                break
            steps_since_vote = vote_every
            continue
        # start_together: while freshly started lanes are still in their opening run of sub-DAG entries and descents,
        # only they step (the others wait), so that the run is walked in lock step
        stepping = live
        if start_together:
            opening = [l for l in fresh if lane[l] is not None and lane[l][0][lane[l][1]] in "OD"]
            fresh.intersection_update(opening)
            if opening:
                stepping = opening
        kinds = set()
        for l in stepping:
            s, p = lane[l]
            k = s[p]
            kinds.add(k)
            use[k][1] += 1
        for k in kinds:
            use[k][0] += 1
        total += step_cost(kinds)
        for l in stepping:
            lane[l][1] += 1
            if lane[l][1] >= len(lane[l][0]):
                lane[l] = None
                fresh.discard(l)
        steps_since_vote += 1
    return total, use


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=12)
    ap.add_argument("--rows", type=int, default=96)
    ap.add_argument("--warps", type=int, default=48)
    args = ap.parse_args()
    rays = [r for r in bounce_events(args) if r]
    n = len(rays)
    ev = sum(len(r) for r in rays)
    opening = np.mean([len(r) - len(r.lstrip("OD")) for r in rays])
    print("bounce rays %d, events/ray %.1f (D %.0f%% A %.0f%% P %.0f%%), opening run of O/D events %.1f" % (
        n, ev / n, 100 * sum(r.count("D") for r in rays) / ev, 100 * sum(r.count("A") for r in rays) / ev, 100 * sum(r.count("P") for r in rays) / ev, opening))
    per = n // args.warps
    chunks = [rays[i * per:(i + 1) * per] for i in range(args.warps)]       # each model warp drains a contiguous run of the queue
    for name, kw in [("refill at 8 idle lanes (the kernel's default)", dict(threshold=8)), ("refill at 4", dict(threshold=4)), ("refill at 16", dict(threshold=16)),
                     ("whole warp at a time (32)", dict(threshold=32)), ("refill at 8, new lanes walk their opening run alone", dict(threshold=8, start_together=True)),
                     ("refill at 16, opening run alone", dict(threshold=16, start_together=True))]:
        tot = 0
        use = {k: [0, 0] for k in "DAPHO"}
        for c in chunks:
            t, u = simulate(c, **kw)
            tot += t
            for k in use:
                use[k][0] += u[k][0]; use[k][1] += u[k][1]
        lanes = {k: (use[k][1] / use[k][0] if use[k][0] else 0) for k in use}
        print("%-52s %6.1f warp instr / ray   lanes per section  D %.1f  A %.1f  P %.1f" % (name, tot / (per * args.warps), lanes["D"], lanes["A"], lanes["P"]))


if __name__ == "__main__":
    main()

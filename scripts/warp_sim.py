"""Offline warp-scheduling model of the ray-cast kernel (no GPU needed).

Takes the per-ray event strings of the instrumented oracle (O sub-DAG entry, D descend, A advance, P advance+pop,
H hit) for the bench frame, groups 32 consecutive rays into a warp exactly as the kernel's ticket queue does, and
prices scheduling policies with the per-section SASS instruction counts of profiles/r01_analysis.md. Used to decide
whether a restructuring is worth GPU time; results are quoted in profiles/r01_analysis.md.

    python scripts/warp_sim.py [--log2 12] [--width 1920 --height 1080] [--sample 4096 warps] [--order tile|row]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

COST = dict(loop=11, common=17, occupied=9, D=49, A=37, P=48, H=30, O=56, refill=65)


def frame_events(args):
    import ctypes
    from cubiquity_b200 import api
    from oracle import pyoracle
    import bench
    port = pyoracle.Port()
    scene = api.Scene("terrain", args.log2, 1)
    class B: pass
    b = B(); b.lower, b.upper = scene.lower, scene.upper
    cam, pos, yaw = bench.orbit_camera(api, b, 0)
    rays = port.camera_rays(cam, args.width, args.height).reshape(args.height, args.width)
    if args.order == "tile":
        t = rays.reshape(args.height // 4, 4, args.width // 8, 8).transpose(0, 2, 1, 3).reshape(-1)
    else:
        t = rays.reshape(-1)
    rng = np.random.default_rng(0)
    nwarps = t.size // 32
    pick = np.sort(rng.choice(nwarps, size=min(args.sample, nwarps), replace=False))
    sel = np.ascontiguousarray(t.reshape(nwarps, 32)[pick].reshape(-1))
    sub = port.find_subdags(scene.nodes, scene.root)
    cap = 512
    ev = np.zeros((sel.size, cap), dtype=np.uint8)
    counts = np.zeros(sel.size, dtype=np.uint32)
    port.lib.cbqo_trace_events.restype = None
    port.lib.cbqo_trace_events(pyoracle._ptr(scene.nodes), pyoracle._ptr(sub), pyoracle._ptr(sel), ctypes.c_uint64(sel.size),
                               1, ctypes.c_float(-1.0), pyoracle._ptr(ev), ctypes.c_uint32(cap), pyoracle._ptr(counts))
    return [[bytes(ev[w * 32 + l, :counts[w * 32 + l]]).decode() for l in range(32)] for w in range(pick.size)]


def esvo_cost(kinds):
    c = COST["loop"]
    if kinds & set("DAPH"):
        c += COST["common"]
    if kinds & set("DH"):
        c += COST["occupied"]
    for k in kinds:
        c += COST[k]
    return c


def policy_current(warp):
    """One event per live lane per step; the step costs every section some lane needs."""
    pos = [0] * 32
    cost = COST["refill"]
    while True:
        kinds = {warp[l][pos[l]] for l in range(32) if pos[l] < len(warp[l])}
        if not kinds:
            return cost
        cost += esvo_cost(kinds)
        for l in range(32):
            if pos[l] < len(warp[l]):
                pos[l] += 1


def policy_majority(warp, order="DAPHO"):
    """Each step runs ONE section: the one most lanes are waiting for (ties by `order`)."""
    pos = [0] * 32
    cost = COST["refill"]
    while True:
        want = {}
        for l in range(32):
            if pos[l] < len(warp[l]):
                want.setdefault(warp[l][pos[l]], []).append(l)
        if not want:
            return cost
        k = max(want, key=lambda x: (len(want[x]), -order.index(x)))
        cost += esvo_cost({k})
        for l in want[k]:
            pos[l] += 1


def policy_sections(warp):
    """Descend section then advance section per trip (V3)."""
    pos = [0] * 32
    cost = COST["refill"]
    while True:
        live = [l for l in range(32) if pos[l] < len(warp[l])]
        if not live:
            return cost
        cost += COST["loop"]
        first = {warp[l][pos[l]] for l in live} & set("DHO")
        if first:
            cost += COST["common"] + COST["occupied"] + sum(COST[k] for k in first)
            for l in live:
                if warp[l][pos[l]] in first:
                    pos[l] += 1
        live = [l for l in live if pos[l] < len(warp[l])]
        second = {warp[l][pos[l]] for l in live} & set("AP")
        if second:
            cost += COST["common"] + sum(COST[k] for k in second)
            for l in live:
                if warp[l][pos[l]] in second:
                    pos[l] += 1


def policy_unified(warp, body=60, pop_extra=35):
    """One branch-free body shared by descend and advance; pop, hit and sub-DAG entry stay branches."""
    pos = [0] * 32
    cost = COST["refill"]
    while True:
        kinds = {warp[l][pos[l]] for l in range(32) if pos[l] < len(warp[l])}
        if not kinds:
            return cost
        cost += COST["loop"]
        if kinds & set("DAPH"):
            cost += COST["common"] + COST["occupied"] + body
        if "P" in kinds:
            cost += pop_extra
        if "H" in kinds:
            cost += COST["H"]
        if "O" in kinds:
            cost += COST["O"]
        for l in range(32):
            if pos[l] < len(warp[l]):
                pos[l] += 1


def bound_packed(warp):
    total = sum(COST[e] + COST["common"] + (COST["occupied"] if e in "DH" else 0) for lane in warp for e in lane)
    return COST["refill"] + total / 32.0


def bound_longest_lane(warp):
    """No schedule can run fewer sections of a kind than the lane that needs most of that kind."""
    c = COST["refill"]
    for k in "DAPHO":
        c += max(lane.count(k) for lane in warp) * esvo_cost({k}) - 0
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=12)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--sample", type=int, default=2048)
    ap.add_argument("--order", default="tile")
    args = ap.parse_args()
    warps = frame_events(args)
    n = len(warps) * 32
    ev = sum(len(l) for w in warps for l in w)
    print("warps %d, events/ray %.1f" % (len(warps), ev / n))
    for name, fn in [("packed bound", bound_packed), ("longest-lane bound", bound_longest_lane), ("current (one event per lane per step)", policy_current),
                     ("majority section per step", policy_majority), ("unified descend/advance body (60 + 35 pop)", policy_unified),
                     ("unified, optimistic body (50 + 30 pop)", lambda w: policy_unified(w, 50, 30)), ("descend section + advance section (V3)", policy_sections)]:
        tot = sum(fn(w) for w in warps)
        print("%-45s %7.1f warp instructions per ray" % (name, tot / n))


if __name__ == "__main__":
    main()

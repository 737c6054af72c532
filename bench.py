#!/usr/bin/env python
"""Benchmark of the SVDAG ray-cast hot path (BASELINE.json metric: Grays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload primary|random|pathtrace]

Default workload = BASELINE.json configs[1]: primary-ray ray cast, 1920x1080, against the procedural
4096^3 multi-material terrain, LOD off, surface properties on. One "step" = one pass of
cbq_trace_device over one frame of rays (2 073 600 rays) that already sit in HBM.

  value      whole-job Grays/s, device-timed (CUDA events on the launching stream, one pair per step,
             L2 flushed between steps), max over ranks.
  e2e        the same metric through the public host-buffer call cbq_trace(): pinned host rays in,
             pinned host hits out, both copies inside the timed region.
  roofline   algorithmic bytes per ray (32 B x node visits V + 24 B ray + 40 B hit; V counted by the
             instrumented oracle on a sample of the same rays) x rays / kernel time, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own intersectVolume (oracle/_ref, compiled from /root/reference) on the
             GPU box's host cores, same rays. oracle/ is used ONLY here and in --impl reference.

N > 1 (torchrun): the DAG is built on rank 0 and replicated with one NCCL broadcast; every rank then
traces its own copy of the frame -- weak scaling, no collective in the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
SCENE_KIND, SCENE_LOG2, SCENE_SEED = "terrain", 12, 1
PI_F = float(np.float32(3.14159265358979))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="primary", choices=["primary", "random", "pathtrace"])
    ap.add_argument("--scene-log2", type=int, default=SCENE_LOG2)
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (reported as such)")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to cbq_set_option")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--random-rays", type=int, default=100_000_000)
    ap.add_argument("--config", type=int, default=0, choices=[0, 4, 5],
                    help="pathtrace presets: 4 = BASELINE configs[3] (16384^3 solids, 1080p, 4 bounces, 64 spp); "
                         "5 = configs[4] (65536^3 city, 3840x2160, 256 spp, a sphere-brush edit + delta re-upload before every frame)")
    ap.add_argument("--scene", default=None, choices=[None, "terrain", "soup", "city", "sphere_noise"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--device-edits", action="store_true", help="pathtrace --edits: apply the brush on every GPU with cbq_fill_sphere instead of on the host + delta upload/broadcast")
    ap.add_argument("--edits", action="store_true", help="pathtrace: carve a radius-30 sphere and delta re-upload before every frame")
    return ap.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        t = json.load(open(path))
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), t.get("source")
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every millisecond from a thread (the timed
    region of the default run is ~10 ms, too short for `nvidia-smi -lms`); nvidia-smi is the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.max_mhz, self.reasons, self.proc, self.nvml = index, [], 0, set(), None, None
        self.stop = threading.Event()

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        while not self.stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.handle))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[0]))
                self.max_mhz = max(self.max_mhz, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                continue

    def __exit__(self, *exc):
        self.stop.set()
        if self.nvml:
            self.thread.join(timeout=1)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def orbit_camera(api, scene, rank):
    """Rank 0 = the viewer's default pose (reference viewer.cpp:71-79); other ranks orbit the scene."""
    lower = np.asarray(scene.lower, dtype=np.float64)
    upper = np.asarray(scene.upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    half_diag = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    yaw = rank * (2.0 * np.pi / 8.0)
    pos = [centre[0] - half_diag * np.sin(yaw), centre[1] - half_diag * np.cos(yaw), centre[2] + half_diag]
    return api.camera_from_pose(pos, -(PI_F / 4.0), yaw), pos, yaw


class OnlyJsonOnStdout:
    """The driver reads ONE JSON line from stdout; libraries (NCCL prints its version banner there) must not add to it.
    File descriptor 1 points at stderr until emit() prints the line."""
    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.saved = None
        print(json.dumps(line), flush=True)


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU intersectVolume, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cubiquity_b200 import api
    from oracle import pyoracle
    scene = api.Scene(SCENE_KIND, args.scene_log2, SCENE_SEED)
    port = pyoracle.Port()
    cam, pos, yaw = orbit_camera(api, scene, 0)
    ocam = port.camera(pos, -(PI_F / 4.0), yaw)
    rays = port.camera_rays(ocam, WIDTH, HEIGHT)
    threads = os.cpu_count() or 1
    kind = "reference"
    try:
        vol = pyoracle.Ref().volume().load_arrays(scene.nodes, scene.root)
        run = lambda r: vol.intersect(r, True, -1.0, threads=threads, want_hits=True)[1]
    except Exception:
        kind = "port"
        sd = port.find_subdags(scene.nodes, scene.root)
        run = lambda r: port.trace(scene.nodes, sd, r, True, -1.0, threads=threads)[1]
    # Bounded sample per step: every 4th 8-row band of the frame (about half a million rays).
    rows = np.arange(HEIGHT).reshape(-1, 8)[::4].reshape(-1)
    sample = np.ascontiguousarray(rays.reshape(HEIGHT, WIDTH)[rows].reshape(-1))
    for _ in range(args.warmup):
        run(sample)
    secs = [run(sample) for _ in range(args.steps)]
    total = float(np.sum(secs))
    value = len(sample) * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": "Grays/s SVDAG traversal (primary rays)", "value": value, "unit": "Grays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": "primary rays 1920x1080 vs procedural 4096^3 terrain SVDAG (BASELINE configs[1]), LOD off, surface properties on",
                   "scene": "%s 2^%d seed %d" % (SCENE_KIND, args.scene_log2, SCENE_SEED), "nodes": int(len(scene.nodes))},
        "cpu_baseline": {"value": value, "unit": "Grays/s", "cores": threads, "kind": kind,
                         "sample": "%d rays per step: every 4th 8-row band of the 1080p frame" % len(sample)},
        "e2e": {"value": value, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def pathtrace_workload(args):
    """--workload pathtrace: tile-sharded 1080p path tracing (BASELINE metric "1080p spp/s", config 4 style:
    recursive bounce loop, `--bounces` bounces, `--spp` samples, LOD 0.0035). Strong scaling: the frame's
    64-row tile bands are dealt round-robin to the ranks, the DAG is replicated, and ONE collective -- a sum of
    the disjoint partial images onto rank 0 -- ends the frame (inside the timed region)."""
    out = OnlyJsonOnStdout()
    import torch
    import torch.distributed as dist
    from cubiquity_b200 import api, sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t_start = time.time()
    if args.config == 4:
        args.scene, args.scene_log2, args.spp, args.bounces = args.scene or "soup", 14, 64, 4
    elif args.config == 5:
        args.scene, args.scene_log2, args.spp, args.bounces, args.width, args.height, args.edits = args.scene or "city", 16, 256, 4, 3840, 2160, True
    kind, log2 = (args.scene or "terrain", args.scene_log2)
    W, H = args.width, args.height
    t_build = time.time()
    scene = api.Scene(kind, log2, SCENE_SEED) if rank == 0 else None
    t_build = time.time() - t_build
    if world > 1:
        nodes, root = sharding.broadcast_volume(dist, scene.nodes if rank == 0 else None, scene.root if rank == 0 else None, device=dev)
        meta = torch.zeros(6, dtype=torch.int64, device=dev)
        col = torch.zeros(256, 3, device=dev)
        if rank == 0:
            meta[:3] = torch.from_numpy(scene.lower.astype(np.int64)); meta[3:] = torch.from_numpy(scene.upper.astype(np.int64))
            col = torch.from_numpy(scene.colours.copy()).to(dev)
        dist.broadcast(meta, src=0); dist.broadcast(col, src=0)
        lower, upper, colours = meta[:3].cpu().numpy(), meta[3:].cpu().numpy(), col.cpu().numpy()
    else:
        nodes, root, lower, upper, colours = scene.nodes, scene.root, scene.lower, scene.upper, scene.colours
    def log(msg):
        if rank == 0:
            sys.stderr.write("[bench pathtrace %.1fs] %s\n" % (time.time() - t_start, msg)); sys.stderr.flush()
    log("scene %s 2^%d: %d nodes, built in %.1f s" % (kind, log2, len(nodes), t_build))
    ctx = api.Context(local)
    for kv in args.option:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.upload(nodes, root, colours)
    if kind == "city":
        cam = api.camera_from_pose([0.0, -2600.0, 1800.0], -0.6, 0.0)      # over the roofs, looking down the avenues
    else:
        cam = api.default_camera(lower, upper)
    stream = torch.cuda.current_stream().cuda_stream
    accum = torch.zeros(H, W, 3, dtype=torch.float32, device=dev)
    # Runtime edits (config 5): rank 0 owns the editable copy (csrc/edit.cpp restates the reference's checkpoint +
    # sphere brush); every frame it carves a radius-30 sphere, ships the dirty tail to the other ranks, and every rank
    # applies it with cbq_update -- the delta, not the DAG, crosses PCIe and NVLink.
    host_edits = args.edits and not args.device_edits
    editable = api.Editable(nodes, root) if (host_edits and rank == 0) else None
    replica = np.array(nodes, dtype=np.uint32, copy=True) if (host_edits and rank != 0) else None
    synced = len(nodes)
    edit_stats = {"tail_bytes": [], "edit_ms": [], "sync_ms": []}

    def edit(frame):
        nonlocal replica, synced
        if not args.edits:
            return
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.device_edits:
            # Every rank replays the same stroke on its own replica (cbq_fill_sphere): nothing crosses PCIe or NVLink.
            rng = np.random.default_rng(frame)
            before = ctx.node_count()
            ctx.fill_sphere(float(rng.uniform(-1200, 1200)), float(rng.uniform(-1200, 1200)), float(rng.uniform(0, 400)), 30.0, 0)
            edit_stats["tail_bytes"].append((ctx.node_count() - before) * 32)
            edit_stats["edit_ms"].append(1e3 * (time.perf_counter() - t0))
            edit_stats["sync_ms"].append(0.0)
            return
        if rank == 0:
            editable.checkpoint()
            rng = np.random.default_rng(frame)
            editable.fill_sphere(float(rng.uniform(-1200, 1200)), float(rng.uniform(-1200, 1200)), float(rng.uniform(0, 400)), 30.0, 0)
            cur, cur_root, dirty = editable.nodes(), editable.root(), synced
        t1 = time.perf_counter()
        if world > 1:
            if rank == 0:
                fresh, new_root, dirty = sharding.broadcast_tail(dist, cur, dirty, cur_root, device=dev)
            else:
                fresh, new_root, dirty = sharding.broadcast_tail(dist, replica, 0, 0, device=dev)
                replica = fresh
        else:
            fresh, new_root = cur, cur_root
        ctx.update(fresh, dirty, new_root)
        edit_stats["tail_bytes"].append((len(fresh) - dirty) * 32)
        synced = editable.shared_end() if rank == 0 else len(fresh)
        if world > 1:
            t = torch.tensor([synced], dtype=torch.int64, device=dev)
            dist.broadcast(t, src=0)
            synced = int(t.item())
        edit_stats["edit_ms"].append(1e3 * (t1 - t0))
        edit_stats["sync_ms"].append(1e3 * (time.perf_counter() - t1))
    # this rank's share = every world-th 64-row band, rendered by ONE call (cbq_pt_params.band_count/index)
    bands = (world, rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(frame):
        edit(frame)
        accum.zero_()
        p = api.pt_params(W, H, spp=args.spp, bounces=args.bounces, variant=api.VARIANT_RECURSIVE,
                          frame_id=frame * args.spp, bands=bands)
        ctx.render_device(cam, p, accum.data_ptr(), stream)
        if world > 1:
            sharding.reduce_image(dist, accum, dst=0)

    log("uploaded; warming up")
    for i in range(max(args.warmup, 3)):
        step(i)
        torch.cuda.synchronize()
        log("warm-up frame %d done" % i)
    if world > 1:
        dist.barrier()
    ctx.reset_counters()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        for i in range(args.steps):
            flush.zero_()
            starts[i].record(); step(100 + i); stops[i].record()
        torch.cuda.synchronize()
    launches = ctx.counter("kernel_launches")
    total_ms = float(np.sum([a.elapsed_time(b) for a, b in zip(starts, stops)]))
    log("timed region done: %.1f ms per frame, %d abandoned rays" % (total_ms / args.steps, ctx.counter("abandoned_rays")))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
    value = W * H * args.spp / (ms * 1e-3)
    # end to end: the host-image call (image up, render, image down) on this rank's bands
    host = api.PinnedArray(H * W * 3, np.float32)
    img = host.array.reshape(H, W, 3)
    img[:] = 0
    t0 = time.perf_counter()
    ctx.render(cam, api.pt_params(W, H, spp=args.spp, bounces=args.bounces, variant=api.VARIANT_RECURSIVE, bands=bands), img)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        mean = float(accum.mean().item()) / args.spp
        line = {"metric": "%dp path-traced spp/s" % H, "value": value, "unit": "spp/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32+i32", "data": "synthetic",
                "config": {"workload": "path tracing %dx%d, %d spp, %d bounces (traceSingleRayRecurse), sun+sky+noise, maxFootprint 0.0035, procedural %s 2^%d SVDAG%s" % (W, H, args.spp, args.bounces, kind, log2, (", radius-30 sphere edit on the device before every frame" if args.device_edits else ", radius-30 sphere edit + delta re-upload before every frame") if args.edits else ""),
                           "dag_mb": round(len(nodes) * 32 / 1e6, 1), "scene_build_s": round(t_build, 2),
                           "nodes": int(len(nodes)), "l2": "flushed between steps", "options": args.option,
                           "parallelism": "replicated DAG, 64-row tile bands round-robin over %d GPU(s) (one render call per GPU), one NCCL reduce per frame" % world},
                "e2e": {"value": W * H * args.spp / e2e_s, "unit": "spp/s", "h2d_bytes_per_step": H * W * 12,
                        "d2h_bytes_per_step": H * W * 12, "call": "cbq_render (host image in, this rank's bands rendered, host image out)"},
                "gpu_launches": int(launches), "clocks": clocks.summary(), "extra": {"mean_radiance": mean, "edits": {k: [round(float(x), 3) for x in v] for k, v in edit_stats.items()} if args.edits else None}}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload == "pathtrace":
        pathtrace_workload(args)
        return

    out = OnlyJsonOnStdout()
    import torch
    import torch.distributed as dist
    from cubiquity_b200 import api, sharding
    from cubiquity_b200 import rays as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cubiquity_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- scene: built on rank 0, replicated by broadcast -------------------------------------
    t0 = time.time()
    scene = None
    if rank == 0:
        scene = api.Scene(SCENE_KIND, args.scene_log2, SCENE_SEED)
        nodes, root, colours = scene.nodes, scene.root, scene.colours
        lower, upper = scene.lower.copy(), scene.upper.copy()
    build_s = time.time() - t0
    if world > 1:
        nodes, root = sharding.broadcast_volume(dist, nodes if rank == 0 else None, root if rank == 0 else None, device=dev)
        meta = torch.zeros(6, dtype=torch.int64, device=dev)
        if rank == 0:
            meta[:3] = torch.from_numpy(lower.astype(np.int64))
            meta[3:] = torch.from_numpy(upper.astype(np.int64))
        dist.broadcast(meta, src=0)
        lower, upper = meta[:3].cpu().numpy(), meta[3:].cpu().numpy()
        col = torch.from_numpy(colours.copy()).to(dev) if rank == 0 else torch.empty(256, 3, device=dev)
        dist.broadcast(col, src=0)
        colours = col.cpu().numpy()

    class Bounds:
        pass
    b = Bounds()
    b.lower, b.upper = lower, upper

    ctx = api.Context(local)
    options = list(args.option)
    if args.workload == "primary" and not any(o.startswith("refill_threshold=") for o in options):
        # Primary rays are coherent: mid-flight lane refill costs 7-10 % on them (profiles/r01_sweeps.md), so
        # this workload runs with the knob at 32 (= refill only when the whole warp is done). The library
        # default (8) is the robust choice for batches of unknown coherence.
        options.append("refill_threshold=32")
    for kv in options:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.upload(nodes, root, colours)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- inputs resident in HBM --------------------------------------------------------------
    # Weak scaling means the SAME work per GPU at every N: every rank casts the frame of BASELINE configs[1] (rank 0's pose).
    # (Per-rank poses on an orbit made the 8-GPU figure the slowest view's, not a scaling measurement: 38.2 vs 8 x 5.6 Grays/s.)
    cam, pos, yaw = orbit_camera(api, b, 0)
    if args.workload == "random":
        n_rays = args.random_rays
        ext = (np.asarray(upper, dtype=np.float64) - np.asarray(lower, dtype=np.float64)) * 0.1
        d_rays = torch.empty(n_rays * 6, dtype=torch.float32, device=dev)
        ctx.random_rays_device(100 + rank, (lower - ext).astype(np.float32), (upper + ext).astype(np.float32), n_rays, d_rays.data_ptr(), stream)
        workload = "incoherent batch of %d random rays (origin U(bounds+10%%), direction uniform on the sphere, device generated) vs procedural 4096^3 terrain SVDAG (BASELINE configs[2])" % n_rays
    else:
        n_rays = WIDTH * HEIGHT
        d_rays = torch.empty(n_rays * 6, dtype=torch.float32, device=dev)
        # Morton-like ray order (north star): ray i is pixel i % 32 of 8x4-pixel tile i / 32, so the 32 rays a warp
        # claims are one compact tile; hits come back in the same order. +7 % over row-major (profiles/r01_analysis.md).
        ctx.primary_rays_tiled_device(cam, WIDTH, HEIGHT, d_rays.data_ptr(), None, stream)
        workload = "primary rays 1920x1080 (8x4-pixel tile order) vs procedural 4096^3 terrain SVDAG (BASELINE configs[1]), LOD off, surface properties on"
    d_hits = torch.zeros(n_rays * 10, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2
    torch.cuda.synchronize()

    def step():
        ctx.trace_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr(), True, -1.0, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, one CUDA-event pair per step on the launching stream ----
    ctx.reset_counters()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        wall0 = time.time()
        for i in range(args.steps):
            if not args.no_flush:
                flush.zero_()                      # evict the DAG, rays and hits from L2 (untimed)
            starts[i].record()
            step()
            stops[i].record()
        torch.cuda.synchronize()
        wall = time.time() - wall0
    launches = ctx.counter("kernel_launches")
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = float(np.sum(step_ms))
    if world > 1:
        dist.barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_rays / (ms_per_step * 1e-3) / 1e9

    # warm-L2 figure for context (same kernel, no flush)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1) / args.steps

    # the same steps with the cost-feedback ticket order switched off (every launch deals tickets in buffer order)
    order_off_ms = None
    if args.workload == "primary" and ctx.get_option("adaptive_order"):
        ctx.set_option("adaptive_order", 0)
        off = []
        for _ in range(min(args.steps, 10)):
            if not args.no_flush:
                flush.zero_()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b2.record()
            torch.cuda.synchronize()
            off.append(a.elapsed_time(b2))
        order_off_ms = float(np.mean(off))
        ctx.set_option("adaptive_order", 1)

    # the same frame with the reference path tracer's LOD threshold (maxFootprint 0.0035, pathtracing_demo.h:84); hits land in a scratch buffer
    lod_ms = None
    if args.workload == "primary":
        d_lod = torch.empty_like(d_hits)
        def lod_step():
            ctx.trace_device(d_rays.data_ptr(), n_rays, d_lod.data_ptr(), True, 0.0035, stream)
        for _ in range(3):
            lod_step()
        torch.cuda.synchronize()
        lod = []
        for _ in range(min(args.steps, 10)):
            if not args.no_flush:
                flush.zero_()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lod_step(); b2.record()
            torch.cuda.synchronize()
            lod.append(a.elapsed_time(b2))
        lod_ms = float(np.mean(lod))
        del d_lod
        for _ in range(2):
            step()                 # hand the cost feedback back to the LOD-off frame before anything else is measured
        torch.cuda.synchronize()

    # ---- secondary metric of BASELINE.json: 1080p path-traced samples per second (spp/s) ------------
    pt = None
    if args.workload == "primary":
        d_accum = torch.zeros(HEIGHT * WIDTH * 3, dtype=torch.float32, device=dev)
        saved_threshold = ctx.get_option("refill_threshold")
        ctx.set_option("refill_threshold", 8)      # the library default: secondary rays are incoherent
        p = api.pt_params(WIDTH, HEIGHT, spp=args.spp, bounces=args.bounces, variant=api.VARIANT_RECURSIVE, frame_id=0)
        ctx.render_device(cam, p, d_accum.data_ptr(), stream)     # warm-up with the same shape (buffers get sized here)
        torch.cuda.synchronize()
        pt_times = []
        for _ in range(3):
            d_accum.zero_()
            flush.zero_()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.render_device(cam, p, d_accum.data_ptr(), stream)
            b2.record()
            torch.cuda.synchronize()
            pt_times.append(a.elapsed_time(b2))
        pt_ms = float(np.median(pt_times))
        ctx.set_option("refill_threshold", saved_threshold)
        if world > 1:
            t = torch.tensor([pt_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pt_ms = float(t.item())
        pt = {"spp_per_s": world * WIDTH * HEIGHT * args.spp / (pt_ms * 1e-3), "ms": pt_ms,
              "config": "1920x1080, %d spp, %d bounces, traceSingleRayRecurse, sun+sky+noise, maxFootprint 0.0035, same terrain" % (args.spp, args.bounces),
              "mean_radiance": float(d_accum.mean().item()) / args.spp}

    # ---- end to end through the public host-buffer call ------------------------------------------
    n_e2e = min(n_rays, 8_000_000)      # the end-to-end leg moves at most 8 M rays (192 MB in, 320 MB out) per call
    pin_rays = api.PinnedArray(n_e2e, api.RAY_DTYPE)
    pin_hits = api.PinnedArray(n_e2e, api.HIT_DTYPE)
    pin_rays.array[:] = d_rays[: n_e2e * 6].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
    for _ in range(2):
        ctx.intersect_volume(pin_rays.array, True, -1.0, out=pin_hits.array)
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.intersect_volume(pin_rays.array, True, -1.0, out=pin_hits.array)   # returns after the D2H copy
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_e2e / e2e_s / 1e9
    hit_fraction = float((pin_hits.array["hit"] == 1).mean())
    device_hits = d_hits[: n_e2e * 10].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    same = device_hits.tobytes() == pin_hits.array.tobytes()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0: oracle-derived roofline and CPU baseline (the only use of oracle/ in this file) ---
    from oracle import pyoracle
    port = pyoracle.Port()
    host_rays = pin_rays.array.copy()
    sd = port.find_subdags(nodes, root)
    pick = np.arange(0, n_e2e, max(1, n_e2e // 200000))
    _, _, st = port.trace(nodes, sd, host_rays[pick], True, -1.0, threads=os.cpu_count() or 1, want_hits=False, want_stats=True)
    visits = st.node_visits() / len(pick)
    bytes_per_ray = 32.0 * visits + 24.0 + 40.0
    peak, peak_src = peaks()
    achieved = n_rays * bytes_per_ray / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic() if args.workload == "primary" else (None, None)
    # parity spot check of what was just timed
    want, _, _ = port.trace(nodes, sd, host_rays[pick], True, -1.0, threads=os.cpu_count() or 1)
    parity = float((want.view(np.uint32).reshape(-1, 10) == device_hits[pick].view(np.uint32).reshape(-1, 10)).all(axis=1).mean())

    threads = os.cpu_count() or 1
    kind = "reference"
    try:
        vol = pyoracle.Ref().volume().load_arrays(nodes, root)
        run = lambda r, t: vol.intersect(r, True, -1.0, threads=t, want_hits=True)[1]
    except Exception:
        kind = "port"
        run = lambda r, t: port.trace(nodes, sd, r, True, -1.0, threads=t)[1]
    cpu_sample = host_rays[: min(n_e2e, 1 << 20)]
    run(cpu_sample[:50000], threads)
    cpu_all = len(cpu_sample) / run(cpu_sample, threads) / 1e9
    cpu_one_sample = cpu_sample[:: 8]
    cpu_one = len(cpu_one_sample) / run(cpu_one_sample, 1) / 1e9

    line = {
        "metric": "Grays/s SVDAG traversal (%s rays)" % ("primary" if args.workload == "primary" else "incoherent"), "value": value, "unit": "Grays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": workload, "scene": "%s 2^%d seed %d" % (SCENE_KIND, args.scene_log2, SCENE_SEED),
                   "nodes": int(len(nodes)), "dag_mb": round(len(nodes) * 32 / 1e6, 1), "rays_per_step_per_gpu": n_rays,
                   "l2": "warm (no flush)" if args.no_flush else "flushed between steps (256 MB memset, untimed)",
                   "hit_fraction": round(hit_fraction, 4), "options": options,
                   "parallelism": "replicated DAG, one 1080p frame per GPU (the same view on every GPU)" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "Grays/s", "h2d_bytes_per_step": n_e2e * 24, "d2h_bytes_per_step": n_e2e * 40, "rays_per_step": n_e2e,
                "ms_per_step": 1e3 * e2e_s, "call": "cbq_trace (pinned host rays -> pinned host hits, 3-stage copy/compute pipeline)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": n_rays * bytes_per_ray,
                     "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "node_visits_per_ray": visits,
                     "kernel": "tracePersistent<surface=1, lodOff=1, BufferSource>", "note": "algorithmic bytes = 32 B x V + 24 B ray + 40 B hit; node re-visits are served by L1/L2, the kernel is issue-bound under SIMT divergence, not HBM-bound (DESIGN.md 4.4, profiles/r01_analysis.md)"},
        "cpu_baseline": {"value": cpu_all, "unit": "Grays/s", "cores": threads, "kind": kind,
                         "sample": "first %d rays of the same frame, all host threads" % len(cpu_sample),
                         "single_thread_value": cpu_one, "single_thread_sample": "%d rays (every 8th of the sample)" % len(cpu_one_sample)},
        "clocks": clocks.summary(),
        "extra": {"warm_l2_ms_per_step": warm_ms, "warm_l2_value": world * n_rays / (warm_ms * 1e-3) / 1e9,
                  "step_ms_min": float(np.min(step_ms)), "step_ms_max": float(np.max(step_ms)), "timed_wall_s": wall,
                  "scene_build_s": build_s, "parity_vs_oracle_sample": parity, "e2e_equals_device_path": bool(same),
                  "ticket_order": "cost feedback: from the 2nd launch over the same ray buffer the 32-ray tickets are dealt longest first (adaptive_order=1; one extra 1-block sort kernel per step, inside the timed region)" if order_off_ms is not None else "buffer order",
                  "adaptive_order_off_ms_per_step": order_off_ms,
                  "adaptive_order_off_value": (world * n_rays / (order_off_ms * 1e-3) / 1e9) if order_off_ms else None,
                  "max_footprint_0.0035_ms_per_step": lod_ms,
                  "max_footprint_0.0035_value": (world * n_rays / (lod_ms * 1e-3) / 1e9) if lod_ms else None,
                  "pathtrace": pt},
    }
    out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

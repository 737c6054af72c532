#!/usr/bin/env python
"""Benchmark of the SVDAG ray-cast hot path (BASELINE.json metric: Grays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload primary|random|pathtrace]

Default workload = BASELINE.json configs[1] as a JOB: FRAMES (64) primary-ray frames of 1920x1080 -- the poses of one
orbit round the procedural 4096^3 multi-material terrain, pose 0 being the reference viewer's start pose -- LOD off,
surface properties on. One "step" = one pass of the hot path over the job: every ray of every frame cast once, no
launch sees a ray buffer twice in a row (tickets in buffer order, `adaptive_order` off: the COLD rate), results as
8-byte compact records (cbq_hit_compact: every field of RayVolumeIntersection except position, which is origin + dir *
distance).

  value      whole-job Grays/s, device-timed: CUDA events on the launching stream, one pair per step, L2 flushed between
             steps (the job's rays alone, 3.2 GB, are 25x the L2), max over ranks.
  N > 1      STRONG scaling of the same job: the DAG is replicated (one NCCL broadcast into cbq_upload_device), the frames
             are cut into contiguous 1/N slices, and inside the timed region every rank ships its results to rank 0 over
             NCCL (send/recv per chunk of frames, overlapped with the trace of the next chunk). N = 1 is the same code
             without the exchange.
  e2e        the same job through the public host-buffer call cbq_trace_compact(): pinned host rays in, pinned host
             results out, both copies inside the timed region (24 + 8 bytes per ray); at N > 1 each rank moves its slice.
  roofline   algorithmic bytes per ray (32 B x node visits V + 24 B ray + 8 B result; V counted by the instrumented
             oracle on a sample of the job's rays) x rays / kernel time against the measured HBM copy bandwidth, plus
             the two bounds that actually bind this kernel: L2 sectors and instruction issue (from the committed ncu
             capture, profiles/r02_traffic.json).
  cpu_baseline  the reference's own intersectVolume (oracle/_ref, compiled from /root/reference) on the GPU box's host
             cores, and its single-threaded CPU path tracer on BASELINE configs[0]. oracle/ is used ONLY here and in
             --impl reference.
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
FRAMES = 64
SCENE_KIND, SCENE_LOG2, SCENE_SEED = "terrain", 12, 1
PI_F = float(np.float32(3.14159265358979))
PITCH = -(PI_F / 4.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="primary", choices=["primary", "random", "pathtrace"])
    ap.add_argument("--scene-log2", type=int, default=SCENE_LOG2)
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames (orbit poses) in the job")
    ap.add_argument("--chunk-frames", type=int, default=1, help="frames per launch and per NCCL send when N > 1")
    ap.add_argument("--exchange", default="direct", choices=["direct", "dma", "nccl"],
                    help="N > 1, how results reach rank 0: direct = the ray-cast kernel stores them in rank 0's buffer over NVLink (peer memory), one launch "
                         "per rank; dma = per-chunk peer copies on the copy engines; nccl = per-chunk NCCL send/recv")
    ap.add_argument("--no-exchange", action="store_true", help="diagnostic: N > 1 without shipping the results to rank 0 (the number is then NOT the job's)")
    ap.add_argument("--lanes", type=int, default=2, help="streams the chunk launches alternate between when N > 1")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to cbq_set_option")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--random-rays", type=int, default=100_000_000)
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements (extra, cpu_baseline): for sweeps")
    ap.add_argument("--config", type=int, default=0, choices=[0, 4, 5],
                    help="pathtrace presets: 4 = BASELINE configs[3] (16384^3 solids, 1080p, 4 bounces, 64 spp); "
                         "5 = configs[4] (65536^3 city, 3840x2160, 256 spp, a sphere-brush edit + delta re-upload before every frame)")
    ap.add_argument("--scene", default=None, choices=[None, "terrain", "soup", "city", "sphere_noise"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--pt-split", default="samples", choices=["samples", "bands"],
                    help="pathtrace at N > 1: every rank renders spp / N samples of the whole frame (equal work by construction), "
                         "or every N-th 64-row band with all samples")
    ap.add_argument("--device-edits", action="store_true", help="pathtrace --edits: apply the brush on every GPU with cbq_fill_sphere instead of on the host + delta upload/broadcast")
    ap.add_argument("--edits", action="store_true", help="pathtrace: carve a radius-30 sphere and delta re-upload before every frame")
    return ap.parse_args()


def workload_text(frames):
    return ("%d frames of primary rays 1920x1080 (orbit poses, 8x4-pixel tile order) vs procedural 4096^3 terrain SVDAG (BASELINE configs[1]), "
            "LOD off, surface properties on, every ray cast once per step, 8-byte compact results" % frames)


def profile_counters():
    """Per-launch counters of the dominant kernel from the committed ncu --set full capture (profiles/r02_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every millisecond from a thread; nvidia-smi is
    the fallback."""

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


def orbit_pose(lower, upper, k, frames):
    """Pose k of the orbit: position and yaw. Pose 0 is the viewer's default for a solid object (reference
    viewer.cpp:71-79): centred, back and up by half the bounding-box diagonal, pitch -45 degrees."""
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    half_diag = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    yaw = k * (2.0 * np.pi / frames)
    return [centre[0] - half_diag * np.sin(yaw), centre[1] - half_diag * np.cos(yaw), centre[2] + half_diag], yaw


def orbit_camera(api, bounds, k, frames=8):
    """(camera, position, yaw) of orbit pose k: what the analysis scripts under scripts/ use."""
    pos, yaw = orbit_pose(bounds.lower, bounds.upper, k, frames)
    return api.camera_from_pose(pos, PITCH, yaw), pos, yaw


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


def pin_to_share_of_cpus(world, local):
    """One process per GPU: give each an equal, disjoint share of the CPUs it is allowed on, so that the pinned staging
    buffers a rank allocates (first touch) and the threads that fill them stay together."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cpus) >= world:
            per = len(cpus) // world
            os.sched_setaffinity(0, cpus[local * per:(local + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU intersectVolume on all host threads, on the same job: one full
    1080p frame of it per step (pose = step index). Nothing of the product is imported or mapped here: the scene comes
    from the host-only generator (scenes/), rays and traversal from the compiled reference (oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scenes
    from oracle import pyoracle
    scene = scenes.Scene(SCENE_KIND, args.scene_log2, SCENE_SEED)
    port = pyoracle.Port()
    threads = os.cpu_count() or 1
    kind = "reference"
    try:
        vol = pyoracle.Ref().volume().load_arrays(scene.nodes, scene.root)
        run = lambda r: vol.intersect(r, True, -1.0, threads=threads, want_hits=True)[1]
    except Exception:
        kind = "port"
        sd = port.find_subdags(scene.nodes, scene.root)
        run = lambda r: port.trace(scene.nodes, sd, r, True, -1.0, threads=threads)[1]

    def frame(k):
        pos, yaw = orbit_pose(scene.lower, scene.upper, k % args.frames, args.frames)
        return port.camera_rays(port.camera(pos, PITCH, yaw), WIDTH, HEIGHT)

    for i in range(args.warmup):
        run(frame(args.steps + i))
    secs = [run(frame(i)) for i in range(args.steps)]
    total = float(np.sum(secs))
    value = WIDTH * HEIGHT * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": "Grays/s SVDAG traversal (primary rays)", "value": value, "unit": "Grays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": workload_text(args.frames), "scene": "%s 2^%d seed %d" % (SCENE_KIND, args.scene_log2, SCENE_SEED),
                   "nodes": int(len(scene.nodes))},
        "cpu_baseline": {"value": value, "unit": "Grays/s", "cores": threads, "kind": kind,
                         "sample": "one full 1080p frame of the job per step (2 073 600 rays, pose = step index), Cubiquity::intersectVolume on %d host threads" % threads},
        "e2e": {"value": value, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def init_ranks():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cubiquity_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, world, rank, local, dev


def replicate_scene(api, sharding, dist, ctx, world, rank, dev, kind, log2):
    """Scene built on rank 0; at N > 1 one NCCL broadcast lands it in device memory and cbq_upload_device takes it from
    there. Returns (node_count, root, lower, upper, colours, build_seconds, host nodes on rank 0 else None)."""
    import torch
    t0 = time.time()
    scene = api.Scene(kind, log2, SCENE_SEED) if rank == 0 else None
    build_s = time.time() - t0
    if world == 1:
        ctx.upload(scene.nodes, scene.root, scene.colours)
        return len(scene.nodes), scene.root, scene.lower.copy(), scene.upper.copy(), scene.colours, build_s, scene
    count, root, colours = sharding.replicate_volume(dist, ctx, scene.nodes if rank == 0 else None, scene.root if rank == 0 else None,
                                                     scene.colours if rank == 0 else None, dev)
    meta = torch.zeros(6, dtype=torch.int64, device=dev)
    if rank == 0:
        meta[:3] = torch.from_numpy(scene.lower.astype(np.int64))
        meta[3:] = torch.from_numpy(scene.upper.astype(np.int64))
    dist.broadcast(meta, src=0)
    return count, root, meta[:3].cpu().numpy(), meta[3:].cpu().numpy(), colours, build_s, scene


def pathtrace_workload(args):
    """--workload pathtrace: 1080p path tracing (BASELINE metric "1080p spp/s", config 4 style: recursive bounce loop,
    `--bounces` bounces, `--spp` samples, LOD 0.0035). Strong scaling: the DAG is replicated, the frame's samples (default)
    or its 64-row tile bands are shared out over the ranks, and ONE collective -- a sum of the partial images onto rank 0 --
    ends the frame (inside the timed region)."""
    out = OnlyJsonOnStdout()
    torch, dist, world, rank, local, dev = init_ranks()
    from cubiquity_b200 import api, sharding
    t_start = time.time()
    if args.config == 4:
        args.scene, args.scene_log2, args.spp, args.bounces = args.scene or "soup", 14, 64, 4
    elif args.config == 5:
        args.scene, args.scene_log2, args.spp, args.bounces, args.width, args.height, args.edits = args.scene or "city", 16, 256, 4, 3840, 2160, True
    kind, log2 = (args.scene or "terrain", args.scene_log2)
    W, H = args.width, args.height
    ctx = api.Context(local)
    for kv in args.option:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    count, root, lower, upper, colours, t_build, scene = replicate_scene(api, sharding, dist, ctx, world, rank, dev, kind, log2)

    def log(msg):
        if rank == 0:
            sys.stderr.write("[bench pathtrace %.1fs] %s\n" % (time.time() - t_start, msg)); sys.stderr.flush()
    log("scene %s 2^%d: %d nodes, built in %.1f s" % (kind, log2, count, t_build))
    if kind == "city":
        cam = api.camera_from_pose([0.0, -2600.0, 1800.0], -0.6, 0.0)      # over the roofs, looking down the avenues
    else:
        cam = api.default_camera(lower, upper)
    stream = torch.cuda.current_stream().cuda_stream
    accum = torch.zeros(H, W, 3, dtype=torch.float32, device=dev)
    # Runtime edits (config 5): rank 0 owns the editable copy (csrc/edit.cpp restates the reference's checkpoint +
    # sphere brush); every frame it carves a radius-30 sphere and ships the dirty tail; every rank applies it with
    # cbq_update_device -- the delta, not the DAG, crosses PCIe (once, on rank 0) and NVLink.
    host_edits = args.edits and not args.device_edits
    editable = api.Editable(scene.nodes, scene.root) if (host_edits and rank == 0) else None
    synced = count
    edit_stats = {"tail_bytes": [], "edit_ms": [], "sync_ms": []}

    def edit(frame):
        nonlocal synced
        if not args.edits:
            return
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rng = np.random.default_rng(frame)
        x, y, z = float(rng.uniform(-1200, 1200)), float(rng.uniform(-1200, 1200)), float(rng.uniform(0, 400))
        if args.device_edits:
            # Every rank replays the same stroke on its own replica (cbq_fill_sphere): nothing crosses PCIe or NVLink.
            before = ctx.node_count()
            ctx.fill_sphere(x, y, z, 30.0, 0)
            edit_stats["tail_bytes"].append((ctx.node_count() - before) * 32)
            edit_stats["edit_ms"].append(1e3 * (time.perf_counter() - t0))
            edit_stats["sync_ms"].append(0.0)
            return
        cur = cur_root = None
        if rank == 0:
            editable.checkpoint()
            editable.fill_sphere(x, y, z, 30.0, 0)
            cur, cur_root = editable.nodes(), editable.root()
        t1 = time.perf_counter()
        if world > 1:
            n, _, tail_bytes = sharding.replicate_tail(dist, ctx, cur, synced, cur_root, dev)
        else:
            ctx.update(cur, synced, cur_root)
            n, tail_bytes = len(cur), (len(cur) - synced) * 32
        edit_stats["tail_bytes"].append(tail_bytes)
        nxt = torch.tensor([editable.shared_end() if rank == 0 else 0], dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(nxt, src=0)
        synced = int(nxt.item())
        edit_stats["edit_ms"].append(1e3 * (t1 - t0))
        edit_stats["sync_ms"].append(1e3 * (time.perf_counter() - t1))

    by_samples = args.pt_split == "samples" and world > 1
    s0, s1 = sharding.split_range(args.spp, world, rank) if by_samples else (0, args.spp)
    bands = (world, rank) if (world > 1 and not by_samples) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(frame):
        edit(frame)
        accum.zero_()
        if s1 > s0:
            p = api.pt_params(W, H, spp=s1 - s0, bounces=args.bounces, variant=api.VARIANT_RECURSIVE,
                              frame_id=frame * args.spp + s0, bands=bands)
            ctx.render_device(cam, p, accum.data_ptr(), stream)
        if world > 1:
            sharding.reduce_image(dist, accum, dst=0)

    log("uploaded; warming up")
    for i in range(max(args.warmup, 3)):
        step(i)
        torch.cuda.synchronize()
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
    # end to end: the host-image call (image up, render, image down) on this rank's share
    host = api.PinnedArray(H * W * 3, np.float32)
    img = host.array.reshape(H, W, 3)
    img[:] = 0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if s1 > s0:
        ctx.render(cam, api.pt_params(W, H, spp=s1 - s0, bounces=args.bounces, variant=api.VARIANT_RECURSIVE, frame_id=s0, bands=bands), img)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        mean = float(accum.mean().item()) / args.spp
        split = ("replicated DAG, every GPU renders %d of the %d samples of the whole frame, one NCCL reduce per frame" % (s1 - s0, args.spp)) if by_samples else \
                ("replicated DAG, 64-row tile bands round-robin over %d GPU(s) (one render call per GPU), one NCCL reduce per frame" % world)
        line = {"metric": "%dp path-traced spp/s" % H, "value": value, "unit": "spp/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32+i32", "data": "synthetic",
                "config": {"workload": "path tracing %dx%d, %d spp, %d bounces (traceSingleRayRecurse), sun+sky+noise, maxFootprint 0.0035, procedural %s 2^%d SVDAG%s" % (W, H, args.spp, args.bounces, kind, log2, (", radius-30 sphere edit on the device before every frame" if args.device_edits else ", radius-30 sphere edit + delta re-upload before every frame") if args.edits else ""),
                           "dag_mb": round(count * 32 / 1e6, 1), "scene_build_s": round(t_build, 2),
                           "nodes": int(count), "l2": "flushed between steps", "options": args.option, "parallelism": split},
                "e2e": {"value": W * H * args.spp / e2e_s, "unit": "spp/s", "h2d_bytes_per_step": H * W * 12,
                        "d2h_bytes_per_step": H * W * 12, "call": "cbq_render (host image in, this rank's share rendered, host image out)"},
                "gpu_launches": int(launches), "clocks": clocks.summary(), "extra": {"mean_radiance": mean, "edits": {k: [round(float(x), 3) for x in v] for k, v in edit_stats.items()} if args.edits else None}}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def timed_calls(torch, flush, fn, steps, warm=3):
    """Mean milliseconds of fn() over `steps` calls, one CUDA-event pair each, L2 flushed before each (untimed)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms))


def config1_pathtracer(api, ctx, torch, flush, stream):
    """BASELINE configs[0] as it stands in the reference (pathtracing_demo.cpp:214-229): procedural 256^3 sphere+noise
    volume, 512x512, 1 sample per pixel, traceSingleRay (one diffuse bounce), viewer start pose. The reference's
    single-threaded CPU path tracer (oracle/_ref: the unmodified pathtracing_demo.cpp) is timed on this box next to the
    same frame on the GPU."""
    from oracle import pyoracle
    sc = api.Scene("sphere_noise", 8, 1)
    W = H = 512
    lower, upper = np.asarray(sc.lower, dtype=np.float64), np.asarray(sc.upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    hd = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    pos = [centre[0], centre[1] - hd, centre[2] + hd]
    out = {"config": "256^3 sphere+noise SVDAG (%d nodes), 512x512, 1 spp, traceSingleRay (1 diffuse bounce), sun+sky+noise, maxFootprint 0.0035" % len(sc.nodes)}
    # GPU: same frame, own context so that the bench volume stays uploaded
    small = api.Context(ctx.device)
    small.upload(sc.nodes, sc.root, sc.colours)
    cam = api.camera_from_pose(pos, PITCH, 0.0)
    p = api.pt_params(W, H, spp=1, bounces=1, variant=api.VARIANT_ONE_BOUNCE)
    acc = torch.zeros(H * W * 3, dtype=torch.float32, device="cuda:%d" % ctx.device)
    ms = timed_calls(torch, flush, lambda: (acc.zero_(), small.render_device(cam, p, acc.data_ptr(), stream)), steps=5)
    ours = acc.cpu().numpy().reshape(H, W, 3)
    small.close()
    out["gpu_spp_per_s"] = W * H / (ms * 1e-3)
    out["gpu_ms"] = ms
    try:
        ref = pyoracle.Ref()
        op = pyoracle.PtParams(W, H, 1, 1, 0, 1, 1, 1, 0.0035, 0, 0, 0, W, H, 0)
        as_is, secs = ref.pt_render(sc.nodes, sc.root, sc.colours, pos, PITCH, 0.0, op, reseed=False)      # the reference as shipped
        seeded, secs2 = ref.pt_render(sc.nodes, sc.root, sc.colours, pos, PITCH, 0.0, op, reseed=True)    # same code, per-pixel streams
        mse = float(((as_is - ours) ** 2).mean())
        out.update({"cpu_spp_per_s": W * H / secs, "cpu_seconds": secs, "cores": 1, "kind": "reference",
                    "cpu_what": "PathtracingDemo::traceSingleRay over the frame, unmodified pathtracing_demo.cpp, one thread (its RNG is a process global)",
                    "mean_radiance_gpu": float(ours.mean()), "mean_radiance_reference_as_is": float(as_is.mean()),
                    "mean_radiance_rel_diff": abs(float(ours.mean()) - float(as_is.mean())) / float(as_is.mean()),
                    "psnr_vs_as_is_db": float(10 * np.log10(1.0 / mse)) if mse > 0 else None,
                    "max_abs_diff_vs_reference_with_per_pixel_streams": float(np.abs(seeded - ours).max()),
                    "speedup": (W * H / (ms * 1e-3)) / (W * H / secs)})
    except Exception as e:                                    # oracle/_ref absent: report the GPU side alone
        out["cpu_unavailable"] = repr(e)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload == "pathtrace":
        pathtrace_workload(args)
        return

    out = OnlyJsonOnStdout()
    torch, dist, world, rank, local, dev = init_ranks()
    from cubiquity_b200 import api, sharding
    host_cpus = pin_to_share_of_cpus(world, local)

    ctx = api.Context(local)
    for kv in args.option:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    count, root, lower, upper, colours, build_s, scene = replicate_scene(api, sharding, dist, ctx, world, rank, dev, SCENE_KIND, args.scene_log2)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.set_option("adaptive_order", 0)           # the headline is the cold rate: no launch learns from an earlier one
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2
    frame_rays = WIDTH * HEIGHT

    # ---- inputs resident in HBM: this rank's contiguous slice of the job ------------------------
    if args.workload == "random":
        total_rays = args.random_rays
        r0, r1 = sharding.split_range(total_rays, world, rank)
        n_mine = r1 - r0
        d_rays = torch.empty(max(n_mine, 1) * 6, dtype=torch.float32, device=dev)
        ext = (np.asarray(upper, dtype=np.float64) - np.asarray(lower, dtype=np.float64)) * 0.1
        # every rank generates its own slice of the batch (counter-based generator, one seed per rank)
        ctx.random_rays_device(100 + rank, (lower - ext).astype(np.float32), (upper + ext).astype(np.float32), n_mine, d_rays.data_ptr(), stream)
        plans = [[sharding.split_range(total_rays, world, r)] for r in range(world)]
        chunks = [(0, n_mine)]
        workload = "incoherent batch of %d random rays (origin U(bounds+10%%), direction uniform on the sphere, device generated) vs procedural 4096^3 terrain SVDAG (BASELINE configs[2]), 8-byte compact results" % total_rays
        ctx.set_option("refill_threshold", 8)
    else:
        f0, f1 = sharding.split_range(args.frames, world, rank)
        n_mine = (f1 - f0) * frame_rays
        total_rays = args.frames * frame_rays
        r0 = f0 * frame_rays
        d_rays = torch.empty(max(n_mine, 1) * 6, dtype=torch.float32, device=dev)
        for f in range(f0, f1):
            pos, yaw = orbit_pose(lower, upper, f, args.frames)
            # Morton-like ray order (north star): ray i of a frame is pixel i % 32 of 8x4-pixel tile i / 32, so the 32 rays a
            # warp claims are one compact tile; results come back in the same order.
            ctx.primary_rays_tiled_device(api.camera_from_pose(pos, PITCH, yaw), WIDTH, HEIGHT, d_rays.data_ptr() + (f - f0) * frame_rays * 24, None, stream)
        # N = 1: the whole job is one launch. N > 1: a launch and a message per chunk of frames.
        per_chunk = args.chunk_frames if world > 1 else args.frames
        plans = [sharding.job_chunks(args.frames, frame_rays, world, r, per_chunk) for r in range(world)]
        chunks = [(b - r0, e - r0) for b, e in plans[rank]]
        workload = workload_text(args.frames)
        ctx.set_option("refill_threshold", 32)    # tile-ordered primary rays: a warp takes its next tile when the whole tile is done
    class DevicePointer:
        """A raw device pointer as something torch.as_tensor can view (no copy)."""
        def __init__(self, ptr, words):
            self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}

    # results: rank 0 holds the whole job's, the others their slice. With --exchange direct / dma the job's buffer is ONE
    # allocation on rank 0's GPU that every process addresses over NVLink (cbq_shared_alloc / cbq_shared_open).
    mode = args.exchange if (world > 1 and not args.no_exchange) else "none"
    shared = sharding.SharedResults(dist, ctx, total_rays * 8, dev) if mode in ("direct", "dma") else None
    if shared is not None and rank == 0:
        d_out = torch.as_tensor(DevicePointer(shared.ptr, total_rays * 2), device=dev)
    elif mode == "direct":
        d_out = torch.as_tensor(DevicePointer(shared.ptr + r0 * 8, max(n_mine, 1) * 2), device=dev)     # rank 0's memory, my slice
    else:
        d_out = torch.zeros((total_rays if rank == 0 else max(n_mine, 1)) * 2, dtype=torch.int32, device=dev)
    my_out = d_out[r0 * 2:(r0 + n_mine) * 2] if rank == 0 else d_out
    if mode == "direct":
        chunks = [(0, n_mine)]            # the kernel's stores are the transfer: nothing to pipeline, one launch per rank
    torch.cuda.synchronize()

    # Chunks alternate between two streams: a launch is a persistent grid, so the next chunk's CTAs move in as the
    # previous chunk's drain and the end-of-kernel tail of one launch is filled with the start of the next.
    main_stream = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream(device=dev) for _ in range(args.lanes)] if (len(chunks) > 1 and args.lanes > 1) else [main_stream]
    copy_stream = torch.cuda.Stream(device=dev) if mode == "dma" else None
    done = torch.zeros(1, dtype=torch.int32, device=dev)

    def step():
        """One pass over the job. Rank r traces its slice; the results reach rank 0 inside the step:
          direct  the kernel writes them there (8-byte stores over NVLink), then one tiny all-reduce says "all landed";
          dma     chunk by chunk, a peer copy on the copy engines beside the trace of the next chunk, then the all-reduce;
          nccl    chunk by chunk, an NCCL send beside the trace of the next chunk; rank 0 posts the receives first."""
        works = sharding.post_receives(dist, d_out, plans, 2) if (mode == "nccl" and rank == 0) else []
        for lane in lanes:
            if lane is not main_stream:
                lane.wait_stream(main_stream)
        for i, (b0, b1) in enumerate(chunks):
            if b1 <= b0:
                continue
            lane = lanes[i % len(lanes)]
            with torch.cuda.stream(lane):
                ctx.trace_compact_device(d_rays.data_ptr() + b0 * 24, b1 - b0, my_out.data_ptr() + b0 * 8, True, -1.0, lane.cuda_stream)
                if mode == "nccl" and rank != 0:
                    works += sharding.send_chunk(dist, my_out[b0 * 2:b1 * 2])     # ordered after the kernel on this lane
                elif mode == "dma" and rank != 0:
                    copy_stream.wait_stream(lane)
                    ctx.copy_device(shared.ptr + (r0 + b0) * 8, my_out.data_ptr() + b0 * 8, (b1 - b0) * 8, copy_stream.cuda_stream)
        for lane in lanes:
            if lane is not main_stream:
                main_stream.wait_stream(lane)
        if copy_stream is not None:
            main_stream.wait_stream(copy_stream)
        for w in works:
            w.wait()                      # joins the NCCL stream into the launching stream: the stop event covers the exchange
        if mode in ("direct", "dma"):
            dist.all_reduce(done)         # stream-ordered after every rank's stores / copies: when it completes on rank 0, the job's results are there

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
            flush.zero_()                      # evict the DAG, rays and results from L2 (untimed)
            if world > 1:
                dist.barrier()                 # a step is a job: all ranks start it together (untimed)
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
    value = total_rays / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public host-buffer call: this rank's slice from pinned host memory ----
    pin_rays = api.PinnedArray(max(n_mine, 1), api.RAY_DTYPE)
    pin_out = api.PinnedArray(max(n_mine, 1), api.COMPACT_DTYPE)
    if n_mine:
        pin_rays.array[:] = d_rays[: n_mine * 6].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
    for _ in range(2):
        ctx.intersect_volume_compact(pin_rays.array[:n_mine], True, -1.0, out=pin_out.array[:n_mine])
    e2e_steps = max(3, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.intersect_volume_compact(pin_rays.array[:n_mine], True, -1.0, out=pin_out.array[:n_mine])   # returns after the D2H copy
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_rays / e2e_s / 1e9
    device_compact = my_out[: n_mine * 2].cpu().numpy().view(api.COMPACT_DTYPE).reshape(-1)
    same = device_compact.tobytes() == pin_out.array[:n_mine].tobytes()
    gathered_ok = None
    if world > 1:
        # the exchange delivered what the ranks computed: compare a checksum of rank 0's gathered buffer with the sum of the slices'
        mine = my_out[: n_mine * 2].to(torch.int64).sum().reshape(1)
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        gathered_ok = bool(rank != 0 or int(d_out.to(torch.int64).sum().item()) == int(mine.item())) if not args.no_exchange else None

    if shared is not None:
        del d_out, my_out
        shared.close(dist)
    if rank != 0:
        pin_rays.free(); pin_out.free()
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "Grays/s SVDAG traversal (%s rays)" % ("primary" if args.workload == "primary" else "incoherent"), "value": value, "unit": "Grays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": workload, "scene": "%s 2^%d seed %d" % (SCENE_KIND, args.scene_log2, SCENE_SEED),
                   "nodes": int(count), "dag_mb": round(count * 32 / 1e6, 1), "rays_per_step": int(total_rays),
                   "ticket_order": "buffer order (adaptive_order = 0): cold, nothing learnt from earlier launches",
                   "l2": "flushed between steps (256 MB memset, untimed); the job's rays are %.1f GB" % (total_rays * 24 / 1e9),
                   "options": args.option,
                   "parallelism": ("replicated DAG (NCCL broadcast -> cbq_upload_device), contiguous 1/%d slices of the job per GPU, results gathered in one buffer on rank 0 inside the timed region: %s"
                                   % (world, {"direct": "the ray-cast kernel of every rank stores its 8-byte results straight into that buffer over NVLink (CUDA IPC peer memory), one launch per rank, one 4-byte all-reduce as the completion barrier",
                                              "dma": "peer copies on the copy engines per %d-frame chunk, beside the trace of the next chunk, one 4-byte all-reduce as the completion barrier" % args.chunk_frames,
                                              "nccl": "NCCL send/recv per %d-frame chunk, beside the trace of the next chunk" % args.chunk_frames,
                                              "none": "NOT gathered (diagnostic run)"}[mode])) if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "Grays/s", "h2d_bytes_per_step": int(total_rays) * 24, "d2h_bytes_per_step": int(total_rays) * 8,
                "rays_per_step": int(total_rays), "ms_per_step": 1e3 * e2e_s,
                "call": "cbq_trace_compact (pinned host rays -> pinned host 8-byte results, 3-stage copy/compute pipeline)" + (", each rank its slice" if world > 1 else "")},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "extra": {"step_ms_min": float(np.min(step_ms)), "step_ms_max": float(np.max(step_ms)), "timed_wall_s": wall, "scene_build_s": build_s,
                  "e2e_equals_device_path": bool(same), "gather_checksum_ok": gathered_ok, "host_cpus_per_rank": host_cpus},
    }
    if world == 1 and not args.quick:
        secondary(args, api, ctx, torch, flush, stream, dev, scene, d_rays, n_mine, device_compact, pin_rays, ms_per_step, total_rays, line)
    pin_rays.free(); pin_out.free()
    out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def secondary(args, api, ctx, torch, flush, stream, dev, scene, d_rays, n_rays, device_compact, pin_rays, ms_per_step, total_rays, line):
    """N = 1 only: roofline, parity spot check, CPU baselines (the only use of oracle/ in this arm) and the other
    rates the analysis refers to."""
    from oracle import pyoracle
    port = pyoracle.Port()
    threads = os.cpu_count() or 1
    nodes, root = scene.nodes, scene.root
    sd = port.find_subdags(nodes, root)
    host_rays = pin_rays.array[:n_rays]
    pick = np.arange(0, n_rays, max(1, n_rays // 200000))
    sample = np.ascontiguousarray(host_rays[pick])
    want, _, st = port.trace(nodes, sd, sample, True, -1.0, threads=threads, want_stats=True)
    visits = st.node_visits() / len(pick)
    got = api.expand_hits(sample, device_compact[pick])
    parity = float((want.view(np.uint32).reshape(-1, 10) == got.view(np.uint32).reshape(-1, 10)).all(axis=1).mean())

    frame_rays = WIDTH * HEIGHT
    extra = line["extra"]
    extra["parity_vs_oracle_sample"] = parity
    extra["parity_sample"] = "%d rays spread over the job, expanded to 40-byte records, every field compared bit for bit" % len(pick)
    if args.workload == "primary":
        d_hits = torch.zeros(frame_rays * 10, dtype=torch.int32, device=dev)
        d_full = torch.zeros(n_rays * 10, dtype=torch.int32, device=dev) if n_rays * 40 < (8 << 30) else None
        one = lambda mf=-1.0: ctx.trace_device(d_rays.data_ptr(), frame_rays, d_hits.data_ptr(), True, mf, stream)
        ms1 = timed_calls(torch, flush, one, steps=10)
        extra["single_frame_cold"] = {"value": frame_rays / ms1 / 1e6, "ms": ms1, "what": "pose 0 alone (2 073 600 rays, 40-byte records), buffer order: the end-of-kernel tail is not amortised"}
        ctx.set_option("adaptive_order", 1)
        ms2 = timed_calls(torch, flush, one, steps=10, warm=6)
        extra["single_frame_replayed_order"] = {"value": frame_rays / ms2 / 1e6, "ms": ms2, "what": "same launch repeated, tickets dealt longest first from the previous launch's costs (adaptive_order = 1)"}
        ctx.set_option("adaptive_order", 0)
        ms3 = timed_calls(torch, flush, lambda: one(0.0035), steps=10)
        extra["single_frame_cold_max_footprint_0.0035"] = {"value": frame_rays / ms3 / 1e6, "ms": ms3}
        if d_full is not None:
            ms4 = timed_calls(torch, flush, lambda: ctx.trace_device(d_rays.data_ptr(), n_rays, d_full.data_ptr(), True, -1.0, stream), steps=5)
            extra["job_with_40_byte_records"] = {"value": n_rays / ms4 / 1e6, "ms": ms4}
        del d_hits, d_full
        # BASELINE configs[2]: 100 M incoherent rays
        m = args.random_rays
        rr = torch.empty(m * 6, dtype=torch.float32, device=dev)
        ro = torch.zeros(m * 2, dtype=torch.int32, device=dev)
        ext = (np.asarray(scene.upper, dtype=np.float64) - np.asarray(scene.lower, dtype=np.float64)) * 0.1
        ctx.random_rays_device(100, (scene.lower - ext).astype(np.float32), (scene.upper + ext).astype(np.float32), m, rr.data_ptr(), stream)
        ctx.set_option("refill_threshold", 8)
        ms5 = timed_calls(torch, flush, lambda: ctx.trace_compact_device(rr.data_ptr(), m, ro.data_ptr(), True, -1.0, stream), steps=3, warm=1)
        extra["random_100M"] = {"value": m / ms5 / 1e6, "ms": ms5, "what": "BASELINE configs[2]: %d random rays, refill threshold 8, compact results" % m}
        del rr, ro
        # BASELINE secondary metric: 1080p path-traced samples per second on the same terrain
        acc = torch.zeros(HEIGHT * WIDTH * 3, dtype=torch.float32, device=dev)
        pos, yaw = orbit_pose(scene.lower, scene.upper, 0, args.frames)
        cam = api.camera_from_pose(pos, PITCH, yaw)
        p = api.pt_params(WIDTH, HEIGHT, spp=args.spp, bounces=args.bounces, variant=api.VARIANT_RECURSIVE, frame_id=0)
        ms6 = timed_calls(torch, flush, lambda: (acc.zero_(), ctx.render_device(cam, p, acc.data_ptr(), stream)), steps=3, warm=1)
        extra["pathtrace"] = {"spp_per_s": WIDTH * HEIGHT * args.spp / (ms6 * 1e-3), "ms": ms6,
                              "config": "1920x1080, %d spp, %d bounces, traceSingleRayRecurse, sun+sky+noise, maxFootprint 0.0035, same terrain" % (args.spp, args.bounces),
                              "mean_radiance": float(acc.mean().item()) / args.spp}
        # its roofline: B_spp = sum over the rays of a sample of (32 B x V + 24 B + 40 B) + 12 B accumulate; rays and node
        # visits per sample counted by the instrumented oracle on six 8-row bands of the same frame, 1 sample per pixel
        ocam = port.camera(pos, PITCH, yaw)
        o_rays = o_visits = o_px = 0
        for y0 in range(60, HEIGHT, 180):
            op = pyoracle.PtParams(WIDTH, HEIGHT, 1, args.bounces, 1, 1, 1, 1, 0.0035, 0, 0, y0, WIDTH, y0 + 8, 0)
            _, _, nr = port.render(nodes, sd, scene.colours, ocam, op, threads=threads)
            o_rays += nr; o_visits += port.last_render_stats().node_visits(); o_px += WIDTH * 8
        b_spp = 32.0 * o_visits / o_px + 64.0 * o_rays / o_px + 12.0
        spp_rate = extra["pathtrace"]["spp_per_s"]
        peak_pt, _ = peaks()
        extra["pathtrace"]["roofline"] = {"bound": "hbm", "bytes_per_spp": b_spp, "rays_per_spp": o_rays / o_px, "node_visits_per_spp": o_visits / o_px,
                                          "achieved": b_spp * spp_rate / 1e9, "peak": peak_pt, "unit": "GB/s", "frac": b_spp * spp_rate / 1e9 / peak_pt,
                                          "rays_per_s": o_rays / o_px * spp_rate,
                                          "sample": "%d pixels (six 8-row bands), 1 spp, instrumented oracle" % o_px}
        del acc
        ctx.set_option("refill_threshold", 32)

    bytes_per_ray = 32.0 * visits + 24.0 + 8.0
    peak, peak_src = peaks()
    achieved = total_rays * bytes_per_ray / (ms_per_step * 1e-3) / 1e9
    prof = profile_counters() if args.workload == "primary" else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (prof["dram_bytes_read"] + prof["dram_bytes_write"]) if prof else None,
            "traffic_source": prof.get("source") if prof else None,
            "algorithmic_bytes_per_launch": total_rays * bytes_per_ray, "peak_source": peak_src, "bytes_per_ray": bytes_per_ray,
            "node_visits_per_ray": visits, "kernel": "tracePersistent<surface=1, lodOff=1, BufferSource, CompactSink>",
            "note": "algorithmic bytes = 32 B x V + 24 B ray + 8 B result. Node re-visits are served by L1/L2 (DRAM traffic is a fraction of the algorithmic bytes), "
                    "so HBM does not bind this kernel; the two bounds below do (DESIGN.md 4.4, profiles/r02_analysis.md)"}
    if prof:
        rays_prof = float(prof["rays"])
        rate = total_rays / (ms_per_step * 1e-3)
        # L2: sectors the kernel requests from L2 per ray x 32 B, against the measured L2 bandwidth of the guide
        if prof.get("lts_sectors") and prof.get("l2_peak_gbs"):
            l2_gbs = prof["lts_sectors"] / rays_prof * 32.0 * rate / 1e9
            roof["l2_sector_bound"] = {"achieved": l2_gbs, "peak": prof["l2_peak_gbs"], "unit": "GB/s", "frac": l2_gbs / prof["l2_peak_gbs"],
                                       "sectors_per_ray": prof["lts_sectors"] / rays_prof, "peak_source": prof.get("l2_peak_source")}
        # issue: warp instructions per ray x ray rate, against SMs x 4 schedulers x clock
        if prof.get("warp_instructions"):
            sm_clock = (line["clocks"].get("sm_mhz") or line["clocks"].get("sm_max_mhz") or 1965.0) * 1e6
            issue_peak = 148 * 4 * sm_clock
            ipr = prof["warp_instructions"] / rays_prof
            roof["issue_bound"] = {"achieved": ipr * rate / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s", "frac": ipr * rate / issue_peak,
                                   "warp_instructions_per_ray": ipr, "active_lanes_per_instruction": prof.get("thread_instructions", 0) / prof["warp_instructions"],
                                   "note": "the binding limit: one instruction per scheduler per cycle, executed at about half the lanes under SIMT divergence"}
    line["roofline"] = roof

    kind = "reference"
    try:
        vol = pyoracle.Ref().volume().load_arrays(nodes, root)
        run = lambda r, t: vol.intersect(r, True, -1.0, threads=t, want_hits=True)[1]
    except Exception:
        kind = "port"
        run = lambda r, t: port.trace(nodes, sd, r, True, -1.0, threads=t)[1]
    cpu_sample = np.ascontiguousarray(host_rays[: min(n_rays, 1 << 20)])
    run(cpu_sample[:50000], threads)
    cpu_all = len(cpu_sample) / run(cpu_sample, threads) / 1e9
    cpu_one_sample = np.ascontiguousarray(cpu_sample[:: 8])
    cpu_one = len(cpu_one_sample) / run(cpu_one_sample, 1) / 1e9
    line["cpu_baseline"] = {"value": cpu_all, "unit": "Grays/s", "cores": threads, "kind": kind,
                            "sample": "first %d rays of the job (frame 0), Cubiquity::intersectVolume on all host threads" % len(cpu_sample),
                            "single_thread_value": cpu_one, "single_thread_sample": "%d rays (every 8th of the sample)" % len(cpu_one_sample)}
    if args.workload == "primary":
        line["cpu_baseline"]["pathtrace"] = config1_pathtracer(api, ctx, torch, flush, stream)


if __name__ == "__main__":
    main()

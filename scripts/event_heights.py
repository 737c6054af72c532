"""Where in the tree does a ray spend its events? (no GPU) Events of the instrumented oracle by the height of the node
they happen in, for the bench frame's primary rays and for its diffuse bounce rays. Input to DESIGN.md section 10 (how much
a brick layer over the bottom levels could remove)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def histogram(port, pyoracle, scene, sub, rays, mf):
    hist = np.zeros(5 * 34, dtype=np.uint64)
    port.lib.cbqo_trace_event_heights.restype = None
    port.lib.cbqo_trace_event_heights(pyoracle._ptr(scene.nodes), pyoracle._ptr(sub), pyoracle._ptr(rays), ctypes.c_uint64(len(rays)), 1,
                                      ctypes.c_float(mf), pyoracle._ptr(hist))
    return hist.reshape(5, 34).astype(np.float64)


def report(name, h, n):
    tot = h[:4].sum()
    print("%s: %.1f events per ray" % (name, tot / n))
    cum = 0.0
    for height in range(1, 14):
        row = h[:4, height]
        cum += row.sum()
        print("  node height %2d (children %4d voxels wide): D %5.2f  A %5.2f  P %5.2f  H %5.2f per ray   cumulative %5.1f %%" % (
            height, 1 << (height - 1), row[0] / n, row[1] / n, row[2] / n, row[3] / n, 100 * cum / tot))


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
    band = np.ascontiguousarray(rays[::12].reshape(-1))          # every 12th row: 172 800 rays
    sub = port.find_subdags(scene.nodes, scene.root)
    report("primary rays, LOD off", histogram(port, pyoracle, scene, sub, band, -1.0), len(band))
    report("primary rays, LOD 0.0035", histogram(port, pyoracle, scene, sub, band, 0.0035), len(band))
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
    report("bounce rays, LOD 0.0035", histogram(port, pyoracle, scene, sub, out, 0.0035), n)


if __name__ == "__main__":
    main()

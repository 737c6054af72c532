"""Design study for DESIGN.md section 10 (no GPU): if the bottom levels of the DAG were walked by a voxel DDA ("occupancy
bricks") instead of the reference's descend / advance / pop, would every returned field still be the reference's?
The oracle port has an experimental switch that does exactly that for cubes the ray enters from outside; this script
traces the same rays with the switch off and on and counts the records that differ."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main():
    from cubiquity_b200 import api
    from oracle import pyoracle
    import bench
    from conftest import mixed_rays
    port = pyoracle.Port(experiments=True)        # the design-study build; the parity checker has no such switches
    port.lib.cbqo_experiment_set_brick_height.restype = None
    port.lib.cbqo_experiment_set_grid_height.restype = None
    threads = 1                                   # the switch is a process-wide global: keep the comparison single-threaded per call
    for kind, log2 in (("terrain", 12), ("sphere_noise", 8), ("soup", 9), ("city", 12)):
        sc = api.Scene(kind, log2, 1)
        sub = port.find_subdags(sc.nodes, sc.root)
        pos, yaw = bench.orbit_pose(sc.lower, sc.upper, 0, bench.FRAMES)
        prim = port.camera_rays(port.camera(pos, bench.PITCH, yaw), 1920, 1080).reshape(1080, 1920)[::9].reshape(-1)
        sets = {"primary (every 9th row)": np.ascontiguousarray(prim), "mixed incl. degenerate": mixed_rays(sc.lower, sc.upper, 200000, seed=3)}
        for name, rays in sets.items():
            port.lib.cbqo_experiment_set_brick_height(0)
            want, _, _ = port.trace(sc.nodes, sub, rays, True, -1.0, threads=threads)
            line = "%-13s %-26s %7d rays" % (kind, name, len(rays))
            for h in (2, 3, 4):
                port.lib.cbqo_experiment_set_brick_height(h)
                got, _, _ = port.trace(sc.nodes, sub, rays, True, -1.0, threads=threads)
                port.lib.cbqo_experiment_set_brick_height(0)
                a = got.view(np.uint32).reshape(len(rays), -1)
                w = want.view(np.uint32).reshape(len(rays), -1)
                bad = np.nonzero((a != w).any(axis=1))[0]
                line += "   %d^3 bricks: %d differ" % (1 << h, len(bad))
                if len(bad) and h == 3:
                    i = bad[0]
                    line += " (first: ray %d want %s got %s)" % (i, want[i], got[i])
            for g in (5, 6, 7):
                port.lib.cbqo_experiment_set_grid_height(g)
                got, _, _ = port.trace(sc.nodes, sub, rays, True, -1.0, threads=threads)
                port.lib.cbqo_experiment_set_grid_height(0)
                a = got.view(np.uint32).reshape(len(rays), -1)
                w = want.view(np.uint32).reshape(len(rays), -1)
                line += "   top grid of %d-voxel cells: %d differ" % (1 << g, int((a != w).any(axis=1).sum()))
            print(line, flush=True)


if __name__ == "__main__":
    main()

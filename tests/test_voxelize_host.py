"""cbq_mesh_analyse (csrc/voxelize_host.cpp): Mesh::build's verdict, against the reference's own (voxelization.cpp:765-823)."""
import numpy as np

import meshes


def test_closed_open_and_inside_out_are_told_apart_like_the_reference(api, ref):
    tris, mats = meshes.soup()
    for name, t in (("closed", tris), ("inside out", meshes.flipped(tris)), ("open", tris[200:])):
        info = api.mesh_analyse(t)
        # the reference's verdict comes with a voxelisation: run it on the mesh scaled down tenfold to keep that cheap
        closed, inside_out, _ = ref.volume().voxelize(t * np.float32(0.1), np.ones(len(t), dtype=np.uint8), 1)
        small = api.mesh_analyse(t * np.float32(0.1))
        assert (bool(small.is_closed), bool(small.is_inside_out)) == (closed, inside_out), name
        assert bool(info.is_closed) == (name != "open") and bool(info.is_inside_out) == (name == "inside out"), name
        p = t.reshape(-1, 3)
        assert np.array_equal(np.array(info.lower[:]), p.min(axis=0)) and np.array_equal(np.array(info.upper[:]), p.max(axis=0))

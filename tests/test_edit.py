"""csrc/edit.cpp (host-side COW edits) against the reference's own Volume::checkpoint / fillBrush / undo / redo:
the node ARRAYS must be identical word for word, not merely the voxels."""
import numpy as np
import pytest


def both(api, ref, scenes, kind="sphere_noise", size_log2=6):
    sc = scenes(kind, size_log2)
    return sc, api.Editable(sc.nodes, sc.root), ref.volume().load_arrays(sc.nodes, sc.root)


def same(ed, v):
    return ed.root() == v.root() and ed.shared_end() == v.shared_end() and np.array_equal(ed.nodes(), v.nodes())


def test_edit_sequence_matches_reference_word_for_word(api, ref, scenes):
    sc, ed, v = both(api, ref, scenes)
    assert same(ed, v)
    rng = np.random.default_rng(3)
    for step in range(6):
        ed.checkpoint(); v.checkpoint()
        assert same(ed, v)
        c = rng.uniform(-30, 30, 3)
        r = float(rng.choice([1.0, 2.5, 7.0, 12.0, 30.0]))
        m = int(rng.choice([0, 0, 5, 200]))
        ed.fill_sphere(c[0], c[1], c[2], r, m); v.fill_sphere(c[0], c[1], c[2], r, m)
        assert same(ed, v), "edit %d" % step
    # two brushes in one checkpoint, then history navigation
    ed.fill_sphere(0, 0, 0, 9, 3); v.fill_sphere(0, 0, 0, 9, 3)
    assert same(ed, v)
    for op in ("undo", "undo", "redo", "undo", "undo", "undo", "redo"):
        getattr(ed, op)(); getattr(v, op)()
        assert same(ed, v), op
    ed.checkpoint(); v.checkpoint()          # a checkpoint after undo drops the redo entries
    ed.fill_sphere(5, 5, 5, 4, 0); v.fill_sphere(5, 5, 5, 4, 0)
    assert same(ed, v)
    ed.redo(); v.redo()
    assert same(ed, v)


def test_edit_far_from_the_origin_and_on_an_empty_volume(api, ref):
    from cubiquity_b200.dagfile import material_nodes
    ed, v = api.Editable(material_nodes(), 0), ref.volume()
    ed.checkpoint(); v.checkpoint()
    for c, r, m in [((1000.5, -2000.25, 3000.0), 6.0, 9), ((-(2.0 ** 20), 2.0 ** 20, 17.0), 3.0, 1), ((0.0, 0.0, 0.0), 1.0, 2)]:
        ed.fill_sphere(c[0], c[1], c[2], r, m); v.fill_sphere(c[0], c[1], c[2], r, m)
        assert same(ed, v)
    assert len(ed.nodes()) > 300


@pytest.mark.gpu
def test_config5_loop_without_the_reference_library(gpu, port, api, scenes):
    """Edit -> delta re-upload -> render, entirely with the product library (BASELINE config 5's loop)."""
    from oracle import pyoracle
    sc = api.Scene("city", 14, seed=2)
    ed = api.Editable(sc.nodes, sc.root)
    ed.sync(gpu, first_upload=True, colours=sc.colours)
    cam = api.camera_from_pose([0.0, -2600.0, 1800.0], -0.6, 0.0)
    ocam = port.camera([0.0, -2600.0, 1800.0], -0.6, 0.0)
    for frame, centre in enumerate([(128.0, 128.0, 300.0), (-380.0, 250.0, 40.0), (600.0, -700.0, 5.0)]):
        ed.checkpoint()
        ed.fill_sphere(centre[0], centre[1], centre[2], 30.0, 0)
        before = gpu.counter("bytes_h2d")
        ed.sync(gpu)
        assert gpu.counter("bytes_h2d") - before < 400000
        nodes, root = ed.nodes().copy(), ed.root()
        assert np.array_equal(gpu.download_nodes(), nodes)
        p = api.pt_params(192, 108, spp=2, bounces=2, variant=api.VARIANT_RECURSIVE, frame_id=2 * frame)
        want, _, _ = port.render(nodes, port.find_subdags(nodes, root), sc.colours, ocam, pyoracle.pt_params_from(p), threads=8)
        assert np.array_equal(gpu.render(cam, p), want)

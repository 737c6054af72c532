"""Delta re-upload after runtime edits: cbq_update with the dirty tail only."""
import numpy as np
import pytest

from conftest import assert_hits_identical, mixed_rays

pytestmark = pytest.mark.gpu


def test_edit_protocol_with_reference_edits(gpu, port, ref, scenes):
    """The reference's own edit path (viewer.cpp:152-172): checkpoint -> fillBrush -> re-sync."""
    sc = scenes("sphere_noise", 7)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    gpu.upload(v.nodes(), v.root())
    rays = mixed_rays(sc.lower, sc.upper, 100000, seed=2)
    synced = v.shared_end()
    rng = np.random.default_rng(1)
    for step in range(4):
        v.checkpoint()
        c = rng.uniform(-40, 40, 3)
        v.fill_sphere(c[0], c[1], c[2], 12.0, 0 if step % 2 == 0 else 5)
        nodes, root = v.nodes(), v.root()
        before = gpu.counter("bytes_h2d")
        gpu.update(nodes, synced, root)
        sent = gpu.counter("bytes_h2d") - before
        assert sent == (len(nodes) - synced) * 32 + 512            # the tail + header + 8 sub-DAGs, nothing else
        assert sent < 0.2 * nodes.nbytes
        synced = v.shared_end()
        assert np.array_equal(gpu.download_nodes(), nodes)
        assert gpu.subdags().tobytes() == port.find_subdags(nodes, root).tobytes()
        want, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, True, -1.0, threads=8)
        assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "after edit %d" % step)
    # undo = just a different root over the same array
    v.undo()
    gpu.update(v.nodes(), len(v.nodes()), v.root())
    want, _, _ = port.trace(v.nodes(), port.find_subdags(v.nodes(), v.root()), rays, True, -1.0, threads=8)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "after undo")
    # bake relocates everything: full upload
    v.bake()
    gpu.upload(v.nodes(), v.root())
    want, _, _ = port.trace(v.nodes(), port.find_subdags(v.nodes(), v.root()), rays, True, -1.0, threads=8)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "after bake")


def test_update_grows_past_capacity(gpu, port, scenes):
    sc = scenes("sphere_noise", 6)
    nodes = sc.nodes.copy()
    gpu.upload(nodes, sc.root)
    # Append > capacity worth of (unreachable) nodes plus a new root that is a copy of the old one.
    extra = np.zeros((200000, 8), dtype=np.uint32)
    extra[-1] = nodes[sc.root]
    grown = np.concatenate([nodes, extra])
    gpu.update(grown, len(nodes), len(grown) - 1)
    assert gpu.node_count() == len(grown)
    assert np.array_equal(gpu.download_nodes(0, len(nodes)), nodes)
    rays = mixed_rays(sc.lower, sc.upper, 50000, seed=4)
    want, _, _ = port.trace(nodes, port.find_subdags(nodes, sc.root), rays, True, -1.0)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "grown")


def test_update_argument_checks(gpu, api, scenes):
    sc = scenes("sphere_noise", 6)
    gpu.upload(sc.nodes, sc.root)
    with pytest.raises(api.CubiquityError):
        gpu.update(sc.nodes, len(sc.nodes) + 5, sc.root)
    bad = sc.nodes.copy()
    bad[sc.root] = 0xfffffff0
    with pytest.raises(api.CubiquityError) as e:
        gpu.update(bad, 256, sc.root)
    assert e.value.code == api.ERROR_CORRUPT_VOLUME


def test_city_65536_with_runtime_edits_and_lod(gpu, port, ref, api):
    """BASELINE config 5 in miniature: the 65536^3 city, three reference edits (checkpoint + radius-30
    fillBrush, viewer.cpp:165-168) each shipped as a delta, ray cast with the path tracer's LOD threshold and
    path-traced after every edit."""
    from oracle import pyoracle
    sc = api.Scene("city", 16, seed=2)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    gpu.upload(v.nodes(), v.root(), sc.colours)
    synced = v.shared_end()
    lower, upper = np.array([-1500, -1500, -64]), np.array([1500, 1500, 1100])
    rays = mixed_rays(lower, upper, 120000, seed=8)
    cam = api.camera_from_pose([0.0, -2600.0, 1800.0], -0.6, 0.0)
    ocam = port.camera([0.0, -2600.0, 1800.0], -0.6, 0.0)
    for step, centre in enumerate([(128.0, 128.0, 300.0), (-380.0, 250.0, 40.0), (600.0, -700.0, 5.0)]):
        v.checkpoint()
        v.fill_sphere(centre[0], centre[1], centre[2], 30.0, 0)
        nodes, root = v.nodes(), v.root()
        tail = len(nodes) - synced
        gpu.update(nodes, synced, root)
        synced = v.shared_end()
        assert 0 < tail * 32 < 400000                      # tens of KB per edit, not the whole DAG
        sd = port.find_subdags(nodes, root)
        assert gpu.subdags().tobytes() == sd.tobytes()
        for mf in (-1.0, 0.0035):
            want, _, _ = port.trace(nodes, sd, rays, True, mf, threads=8)
            got = gpu.intersect_volume(rays, True, mf)
            assert_hits_identical(got, want, "city edit %d mf %g" % (step, mf))
            # one hop: the edited volume `v` IS the reference's object; trace the rays the port could finish with it
            ok = want["pad"] == 0
            direct, _ = v.intersect(np.ascontiguousarray(rays[ok]), True, mf, threads=8)
            assert_hits_identical(np.ascontiguousarray(got[ok]), direct, "city edit %d mf %g vs the reference itself" % (step, mf))
        p = api.pt_params(160, 90, spp=1, bounces=2, variant=api.VARIANT_RECURSIVE, frame_id=step)
        op = pyoracle.pt_params_from(p)
        want_img, _, _ = port.render(nodes, sd, sc.colours, ocam, op, threads=8)
        assert np.array_equal(gpu.render(cam, p), want_img)


def test_device_resident_upload_and_tail(gpu, port, ref, api, scenes):
    """cbq_upload_device / cbq_update_device: the node array (or its dirty tail) is already in device memory, as
    after an NCCL broadcast; the result must be the volume cbq_upload / cbq_update would have made."""
    torch = pytest.importorskip("torch")
    sc = scenes("sphere_noise", 7)
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    nodes, root = v.nodes(), v.root()
    d_nodes = torch.from_numpy(nodes.view(np.int32).reshape(-1)).cuda()
    gpu.upload_device(d_nodes.data_ptr(), len(nodes), root, sc.colours)
    assert np.array_equal(gpu.download_nodes(), nodes)
    assert gpu.subdags().tobytes() == port.find_subdags(nodes, root).tobytes()
    rays = mixed_rays(sc.lower, sc.upper, 50000, seed=2)
    sd = port.find_subdags(nodes, root)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), port.trace(nodes, sd, rays, True, -1.0)[0], "device upload")
    # the reference's edit, shipped as a device-resident tail
    synced = len(nodes)
    v.checkpoint()
    v.fill_sphere(0.0, 0.0, float(sc.upper[2]) - 10.0, 14.0, 0)
    nodes, root = v.nodes(), v.root()
    d_tail = torch.from_numpy(nodes[synced:].view(np.int32).reshape(-1).copy()).cuda()
    before = gpu.counter("bytes_h2d")
    gpu.update_device(d_tail.data_ptr(), synced, len(nodes), root)
    assert gpu.counter("bytes_h2d") - before == 512                     # header + sub-DAGs; the tail never touched the host link
    assert np.array_equal(gpu.download_nodes(), nodes)
    sd = port.find_subdags(nodes, root)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), port.trace(nodes, sd, rays, True, -1.0)[0], "device tail")
    # a child index past the end is refused and the previous state survives
    bad = nodes[synced:].copy()
    bad[0, 0] = len(nodes) + 5
    d_bad = torch.from_numpy(bad.view(np.int32).reshape(-1)).cuda()
    with pytest.raises(api.CubiquityError) as e:
        gpu.update_device(d_bad.data_ptr(), synced, len(nodes), root)
    assert e.value.code == api.ERROR_CORRUPT_VOLUME

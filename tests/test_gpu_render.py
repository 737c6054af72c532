"""GPU parity of the path tracer (cbq_render) against the oracle's restatement of the bounce loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PI_F = float(np.float32(3.14159265358979))


def both_cameras(api, port, sc):
    cam = api.default_camera(sc.lower, sc.upper)
    ocam = port.camera(list(cam.position), -(PI_F / 4.0), 0.0)
    assert bytes(cam) == bytes(ocam)
    return cam, ocam


def oracle_params(pyoracle, p):
    return pyoracle.pt_params_from(p)


def reference_image(ref, sc, cam, p):
    """ONE hop: the same frame from the reference's own PathtracingDemo::traceSingleRay / traceSingleRayRecurse
    (unmodified pathtracing_demo.cpp in oracle/_ref, per-pixel RNG streams seeded as cbq_render seeds them)."""
    from oracle import pyoracle
    op = pyoracle.pt_params_from(p)
    img, _ = ref.pt_render(sc.nodes, sc.root, sc.colours, list(cam.position), -(PI_F / 4.0), 0.0, op)
    return img


@pytest.mark.parametrize("kind,size_log2", [("sphere_noise", 8), ("soup", 8)])
@pytest.mark.parametrize("bounces,spp", [(0, 1), (1, 2), (4, 3)])
def test_recursive_variant_is_bit_exact(gpu, port, ref, api, scenes, kind, size_log2, bounces, spp):
    from oracle import pyoracle
    sc = scenes(kind, size_log2)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, ocam = both_cameras(api, port, sc)
    p = api.pt_params(160, 120, spp=spp, bounces=bounces, variant=api.VARIANT_RECURSIVE, frame_id=7)
    sd = port.find_subdags(sc.nodes, sc.root)
    want, _, nrays = port.render(sc.nodes, sd, sc.colours, ocam, oracle_params(pyoracle, p), threads=8)
    got = gpu.render(cam, p)
    assert nrays > 160 * 120 * spp
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "max abs diff %g" % np.abs(got - want).max()
    assert got.std() > 0.05
    direct = reference_image(ref, sc, cam, p)
    assert np.array_equal(got.view(np.uint32), direct.view(np.uint32)), "vs the reference itself: max abs diff %g" % np.abs(got - direct).max()


def test_one_bounce_variant_within_gamma_tolerance(gpu, port, api, scenes):
    """BASELINE config 1: traceSingleRay as used by the reference (pathtracing_demo.cpp:222). Its
    per-sample powf gamma is not bit-portable; everything before the powf is."""
    from oracle import pyoracle
    sc = scenes("sphere_noise", 8)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, ocam = both_cameras(api, port, sc)
    p = api.pt_params(256, 256, spp=2, variant=api.VARIANT_ONE_BOUNCE)
    sd = port.find_subdags(sc.nodes, sc.root)
    want, _, _ = port.render(sc.nodes, sd, sc.colours, ocam, oracle_params(pyoracle, p), threads=8)
    got = gpu.render(cam, p)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
    mse = float(((got - want) ** 2).mean())
    psnr = 10 * np.log10((want.max() ** 2) / max(mse, 1e-30))
    assert psnr > 100.0


def test_flags_tiles_and_accumulation(gpu, port, api, scenes):
    from oracle import pyoracle
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, ocam = both_cameras(api, port, sc)
    sd = port.find_subdags(sc.nodes, sc.root)
    for sun, sky, noise in [(0, 1, 1), (1, 0, 0), (0, 0, 1)]:
        p = api.pt_params(96, 64, spp=1, bounces=2, variant=1, include_sun=sun, include_sky=sky, add_noise=noise, max_footprint=-1.0)
        want, _, _ = port.render(sc.nodes, sd, sc.colours, ocam, oracle_params(pyoracle, p), threads=4)
        assert np.array_equal(gpu.render(cam, p), want)
    # Tiles: four rectangles rendered separately into one image == the whole frame (tile sharding is exact).
    w, h = 100, 70
    whole = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=1, variant=1))
    pieces = np.zeros_like(whole)
    for rect in [(0, 0, 37, 41), (37, 0, 100, 41), (0, 41, 64, 70), (64, 41, 100, 70)]:
        gpu.render(cam, api.pt_params(w, h, spp=2, bounces=1, variant=1, rect=rect), pieces)
    assert np.array_equal(whole, pieces)
    # Accumulation: 1 + 1 samples with consecutive frame ids == 2 samples in one call (mImage += pixel).
    acc = gpu.render(cam, api.pt_params(w, h, spp=1, bounces=1, variant=1, frame_id=0))
    gpu.render(cam, api.pt_params(w, h, spp=1, bounces=1, variant=1, frame_id=1), acc)
    assert np.array_equal(acc, whole)


def test_soup_four_bounces_config4_in_miniature(gpu, port, ref, api):
    """BASELINE config 4 at reduced size: voxelised solids (2048^3), 4 bounces, 2 spp, LOD 0.0035."""
    from oracle import pyoracle
    sc = api.Scene("soup", 11, seed=5)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, ocam = both_cameras(api, port, sc)
    p = api.pt_params(320, 180, spp=2, bounces=4, variant=api.VARIANT_RECURSIVE)
    sd = port.find_subdags(sc.nodes, sc.root)
    want, _, nrays = port.render(sc.nodes, sd, sc.colours, ocam, oracle_params(pyoracle, p), threads=8)
    got = gpu.render(cam, p)
    assert nrays > 4 * 320 * 180
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # one hop, on a 96-row band (the reference's tracer is single threaded)
    band = api.pt_params(320, 180, spp=2, bounces=4, variant=api.VARIANT_RECURSIVE, rect=(0, 40, 320, 136))
    direct = reference_image(ref, sc, cam, band)
    assert np.array_equal(got[40:136].view(np.uint32), direct[40:136].view(np.uint32))


def test_render_device_pointer_and_sample_sharding(gpu, port, api, scenes):
    """Device-pointer entry point; and samples split over two calls (as two GPUs would) sum to the same image
    up to float re-association of the per-pixel sums."""
    torch = pytest.importorskip("torch")
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, _ = both_cameras(api, port, sc)
    w, h = 128, 96
    whole = gpu.render(cam, api.pt_params(w, h, spp=4, bounces=1, variant=1))
    stream = torch.cuda.current_stream().cuda_stream
    a = torch.zeros(h, w, 3, device="cuda")
    b = torch.zeros(h, w, 3, device="cuda")
    gpu.render_device(cam, api.pt_params(w, h, spp=2, bounces=1, variant=1, frame_id=0), a.data_ptr(), stream)
    gpu.render_device(cam, api.pt_params(w, h, spp=2, bounces=1, variant=1, frame_id=2), b.data_ptr(), stream)
    torch.cuda.synchronize()
    np.testing.assert_allclose((a + b).cpu().numpy(), whole, rtol=1e-6, atol=1e-6)
    c = torch.zeros(h, w, 3, device="cuda")
    gpu.render_device(cam, api.pt_params(w, h, spp=4, bounces=1, variant=1), c.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(c.cpu().numpy(), whole)


def test_sample_group_size_does_not_change_the_image(gpu, port, api, scenes):
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, _ = both_cameras(api, port, sc)
    p = api.pt_params(120, 90, spp=7, bounces=2, variant=1)
    imgs = []
    try:
        for g in (1, 3, 4, 16):
            gpu.set_option("sample_group", g)
            imgs.append(gpu.render(cam, p))
    finally:
        gpu.set_option("sample_group", 0)
    for im in imgs[1:]:
        assert np.array_equal(im.view(np.uint32), imgs[0].view(np.uint32))


def test_band_interleaved_shares_tile_the_frame_exactly(gpu, port, api, scenes):
    """Round-robin tile sharding in ONE call per GPU: `bands=(count, index)` renders every count-th 64-row
    band. The shares are disjoint, cover the frame, and sum to the whole-frame render bit for bit."""
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, _ = both_cameras(api, port, sc)
    w, h = 90, 200            # 4 bands: 64 + 64 + 64 + 8 rows
    whole = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=2, variant=1))
    for count in (2, 3, 8):
        total = np.zeros_like(whole)
        for index in range(count):
            share = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=2, variant=1, bands=(count, index)))
            rows = np.array([((y // 64) % count) == index for y in range(h)])
            assert not share[~rows].any()                       # nothing outside the share
            assert np.array_equal(share[rows], whole[rows])     # and exactly the frame inside it
            total += share
        assert np.array_equal(total, whole)
    # a banded sub-rectangle
    part = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=2, variant=1, rect=(10, 30, 80, 190), bands=(2, 1)))
    rows = np.array([30 <= y < 190 and (((y - 30) // 64) % 2) == 1 for y in range(h)])
    assert np.array_equal(part[rows][:, 10:80], whole[rows][:, 10:80]) and not part[~rows].any() and not part[:, :10].any()


@pytest.mark.parametrize("sun,sky,bounces,spp", [(1, 0, 3, 5), (1, 1, 0, 3), (1, 1, 5, 2), (0, 1, 0, 2), (1, 0, 0, 17)])
def test_shared_sun_ray_and_depth_limits(gpu, port, ref, api, scenes, sun, sky, bounces, spp):
    """The depth-0 sun ray is traced once per pixel for all samples of a wave: sun only (no per-path shadow ray at depth 0,
    one deeper), no bounces at all (the first shade kernel is also the last), the deepest recursion the API takes, more samples
    than one wave holds, an odd image size (runs of 512 paths straddle sample boundaries) -- bit for bit the oracle's and the
    reference's image."""
    from oracle import pyoracle
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam, ocam = both_cameras(api, port, sc)
    p = api.pt_params(83, 61, spp=spp, bounces=bounces, variant=api.VARIANT_RECURSIVE, include_sun=sun, include_sky=sky)
    sd = port.find_subdags(sc.nodes, sc.root)
    want, _, _ = port.render(sc.nodes, sd, sc.colours, ocam, oracle_params(pyoracle, p), threads=4)
    got = gpu.render(cam, p)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    direct = reference_image(ref, sc, cam, p)
    assert np.array_equal(got.view(np.uint32), direct.view(np.uint32))

"""The GPU path-tracing viewer's passes -- progressive tile groups, RGBA accumulation, normalise, blur --
against numpy restatements of the shaders (glsl/pathtracing.frag:287-296,789-803, glsl/normalise.frag,
glsl/horz_blur.frag, glsl/vert_blur.frag) and against whole-frame renders of the same samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PI_F = float(np.float32(3.14159265358979))


def bit_mix(h):
    """glsl/pathtracing.frag:287-296 on uint32 arrays."""
    h = np.asarray(h, dtype=np.uint64)
    m = np.uint64(0xffffffff)
    h ^= h >> np.uint64(16); h = (h * np.uint64(0x85ebca6b)) & m
    h ^= h >> np.uint64(13); h = (h * np.uint64(0xc2b2ae35)) & m
    h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


def group_of_pixels(w, h, count):
    """tile group of every pixel: bitMix((tileX << 16) | tileY) % count, tiles of 64x64, rows from the top."""
    ys, xs = np.mgrid[0:h, 0:w]
    return bit_mix(((xs >> 6).astype(np.uint64) << np.uint64(16)) | (ys >> 6).astype(np.uint64)) % np.uint32(count)


@pytest.mark.parametrize("variant,bounces", [(1, 2), (0, 1)])
def test_tile_groups_partition_the_frame(gpu, api, scenes, variant, bounces):
    sc = scenes("sphere_noise", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam = api.default_camera(sc.lower, sc.upper)
    w, h = 200, 150                                   # 4 x 3 tiles, the last column and row overhang the image
    whole = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=bounces, variant=variant, frame_id=5))
    for count in (16, 3):
        groups = group_of_pixels(w, h, count)
        total = np.zeros_like(whole)
        for index in range(count):
            part = gpu.render(cam, api.pt_params(w, h, spp=2, bounces=bounces, variant=variant, frame_id=5, tile_group=(count, index)))
            mine = groups == index
            assert not part[~mine].any()
            assert np.array_equal(part[mine].view(np.uint32), whole[mine].view(np.uint32))
            total += part
        assert np.array_equal(total, whole)
    with pytest.raises(api.CubiquityError):
        gpu.render(cam, api.pt_params(w, h, tile_group=(4, 4)))
    with pytest.raises(api.CubiquityError):
        gpu.render(cam, api.pt_params(w, h, tile_group=(4, 1), rect=(0, 0, 100, 150)))


def test_progressive_passes_accumulate_like_the_viewer(gpu, api, scenes):
    torch = pytest.importorskip("torch")
    sc = scenes("soup", 7)
    gpu.upload(sc.nodes, sc.root, sc.colours)
    cam = api.default_camera(sc.lower, sc.upper)
    w, h = 320, 200
    stream = torch.cuda.current_stream().cuda_stream
    rgba = torch.zeros(h, w, 4, device="cuda")
    p = api.pt_params(w, h, spp=1, bounces=1, variant=api.VARIANT_ONE_BOUNCE)       # what the GLSL viewer traces
    for frame in range(32):
        gpu.progressive_pass_device(cam, p, frame, rgba.data_ptr(), stream)
    torch.cuda.synchronize()
    got = rgba.cpu().numpy()
    assert (got[..., 3] == 2.0).all()                                               # every tile group came round twice
    groups = group_of_pixels(w, h, 16)
    want = np.zeros((h, w, 3), dtype=np.float32)
    for frame in range(32):
        full = gpu.render(cam, api.pt_params(w, h, spp=1, bounces=1, variant=api.VARIANT_ONE_BOUNCE, frame_id=frame))
        sel = groups == (frame % 16)
        want[sel] += full[sel]
    assert np.array_equal(got[..., :3], want)
    # normalise.frag
    rgb = torch.zeros(h, w, 3, device="cuda")
    gpu.normalise_device(rgba.data_ptr(), w, h, rgb.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(rgb.cpu().numpy(), got[..., :3] / got[..., 3:4])


def blur_reference(img, vertical):
    """horz_blur.frag / vert_blur.frag in numpy: 9 taps, alpha-difference edge stop, GL_REPEAT wrap; the taps are added in
    the shader's order (-1, +1, -2, +2, ...) so the float sums round the same way."""
    out = img.copy()
    acc = img[..., :3].copy()
    count = np.ones(img.shape[:2], dtype=np.float32)
    axis = 0 if vertical else 1
    for tap in range(1, 5):
        for sgn in (-1, 1):
            t = np.roll(img, -sgn * tap, axis=axis)
            ok = np.abs(img[..., 3] - t[..., 3]) < np.float32(0.5)
            acc = np.where(ok[..., None], acc + t[..., :3], acc)
            count = np.where(ok, count + np.float32(1.0), count)
    out[..., :3] = acc / count[..., None]
    return out


def test_blur_passes_match_the_shaders(gpu, api):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    w, h = 97, 61
    img = rng.uniform(0, 1, (h, w, 4)).astype(np.float32)
    img[..., 3] = rng.integers(1, 4, (h, w)).astype(np.float32)        # sample counts: edges where they differ
    d = torch.from_numpy(img.copy()).cuda()
    scratch = torch.zeros_like(d)
    stream = torch.cuda.current_stream().cuda_stream
    gpu.blur_device(d.data_ptr(), w, h, scratch.data_ptr(), 2, stream)
    torch.cuda.synchronize()
    want = img
    for _ in range(2):
        want = blur_reference(blur_reference(want, False), True)
    np.testing.assert_array_equal(d.cpu().numpy(), want)


def test_upload_dag_reads_the_reference_file(gpu, api, port, ref, scenes, tmp_path):
    from conftest import assert_hits_identical, mixed_rays
    sc = scenes("sphere_noise", 6)
    path = str(tmp_path / "v.dag")
    v = ref.volume().load_arrays(sc.nodes, sc.root)
    v.save(path)                                       # written by the reference
    gpu.upload_dag(path, sc.colours)
    nodes, root = v.nodes(), v.root()
    assert np.array_equal(gpu.download_nodes(), nodes)
    rays = mixed_rays(sc.lower, sc.upper, 20000, seed=1)
    want, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, True, -1.0)
    assert_hits_identical(gpu.intersect_volume(rays, True, -1.0), want, "volume uploaded from a .dag file")

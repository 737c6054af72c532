"""Pins the oracle's restatement of the bounce loop and the camera to the UNMODIFIED reference
(pathtracing_demo.cpp + camera.cpp compiled against a stub SDL.h, oracle/ref_pt_shim.cpp)."""
import numpy as np
import pytest

from oracle import pyoracle

PI_F = float(np.float32(3.14159265358979))


def pose(sc):
    lower, upper = np.asarray(sc.lower, dtype=np.float64), np.asarray(sc.upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    hd = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    return [centre[0], centre[1] - hd, centre[2] + hd], -(PI_F / 4.0), 0.0


@pytest.mark.parametrize("pitch,yaw", [(-(PI_F / 4.0), 0.0), (0.3, 2.1), (-1.2, -0.7)])
def test_camera_rays_bit_exact(port, ref, pitch, yaw):
    pos = [12.5, -300.25, 280.0]
    want = ref.camera_rays(pos, pitch, yaw, 97, 61)
    got = port.camera_rays(port.camera(pos, pitch, yaw), 97, 61)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("variant,bounces,spp", [(0, 1, 1), (1, 0, 1), (1, 1, 2), (1, 4, 1)])
def test_bounce_loop_bit_exact(port, ref, scenes, variant, bounces, spp):
    sc = scenes("sphere_noise", 6)
    pos, pitch, yaw = pose(sc)
    p = pyoracle.PtParams(64, 48, spp, bounces, variant, 1, 1, 1, 0.0035, 3, 0, 0, 64, 48, 0)
    want, _ = ref.pt_render(sc.nodes, sc.root, sc.colours, pos, pitch, yaw, p)
    got, _, _ = port.render(sc.nodes, port.find_subdags(sc.nodes, sc.root), sc.colours, port.camera(pos, pitch, yaw), p)
    assert want.std() > 0.05
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()


def test_flags_and_rectangles(port, ref, scenes):
    sc = scenes("soup", 6)
    pos, pitch, yaw = pose(sc)
    cam = port.camera(pos, pitch, yaw)
    sd = port.find_subdags(sc.nodes, sc.root)
    for sun, sky, noise, mf in [(0, 1, 1, -1.0), (1, 0, 0, 0.0035), (0, 0, 1, 0.05)]:
        p = pyoracle.PtParams(40, 30, 1, 2, 1, sun, sky, noise, mf, 0, 5, 3, 33, 28, 0)
        want, _ = ref.pt_render(sc.nodes, sc.root, sc.colours, pos, pitch, yaw, p)
        got, _, _ = port.render(sc.nodes, sd, sc.colours, cam, p)
        assert np.array_equal(got, want)
        assert not got[:3].any() and not got[:, :5].any()      # outside the rectangle stays untouched


def test_as_is_global_rng_is_statistically_equivalent(port, ref, scenes):
    """The reference as shipped never re-seeds (pathtracing_demo.cpp:33): its image depends on scan order
    and cannot be matched bit for bit by any parallel renderer. Mean radiance must still agree."""
    sc = scenes("sphere_noise", 6)
    pos, pitch, yaw = pose(sc)
    p = pyoracle.PtParams(96, 72, 4, 1, 0, 1, 1, 1, 0.0035, 0, 0, 0, 96, 72, 0)
    as_is, _ = ref.pt_render(sc.nodes, sc.root, sc.colours, pos, pitch, yaw, p, reseed=False)
    ours, _, _ = port.render(sc.nodes, port.find_subdags(sc.nodes, sc.root), sc.colours, port.camera(pos, pitch, yaw), p)
    assert abs(as_is.mean() - ours.mean()) / as_is.mean() < 0.01
    mse = float(((as_is - ours) ** 2).mean()) / 16.0
    assert 10 * np.log10(1.0 / mse) > 15.0     # PSNR between two independent 4-spp estimates (peak 1.0)


BAD_SEED = 1973884838     # maps onto the mixer's fixed point 0, whose candidate point is always rejected (quirk Q8)


def test_rng_cycle_of_rejected_points_terminates(port):
    """The reference's rejection loop would spin for ever on this stream; the oracle gives up after 64 rejected
    candidates and returns the last one (outside the unit ball)."""
    # The mixer's fixed point is state 0 (bit_mix(0) == 0, base.cpp:72-77), whose candidate is (-1, -1, -1): rejected
    # for ever. BAD_SEED is one of the ~1 in 2^32 states that map straight onto it.
    pts, state = port.unit_ball_points(BAD_SEED, 16)
    assert state == 0 and (pts == -1.0).all()
    pts0, state0 = port.unit_ball_points(0, 3)
    assert state0 == 0 and (pts0 == -1.0).all()
    good, _ = port.unit_ball_points(17, 64)                   # the reference's own start state (pathtracing_demo.cpp:33)
    assert ((good.astype(np.float64) ** 2).sum(axis=1) < 1.0).all()


@pytest.mark.gpu
def test_device_rng_matches_oracle_including_the_stuck_stream(gpu, port):
    torch = pytest.importorskip("torch")
    seeds = np.concatenate([[BAD_SEED, 17, 0, 0xffffffff], np.random.default_rng(5).integers(0, 2**32, 5000)]).astype(np.uint32)
    draws = 16
    d_seeds = torch.from_numpy(seeds.view(np.int32)).cuda()
    d_points = torch.zeros(len(seeds) * draws * 3, dtype=torch.float32, device="cuda")
    d_states = torch.zeros(len(seeds), dtype=torch.int32, device="cuda")
    gpu.rng_points_device(d_seeds.data_ptr(), len(seeds), draws, d_points.data_ptr(), d_states.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got_p = d_points.cpu().numpy().reshape(len(seeds), draws, 3)
    got_s = d_states.cpu().numpy().view(np.uint32)
    for i in list(range(4)) + list(range(4, len(seeds), 97)):
        want_p, want_s = port.unit_ball_points(int(seeds[i]), draws)
        assert got_p[i].tobytes() == want_p.tobytes() and int(got_s[i]) == want_s, i

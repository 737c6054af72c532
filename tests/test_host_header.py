"""The C++ host mirror of the reference interface (cubiquity_b200/host/cubiquity_gpu.h)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "sphere32.dag")


@pytest.fixture(scope="module")
def exe(api):
    out = os.path.join(ROOT, "tests", "_build", "host_header_check")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir = os.path.dirname(api.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host_header_check.cpp"), "-o", out,
                    "-L" + libdir, "-lcubiquity_b200", "-Wl,-rpath," + libdir], check=True)
    return out


def run(exe):
    p = subprocess.run([exe, GOLD], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    return [line.split() for line in p.stdout.strip().splitlines()]


def test_header_compiles_and_subdags_match(exe, port):
    from cubiquity_b200 import dagfile
    lines = run(exe)
    nodes, root = dagfile.read_dag(GOLD)
    want = port.find_subdags(nodes, root)
    got = [l for l in lines if l[0] == "subdag"]
    assert len(got) == 8
    for l in got:
        i = int(l[1])
        assert [int(v) for v in l[2:6]] == list(want["lower"][i]) + [int(want["height"][i])]
        assert int(l[6]) == int(want["node"][i])


def test_header_reports_missing_gpu_loudly(exe, api):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    lines = run(exe)
    assert lines[-1][0] == "nogpu" and "fallback" in " ".join(lines[-1])


@pytest.mark.gpu
def test_single_ray_call_matches_oracle(exe, port):
    from cubiquity_b200 import dagfile
    lines = run(exe)
    life = lines.pop()
    assert life[0] == "lifecycle", life
    assert lines[-1][0] == "hit", lines[-1]
    nodes, root = dagfile.read_dag(GOLD)
    ray = np.zeros(1, dtype=[("o", "<f4", 3), ("d", "<f4", 3)])
    ray["o"] = [0.25, 0.125, 100.0]
    ray["d"] = [0.001, 0.002, -1.0]
    want, _, _ = port.trace(nodes, port.find_subdags(nodes, root), ray, True, -1.0)
    got = lines[-1][1:]
    assert int(got[0]) == int(want["hit"][0]) == 1
    assert np.float32(float(got[1])) == want["distance"][0] and int(got[2]) == int(want["material"][0])
    assert [np.float32(float(v)) for v in got[3:6]] == list(want["position"][0])
    # carve at the hit -> the surface recedes; back to the old root -> the same hit again; bake of the un-edited root
    # keeps the scene's node count (the golden volume is already merged) and the same hit
    after_hit, after_d, undone_hit, undone_d, baked_count, _, rebaked_hit, rebaked_d = life[1:]
    assert int(after_hit) == 1 and float(after_d) > float(got[1]) + 1.0
    assert int(undone_hit) == 1 and np.float32(float(undone_d)) == want["distance"][0]
    assert int(baked_count) == len(nodes)
    assert int(rebaked_hit) == 1 and np.float32(float(rebaked_d)) == want["distance"][0]


@pytest.fixture(scope="module")
def render_exe(api):
    out = os.path.join(ROOT, "tests", "_build", "render_dag")
    libdir = os.path.dirname(api.library_path())
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "examples", "render_dag.cpp"), "-I" + os.path.join(ROOT, "include"),
                    "-o", out, "-L" + libdir, "-lcubiquity_b200", "-Wl,-rpath," + libdir], check=True)
    return out


def test_example_program_builds_and_fails_loudly_without_a_gpu(render_exe, api, tmp_path):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([render_exe, GOLD, str(tmp_path / "o.ppm")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "fallback" in p.stderr


@pytest.mark.gpu
def test_example_program_renders_what_the_library_renders(render_exe, gpu, api, ref, tmp_path):
    """The C++ example (file in, PPM out) against the same frame rendered through the Python binding, with the
    camera placed from the REFERENCE's bounds (cubiquity_estimate_bounds) -- byte-identical images."""
    from cubiquity_b200 import dagfile
    out = tmp_path / "o.ppm"
    p = subprocess.run([render_exe, GOLD, str(out), "96", "64", "3", "2"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    data = out.read_bytes()
    header = b"P6\n96 64\n255\n"
    assert data.startswith(header)
    got = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(64, 96, 3)
    nodes, root = dagfile.read_dag(GOLD)
    _, lo, hi = ref.volume().load(GOLD).bounds()
    gpu.upload(nodes, root)                                     # purple default colours, like the example
    img = gpu.render(api.default_camera(lo, hi), api.pt_params(96, 64, spp=3, bounces=2, variant=api.VARIANT_RECURSIVE))
    want = np.clip(img * np.float32(255.0 / 3.0), 0, 255).astype(np.uint8)
    assert np.array_equal(got, want)
    assert got.std() > 10

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402  (tests are one of the few places allowed to use oracle/)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        from cubiquity_b200 import api
        return api.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip silently.
    pass


@pytest.fixture(scope="session")
def port():
    pyoracle.build(port=True, ref=False)
    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build. Available when oracle/_ref was built (here, or shipped to the
    GPU box); skipped only when neither the .so nor /root/reference exists."""
    if not os.path.exists(pyoracle.REF_SO) and not os.path.isdir(pyoracle.REFERENCE_ROOT):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def api():
    from cubiquity_b200 import api as _api
    _api.load_library()
    return _api


@pytest.fixture(scope="session")
def hostcore():
    """tests/host_core_check.cpp: the device traversal core compiled for the host (test only)."""
    src = os.path.join(ROOT, "tests", "host_core_check.cpp")
    out_dir = os.path.join(ROOT, "tests", "_build")
    so = os.path.join(out_dir, "libhostcore.so")
    os.makedirs(out_dir, exist_ok=True)
    dep = os.path.join(ROOT, "cubiquity_b200", "csrc", "traverse.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                        "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.host_core_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_void_p]

    def trace(nodes, subdags, rays, surface=True, max_footprint=-1.0):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        rays = np.ascontiguousarray(rays, dtype=pyoracle.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=pyoracle.HIT_DTYPE)
        lib.host_core_trace(nodes.ctypes.data, subdags.ctypes.data, rays.ctypes.data, len(rays), int(surface),
                            float(max_footprint), hits.ctypes.data)
        return hits
    return trace


@pytest.fixture(scope="session")
def scenes(api):
    """Small procedural volumes shared by the tests, built once."""
    cache = {}

    def get(kind, size_log2, seed=1):
        key = (kind, size_log2, seed)
        if key not in cache:
            cache[key] = api.Scene(kind, size_log2, seed)
        return cache[key]
    return get


@pytest.fixture(scope="session")
def gpu(api):
    if api.device_count() <= 0:
        pytest.fail("this test needs a CUDA device; cubiquity_b200 has no CPU fallback")
    ctx = api.Context(0)
    yield ctx
    ctx.close()


def mixed_rays(lower, upper, n, seed):
    """Reference-style in-bounds rays + outside-looking-in rays + axis-aligned and diagonal corner cases."""
    from cubiquity_b200 import rays as R
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    a = R.in_bounds_rays(n // 2, lower, upper, seed)
    b = R.random_rays(n - n // 2, lower, upper, seed + 1, dilate=0.3)
    rays = np.concatenate([a, b])
    k = min(200, n // 10)
    rng = np.random.default_rng(seed)
    axes = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]], dtype=np.float32)
    rays["d"][:k] = axes[rng.integers(0, 6, k)]
    # voxel-centre origins with exact diagonal directions: edge and corner entries (quirk Q3)
    rays["o"][k:2 * k] = np.round(rays["o"][k:2 * k])
    s = np.sign(rays["d"][k:2 * k])
    s[s == 0] = 1
    rays["d"][k:2 * k] = s * np.float32(1.0 / np.sqrt(3.0))
    return rays


def assert_hits_identical(a, b, what="", nan_payload_insensitive=False):
    """Bit equality of every field (the primary comparison). With nan_payload_insensitive the
    sign/payload bits of NaNs are ignored (x86 and the GPU generate different default NaNs; only rays fed
    NaN or 0/0 inputs ever produce them)."""
    assert a.dtype.itemsize == b.dtype.itemsize == 40 and len(a) == len(b)
    if a.tobytes() == b.tobytes():
        return
    av = a.view(np.uint32).reshape(-1, 10).copy()
    bv = b.view(np.uint32).reshape(-1, 10).copy()
    if nan_payload_insensitive:
        for v in (av, bv):
            f = v[:, [1, 3, 4, 5, 6, 7, 8]]
            f[(f & 0x7fffffff) > 0x7f800000] = 0x7fc00000
            v[:, [1, 3, 4, 5, 6, 7, 8]] = f
        if (av == bv).all():
            return
    bad = np.nonzero((av != bv).any(axis=1))[0]
    raise AssertionError("%s: %d of %d hit records differ, first at %d:\n  %r\n  %r"
                         % (what, len(bad), len(a), bad[0], a[bad[0]], b[bad[0]]))


def reference_hits(volume, port, rays, surface, max_footprint):
    """Runs the port, then the UNMODIFIED reference on every ray the port could finish, and returns
    (port_hits, reference_hits, mask). Rays the port abandons (quirks Q5/Q6: the reference never
    returns for them) are left out of the reference run so the test process cannot hang."""
    nodes, root = volume.nodes(), volume.root()
    got, _, _ = port.trace(nodes, port.find_subdags(nodes, root), rays, surface, max_footprint)
    mask = got["pad"] == 0
    want, _ = volume.intersect(rays[mask], surface, max_footprint)
    return got, want, mask


def assert_matches_reference(ref, nodes, root, rays, got, port_hits, surface, max_footprint, what=""):
    """ONE hop: hits from the CUDA path against the UNMODIFIED reference's intersectVolume (oracle/_ref) on the same rays.
    Rays the port abandons (quirks Q5 / Q6: the reference itself never returns for them) are left out so that the test
    process cannot hang; for those the CUDA path must report `abandoned` as the port does."""
    mask = port_hits["pad"] == 0
    vol = ref.volume().load_arrays(nodes, root)
    want, _ = vol.intersect(np.ascontiguousarray(rays[mask]), surface, max_footprint, threads=os.cpu_count() or 1)
    assert_hits_identical(np.ascontiguousarray(got[mask]), want, what + " vs the reference itself")
    assert (got["status"][~mask] == 1).all()

"""The C-ABI library: it loads, exports every symbol include/cubiquity_b200.h declares, its host-only
helpers agree with the oracle, and without a GPU every compute entry point fails loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "cubiquity_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cbq_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(api):
    lib = api.load_library()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(api.EXPORTS) == names


def test_library_is_in_tree(api):
    assert api.library_path().startswith(os.path.join(ROOT, "cubiquity_b200", "lib"))


def test_record_layouts_match_header(api):
    assert api.RAY_DTYPE.itemsize == 24 and api.HIT_DTYPE.itemsize == 40 and api.SUBDAG_DTYPE.itemsize == 32
    assert C.sizeof(api.Camera) == 104 and C.sizeof(api.PtParams) == 72 and api.COMPACT_DTYPE.itemsize == 8


def test_find_subdags_matches_oracle(api, port, scenes):
    for kind, sl in [("sphere_noise", 6), ("terrain", 7), ("city", 10)]:
        sc = scenes(kind, sl)
        a, b = api.find_subdags(sc.nodes, sc.root), port.find_subdags(sc.nodes, sc.root)
        assert a.tobytes() == b.tobytes()


def test_find_subdags_rejects_corrupt_arrays(api, scenes):
    sc = scenes("sphere_noise", 6)
    bad = sc.nodes.copy()
    bad[sc.root] = 0xfffffff0
    with pytest.raises(api.CubiquityError) as e:
        api.find_subdags(bad, sc.root)
    assert e.value.code == api.ERROR_CORRUPT_VOLUME
    with pytest.raises(api.CubiquityError):
        api.find_subdags(sc.nodes[:100], sc.root)


def test_camera_basis_matches_oracle(api, port):
    for pos, pitch, yaw in [((0, -300, 300), -0.785398, 0.0), ((12.5, 7, -3), 0.3, 2.1), ((0, 0, 0), 0, 0)]:
        a = api.camera_from_pose(pos, pitch, yaw, 60.0)
        b = port.camera(pos, pitch, yaw, 60.0)
        assert bytes(a) == bytes(b)


def test_no_gpu_means_loud_failure_not_fallback(api):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.CubiquityError) as e:
        api.Context(0)
    assert e.value.code == api.ERROR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(api.CubiquityError):
        api.PinnedArray(16, np.uint8)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "cubiquity_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(d, f), errors="replace").read()
                # comments may NAME the checker; nothing may include, import, load or link it
                for needle in ("pyoracle", "libcbq_oracle", "libcbq_ref", "cbq_oracle.h", "import oracle", "from oracle"):
                    assert needle not in text, (f, needle)

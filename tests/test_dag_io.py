"""cbq_dag_load / cbq_dag_save (csrc/host_shim.cpp): the reference's .dag file (storage.cpp:192-206, 505-542) read and
written with 64-bit sizes and validation, against the reference's own Volume::save / Volume::load."""
import os

import numpy as np
import pytest


def test_round_trip_with_the_reference(api, ref, scenes, tmp_path):
    sc = scenes("soup", 7)
    ours = str(tmp_path / "ours.dag")
    api.dag_save(ours, sc.nodes, sc.root)
    v = ref.volume().load(ours)                                  # the reference reads what we wrote ...
    assert np.array_equal(v.nodes(), sc.nodes) and v.root() == sc.root
    theirs = str(tmp_path / "theirs.dag")
    v.save(theirs)                                               # ... (save() bakes: order may change, content not) ...
    nodes, root = api.dag_load(theirs)                           # ... and we read what it writes
    again = ref.volume().load(theirs)
    assert np.array_equal(nodes, again.nodes()) and root == again.root()
    from cubiquity_b200.dagfile import read_dag
    n2, r2 = read_dag(ours)
    a, b = api.dag_load(ours)
    assert np.array_equal(a, n2) and b == r2 and np.array_equal(a, sc.nodes)


def test_damaged_files_are_refused(api, scenes, tmp_path):
    sc = scenes("sphere_noise", 6)
    good = str(tmp_path / "good.dag")
    api.dag_save(good, sc.nodes, sc.root)
    raw = open(good, "rb").read()
    cases = {
        "truncated": raw[:-5],
        "trailing": raw + b"\0" * 32,
        "header_only": raw[:8],
        "short_header": raw[:6],
        "count_too_big": raw[:4] + np.array([len(sc.nodes)], dtype="<u4").tobytes() + raw[8:],     # claims 256 more nodes than stored
        "root_past_end": np.array([len(sc.nodes) + 7], dtype="<u4").tobytes() + raw[4:],
    }
    bad_child = bytearray(raw)
    bad_child[8:12] = np.array([len(sc.nodes) + 1], dtype="<u4").tobytes()
    cases["child_past_end"] = bytes(bad_child)
    for name, data in cases.items():
        path = str(tmp_path / (name + ".dag"))
        open(path, "wb").write(data)
        with pytest.raises(api.CubiquityError) as e:
            api.dag_load(path)
        assert e.value.code == api.ERROR_CORRUPT_VOLUME, name
    with pytest.raises(api.CubiquityError) as e:
        api.dag_load(str(tmp_path / "missing.dag"))
    assert e.value.code == api.ERROR_INVALID_ARGUMENT
    # an empty volume (root = material 0, no stored nodes) is a valid 8-byte file
    from cubiquity_b200.dagfile import material_nodes
    empty = str(tmp_path / "empty.dag")
    api.dag_save(empty, material_nodes(), 0)
    assert os.path.getsize(empty) == 8
    nodes, root = api.dag_load(empty)
    assert len(nodes) == 256 and root == 0
    assert not os.path.exists(empty + ".partial")


def test_log_callback_receives_the_messages(api):
    got = []
    api.set_log_callback(got.append)
    try:
        L = api.load_library()
        assert L.cbq_find_subdags(None, 0, 0, None) == api.ERROR_INVALID_ARGUMENT
    finally:
        api.set_log_callback(None)
    assert got and "node array" in got[0]

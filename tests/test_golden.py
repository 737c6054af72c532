"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py) -- the pin that
travels to machines without /root/reference. Checked against the oracle port, the host-compiled device
core and (on the GPU box) the CUDA path."""
import os

import numpy as np
import pytest

from conftest import assert_hits_identical
from cubiquity_b200 import dagfile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLD, "raycast.npz"))
    nodes, root = dagfile.read_dag(os.path.join(GOLD, "sphere32.dag"))
    assert root == int(z["root"])
    return z, nodes, root


def cases(z):
    return [(bool(s), float(m)) for s, m in z["cases"]]


def test_port_reproduces_reference_vectors(port, gold):
    z, nodes, root = gold
    sd = port.find_subdags(nodes, root)
    for f in ("lower", "height", "node"):
        assert (sd[f] == z["subdags"][f]).all()
    for i, (surf, mf) in enumerate(cases(z)):
        got, _, _ = port.trace(nodes, sd, z["rays"], surf, mf)
        assert_hits_identical(got, z["hits_%d" % i], "baked case %d" % i)
    enodes, eroot = z["edited_nodes"], int(z["edited_root"])
    esd = port.find_subdags(enodes, eroot)
    for f in ("lower", "height", "node"):
        assert (esd[f] == z["edited_subdags"][f]).all()
    for i, (surf, mf) in enumerate(cases(z)):
        got, _, _ = port.trace(enodes, esd, z["edited_rays"], surf, mf)
        assert_hits_identical(got, z["edited_hits_%d" % i], "edited case %d" % i)


def test_device_core_on_host_reproduces_reference_vectors(port, hostcore, gold):
    z, nodes, root = gold
    sd = port.find_subdags(nodes, root)
    for i, (surf, mf) in enumerate(cases(z)):
        assert_hits_identical(hostcore(nodes, sd, z["rays"], surf, mf), z["hits_%d" % i], "case %d" % i)


def test_vectors_exercise_the_quirks(gold):
    z, _, _ = gold
    h = z["hits_0"]
    assert 500 < h["hit"].sum() < len(h)
    assert (h["distance"] < 0).any()                        # Q1: negative-distance node hits
    assert (np.abs(h["normal"]).sum(axis=1) > 1).any()      # Q3: edge / corner normals
    assert (z["hits_3"]["material"] != z["hits_0"]["material"]).any() or (z["hits_3"]["distance"] != z["hits_0"]["distance"]).any()  # LOD bites


@pytest.mark.gpu
def test_gpu_reproduces_reference_vectors(gpu, gold):
    z, nodes, root = gold
    gpu.upload(nodes, root)
    for f in ("lower", "height", "node"):
        assert (gpu.subdags()[f] == z["subdags"][f]).all()
    for i, (surf, mf) in enumerate(cases(z)):
        assert_hits_identical(gpu.intersect_volume(z["rays"], surf, mf), z["hits_%d" % i], "baked case %d" % i)
    # the edit arrives as a delta: upload the shared prefix first, then the tail
    enodes, eroot, shared_end = z["edited_nodes"], int(z["edited_root"]), int(z["edited_shared_end"])
    gpu.upload(enodes[:shared_end], root)
    gpu.update(enodes, shared_end, eroot)
    for i, (surf, mf) in enumerate(cases(z)):
        assert_hits_identical(gpu.intersect_volume(z["edited_rays"], surf, mf), z["edited_hits_%d" % i], "edited case %d" % i)

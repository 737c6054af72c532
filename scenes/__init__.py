"""Procedural SVDAG scenes (scenes/scene_builder.cpp): the synthetic inputs of the tests and benchmarks.

Host only and a library of its own (scenes/lib/libcbq_scenes.so, built with g++): importing this package maps no
product code, so bench.py's reference arm can build the same input as the GPU arm.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libcbq_scenes.so")
_lib = None


def build(force=False):
    src = [os.path.join(HERE, "scene_builder.cpp"), os.path.join(HERE, "cbq_scenes.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in src):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    tmp = LIB + ".partial"
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-o", tmp, src[0], "-lpthread"], check=True)
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.cbq_scene_build.argtypes = [C.c_char_p, u32, u64, C.POINTER(vp)]
        L.cbq_scene_nodes.restype = C.POINTER(u32)
        L.cbq_scene_nodes.argtypes = [vp, C.POINTER(u64)]
        L.cbq_scene_root.restype = u32
        L.cbq_scene_root.argtypes = [vp]
        L.cbq_scene_bounds.restype = None
        L.cbq_scene_bounds.argtypes = [vp, vp, vp]
        L.cbq_scene_colours.restype = None
        L.cbq_scene_colours.argtypes = [vp, vp]
        L.cbq_scene_voxels.restype = None
        L.cbq_scene_voxels.argtypes = [vp, vp, u64, vp]
        L.cbq_scene_free.restype = None
        L.cbq_scene_free.argtypes = [vp]
        _lib = L
    return _lib


class Scene:
    """A procedural volume: nodes (n, 8) uint32 incl. the 256 material nodes, root, lower/upper bounds, colours."""

    def __init__(self, kind, size_log2, seed=1):
        L = load()
        self.kind, self.size_log2, self.seed = kind, int(size_log2), int(seed)
        h = C.c_void_p()
        rc = L.cbq_scene_build(kind.encode(), int(size_log2), int(seed), C.byref(h))
        if rc != 0:
            raise ValueError("cbq_scene_build(%r, %d) failed with code %d" % (kind, size_log2, rc))
        self._h = h
        n = C.c_uint64()
        p = L.cbq_scene_nodes(h, C.byref(n))
        self.nodes = np.ctypeslib.as_array(p, shape=(int(n.value), 8))   # view into the scene's memory
        self.root = int(L.cbq_scene_root(h))
        lo = np.zeros(3, dtype=np.int32)
        hi = np.zeros(3, dtype=np.int32)
        L.cbq_scene_bounds(h, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p))
        self.lower, self.upper = lo, hi
        self.colours = np.zeros((256, 3), dtype=np.float32)
        L.cbq_scene_colours(h, self.colours.ctypes.data_as(C.c_void_p))

    def voxels(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        load().cbq_scene_voxels(self._h, xyz.ctypes.data_as(C.c_void_p), len(xyz), out.ctypes.data_as(C.c_void_p))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self.nodes = None
            load().cbq_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

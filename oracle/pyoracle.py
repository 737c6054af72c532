"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle.

`Port`   -> oracle/_build/libcbq_oracle.so  (plain-C restatement, oracle/cbq_oracle.c)
`Ref`    -> oracle/_ref/libcbq_ref.so       (the unmodified reference compiled from /root/reference
                                             plus oracle/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. Nothing under cubiquity_b200/ does.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libcbq_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcbq_ref.so")
REFERENCE_ROOT = "/root/reference"

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3)])
HIT_DTYPE = np.dtype([("hit", "<u4"), ("distance", "<f4"), ("material", "<u4"),
                      ("position", "<f4", 3), ("normal", "<f4", 3), ("pad", "<u4")])
SUBDAG_DTYPE = np.dtype([("lower", "<i4", 3), ("height", "<i4"), ("pad0", "<u4"),
                         ("node", "<u4"), ("pad1", "<u4"), ("pad2", "<u4")])
assert RAY_DTYPE.itemsize == 24 and HIT_DTYPE.itemsize == 40 and SUBDAG_DTYPE.itemsize == 32


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rays", "hits", "subdag_entries", "iterations", "descents", "pops", "material_steps")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    def node_visits(self):
        """V of SURVEY 8(d): sub-DAG entries + descents + pops."""
        return int(self.subdag_entries + self.descents + self.pops)


class Camera(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("forward", C.c_double * 3), ("up", C.c_double * 3),
                ("right", C.c_double * 3), ("scale", C.c_float), ("pad", C.c_float)]


class PtParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("bounces", C.c_uint32),
                ("variant", C.c_uint32), ("include_sun", C.c_uint32), ("include_sky", C.c_uint32),
                ("add_noise", C.c_uint32), ("max_footprint", C.c_float), ("frame_id", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32), ("pad", C.c_uint32)]


def pt_params_from(p):
    """Oracle PtParams from the product's (cubiquity_b200.api.PtParams): same fields by name; the product's
    band interleave has no oracle counterpart (tests compare banded renders with masked whole frames)."""
    names = [n for n, _ in PtParams._fields_ if n != "pad"]
    return PtParams(**{n: getattr(p, n) for n in names})


def build(port=True, ref=None):
    """Compile the checker(s). `ref` defaults to "if /root/reference exists"."""
    if ref is None:
        ref = os.path.isdir(REFERENCE_ROOT)
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    if targets:
        subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True)


def write_dag(path, nodes, root):
    """The checker's own writer of the reference's .dag file (storage.cpp:192-206: u32 root, u32 count, then the
    count non-material nodes) -- inputs reach the reference without passing through product code."""
    nodes = np.ascontiguousarray(nodes, dtype="<u4").reshape(-1, 8)
    with open(path, "wb") as f:
        np.array([int(root), len(nodes) - 256], dtype="<u4").tofile(f)
        nodes[256:].tofile(f)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Port:
    """The plain-C restatement. experiments=True loads the design-study build (oracle/experiments/ compiled in, `make
    experiments`) instead of the parity checker; only scripts/brick_equivalence.py asks for it."""

    def __init__(self, experiments=False):
        so = PORT_SO
        if experiments:
            so = os.path.join(os.path.dirname(PORT_SO), "libcbq_oracle_experiments.so")
            subprocess.run(["make", "-s", "-C", HERE, "experiments"], check=True)
        elif not os.path.exists(PORT_SO):
            build(port=True, ref=False)
        L = self.lib = C.CDLL(so)
        L.cbqo_trace.restype = C.c_double
        L.cbqo_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_float,
                                 C.c_void_p, C.c_int, C.c_void_p]
        L.cbqo_find_subdags.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.cbqo_camera_from_pose.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.cbqo_camera_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cbqo_render.restype = C.c_double
        L.cbqo_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]
        L.cbqo_bit_mix64.restype = C.c_uint64
        L.cbqo_bit_mix64.argtypes = [C.c_uint64]
        L.cbqo_fnv1a.restype = C.c_uint64
        L.cbqo_fnv1a.argtypes = [C.c_char_p, C.c_int64]
        L.cbqo_fmix32.restype = C.c_uint32
        L.cbqo_fmix32.argtypes = [C.c_uint32]
        L.cbqo_unit_ball_points.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_void_p]
        L.cbqo_pixel_seed.restype = C.c_uint32
        L.cbqo_pixel_seed.argtypes = [C.c_void_p, C.c_uint32]

    def unit_ball_points(self, seed, draws):
        state = C.c_uint32(int(seed))
        out = np.zeros((draws, 3), dtype=np.float32)
        self.lib.cbqo_unit_ball_points(C.byref(state), int(draws), _ptr(out))
        return out, int(state.value)

    def find_subdags(self, nodes, root):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        out = np.zeros(8, dtype=SUBDAG_DTYPE)
        self.lib.cbqo_find_subdags(_ptr(nodes), int(root), _ptr(out))
        return out

    def trace(self, nodes, subdags, rays, surface=True, max_footprint=-1.0, threads=1,
              want_hits=True, want_stats=False):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE) if want_hits else None
        st = Stats() if want_stats else None
        secs = self.lib.cbqo_trace(_ptr(nodes), _ptr(subdags), _ptr(rays), len(rays), int(bool(surface)),
                                   float(max_footprint), _ptr(hits) if want_hits else None, int(threads),
                                   C.byref(st) if want_stats else None)
        return hits, secs, st

    def camera(self, position, pitch, yaw, fov_degrees=60.0):
        cam = Camera()
        pos = (C.c_double * 3)(*[float(v) for v in position])
        self.lib.cbqo_camera_from_pose(pos, float(pitch), float(yaw), float(fov_degrees), C.byref(cam))
        return cam

    def camera_rays(self, cam, width, height):
        rays = np.zeros(width * height, dtype=RAY_DTYPE)
        self.lib.cbqo_camera_rays(C.byref(cam), int(width), int(height), _ptr(rays))
        return rays

    def render(self, nodes, subdags, colours, cam, params, accum=None, threads=1):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        colours = np.ascontiguousarray(colours, dtype=np.float32)
        if accum is None:
            accum = np.zeros((params.height, params.width, 3), dtype=np.float32)
        rays = C.c_uint64(0)
        secs = self.lib.cbqo_render(_ptr(nodes), _ptr(subdags), _ptr(colours), C.byref(cam), C.byref(params),
                                    _ptr(accum), int(threads), C.byref(rays))
        return accum, secs, int(rays.value)

    def last_render_stats(self):
        """Traversal statistics (Stats) of the last render() call: the path tracer's node visits."""
        st = Stats()
        self.lib.cbqo_last_render_stats(C.byref(st))
        return st


class Ref:
    """The unmodified reference library behind oracle/ref_shim.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            if not os.path.isdir(REFERENCE_ROOT):
                raise FileNotFoundError("oracle/_ref/libcbq_ref.so is not built and /root/reference is absent")
            build(port=False, ref=True)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_volume_new.restype = C.c_void_p
        L.ref_volume_free.argtypes = [C.c_void_p]
        L.ref_volume_load.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_volume_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_volume_set_voxel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint8]
        L.ref_volume_set_voxels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ref_volume_voxel.restype = C.c_uint8
        L.ref_volume_voxel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.ref_volume_voxels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_volume_fill.argtypes = [C.c_void_p, C.c_uint8]
        for f in ("ref_volume_bake", "ref_volume_checkpoint", "ref_volume_undo", "ref_volume_redo"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_volume_fill_sphere.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint8]
        L.ref_volume_nodes.restype = C.POINTER(C.c_uint32)
        L.ref_volume_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ref_volume_root.restype = C.c_uint32
        L.ref_volume_root.argtypes = [C.c_void_p]
        L.ref_volume_shared_end.restype = C.c_uint32
        L.ref_volume_shared_end.argtypes = [C.c_void_p]
        L.ref_find_subdags.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_estimate_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_intersect_volume.restype = C.c_double
        L.ref_intersect_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_void_p, C.c_int]
        L.ref_trace_ray_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_pt_render.restype = C.c_double
        L.ref_pt_render.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_float,
                                    C.c_int, C.c_void_p]
        L.ref_camera_rays.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.ref_voxelize.restype = C.c_double
        L.ref_voxelize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8, C.c_uint8, C.c_int, C.c_void_p]
        L.ref_bit_mix.restype = C.c_uint64
        L.ref_bit_mix.argtypes = [C.c_uint64]
        L.ref_fnv1a.restype = C.c_uint64
        L.ref_fnv1a.argtypes = [C.c_char_p, C.c_int64]

    def volume(self):
        return RefVolume(self)

    def camera_rays(self, position, pitch, yaw, width, height):
        """Camera::rayFromViewportPos (reference camera.cpp:12-38) for every pixel, cast to float."""
        pos = np.asarray(position, dtype=np.float64)
        rays = np.zeros(width * height, dtype=RAY_DTYPE)
        self.lib.ref_camera_rays(_ptr(pos), float(pitch), float(yaw), int(width), int(height), _ptr(rays))
        return rays

    def pt_render(self, nodes, root, colours, position, pitch, yaw, params, reseed=True, accum=None):
        """The reference's own PathtracingDemo::traceSingleRay / traceSingleRayRecurse over a pixel rectangle
        (oracle/ref_pt_shim.cpp). `params` is a PtParams. Single threaded: the reference's RNG is a global."""
        colours = np.ascontiguousarray(colours, dtype=np.float32)
        pos = np.asarray(position, dtype=np.float64)
        if accum is None:
            accum = np.zeros((params.height, params.width, 3), dtype=np.float32)
        pv = np.array([params.width, params.height, params.spp, params.bounces, params.variant, params.include_sun,
                       params.include_sky, params.add_noise, params.frame_id, params.x0, params.y0, params.x1, params.y1],
                      dtype=np.uint32)
        with tempfile.NamedTemporaryFile(suffix=".dag", delete=False) as f:
            path = f.name
        try:
            write_dag(path, nodes, root)
            secs = self.lib.ref_pt_render(os.fsencode(path), _ptr(colours), _ptr(pos), float(pitch), float(yaw), _ptr(pv),
                                          float(params.max_footprint), int(bool(reseed)), _ptr(accum))
        finally:
            os.unlink(path)
        return accum, secs


class RefVolume:
    def __init__(self, ref):
        self.L = ref.lib
        self.h = C.c_void_p(self.L.ref_volume_new())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_volume_free(self.h)
            self.h = None

    def load(self, path):
        if self.L.ref_volume_load(self.h, os.fsencode(path)) != 0:
            raise IOError("reference Volume::load failed: %s" % path)
        return self

    def load_arrays(self, nodes, root):
        """Round-trips through the reference's own .dag reader (storage.cpp:505-528)."""
        with tempfile.NamedTemporaryFile(suffix=".dag", delete=False) as f:
            path = f.name
        try:
            write_dag(path, nodes, root)
            self.load(path)
        finally:
            os.unlink(path)
        return self

    def save(self, path):
        self.L.ref_volume_save(self.h, os.fsencode(path))

    def set_voxels(self, xyzm):
        xyzm = np.ascontiguousarray(xyzm, dtype=np.int32).reshape(-1, 4)
        self.L.ref_volume_set_voxels(self.h, _ptr(xyzm), len(xyzm))

    def voxels(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self.L.ref_volume_voxels(self.h, _ptr(xyz), len(xyz), _ptr(out))
        return out

    def fill(self, m):
        self.L.ref_volume_fill(self.h, int(m))

    def bake(self):
        self.L.ref_volume_bake(self.h)

    def checkpoint(self):
        self.L.ref_volume_checkpoint(self.h)

    def undo(self):
        self.L.ref_volume_undo(self.h)

    def redo(self):
        self.L.ref_volume_redo(self.h)

    def fill_sphere(self, x, y, z, radius, material):
        self.L.ref_volume_fill_sphere(self.h, float(x), float(y), float(z), float(radius), int(material))

    def voxelize(self, triangles, materials, fill, background=0, thin=False):
        """The reference's voxelize(volume, mesh, fill, background) with Mesh::build (voxelization.cpp:692-823); the result
        replaces this volume's content. triangles: (n, 9) float32, materials: (n,) uint8. Returns (is_closed,
        is_inside_out, seconds). Runs in a process of its own (oracle/ref_voxelize_helper.py explains why) and comes back
        through the reference's own .dag writer and reader."""
        import sys
        tris = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 9)
        mats = np.ascontiguousarray(materials, dtype=np.uint8)
        with tempfile.TemporaryDirectory() as d:
            tris.tofile(os.path.join(d, "t.f32"))
            mats.tofile(os.path.join(d, "m.u8"))
            out = os.path.join(d, "v.dag")
            r = subprocess.run([sys.executable, os.path.join(HERE, "ref_voxelize_helper.py"), REF_SO, os.path.join(d, "t.f32"), os.path.join(d, "m.u8"),
                                str(int(fill)), str(int(background)), str(int(bool(thin))), out], capture_output=True, text=True, check=True)
            closed, inside_out, secs = r.stdout.split()
            self.load(out)
        return closed == "1", inside_out == "1", float(secs)

    def nodes(self):
        """Copy of the raw node array including the 256 material nodes, shape (n, 8)."""
        n = C.c_uint64(0)
        p = self.L.ref_volume_nodes(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(int(n.value), 8)).copy()

    def root(self):
        return int(self.L.ref_volume_root(self.h))

    def shared_end(self):
        return int(self.L.ref_volume_shared_end(self.h))

    def subdags(self):
        out = np.zeros(8, dtype=SUBDAG_DTYPE)
        self.L.ref_find_subdags(self.h, _ptr(out))
        return out

    def bounds(self):
        outside = C.c_uint8(0)
        lu = np.zeros(6, dtype=np.int32)
        self.L.ref_estimate_bounds(self.h, C.byref(outside), _ptr(lu))
        return int(outside.value), lu[:3].copy(), lu[3:].copy()

    def intersect(self, rays, surface=True, max_footprint=-1.0, threads=1, want_hits=True):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE) if want_hits else None
        secs = self.L.ref_intersect_volume(self.h, _ptr(rays), len(rays), int(bool(surface)), float(max_footprint),
                                           _ptr(hits) if want_hits else None, int(threads))
        return hits, secs

    def trace_ref(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.L.ref_trace_ray_ref(self.h, _ptr(rays), len(rays), _ptr(hits))
        return hits


def dag_signature(nodes, root, collapse=True):
    """Checker for bake parity: the canonical form of the DAG below `root`, independent of node order.

    Follows NodeStore::merge_node (reference storage.cpp:243-290): children first, a node whose eight merged
    children are one material IS that material, otherwise nodes are identified by content. Returns
    (signature of the root, number of distinct non-material nodes); two arrays describe the same volume with the
    same sharing iff both values agree. Iterative post-order, hashlib digests as content ids.
    collapse=False skips the material rule: the signature then identifies the UNFOLDED TREE exactly as stored (an
    un-merged internal node with eight equal material children stays a node), which is what ray traversal sees."""
    import hashlib
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
    sig = {}
    distinct = set()
    open_nodes = set()          # expanded, not finished: meeting one again means the array has a cycle

    def mat(m):
        return b"m" + int(m).to_bytes(4, "little")

    if root < 256:
        return mat(root), 0
    stack = [(int(root), False)]
    while stack:
        i, expanded = stack.pop()
        if i in sig:
            continue
        kids = [int(c) for c in nodes[i]]
        if not expanded:
            if i in open_nodes:
                raise ValueError("node %d is its own ancestor: not a DAG" % i)
            open_nodes.add(i)
            stack.append((i, True))
            for c in kids:
                if c >= len(nodes):
                    raise ValueError("node %d has child %d past the end of the array (%d)" % (i, c, len(nodes)))
                if c >= 256 and c not in sig:
                    stack.append((c, False))
            continue
        open_nodes.discard(i)
        parts = [mat(c) if c < 256 else sig[c] for c in kids]
        if collapse and parts[0][:1] == b"m" and all(p == parts[0] for p in parts):
            sig[i] = parts[0]
        else:
            d = hashlib.blake2b(b"".join(parts), digest_size=16).digest()
            sig[i] = b"n" + d
            distinct.add(d)
    return sig[int(root)], len(distinct)

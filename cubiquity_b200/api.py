"""ctypes binding of include/cubiquity_b200.h -- the host-side mirror used by tests and bench.py.

Names follow the reference (reference src/library/raytracing.h:69-81): `Context.intersect_volume`
is a batch of `Cubiquity::intersectVolume`, `find_subdags` is `findSubDAGs`, `Context.render` is
`PathtracingDemo::raytrace`. Everything here goes through the C ABI; there is no Python or CPU
implementation of the ray cast in this package and the import FAILS if the CUDA library is absent.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3)])
HIT_DTYPE = np.dtype([("hit", "<u4"), ("distance", "<f4"), ("material", "<u4"),
                      ("position", "<f4", 3), ("normal", "<f4", 3), ("status", "<u4")])
SUBDAG_DTYPE = np.dtype([("lower", "<i4", 3), ("height", "<i4"), ("pad0", "<u4"),
                         ("node", "<u4"), ("pad1", "<u4"), ("pad2", "<u4")])
COMPACT_DTYPE = np.dtype([("distance", "<f4"), ("code", "<u4")])      # cbq_hit_compact
assert RAY_DTYPE.itemsize == 24 and HIT_DTYPE.itemsize == 40 and SUBDAG_DTYPE.itemsize == 32 and COMPACT_DTYPE.itemsize == 8

TRACE_SURFACE = 1
MAX_FOOTPRINT_DISABLED = -1.0
VARIANT_ONE_BOUNCE = 0
VARIANT_RECURSIVE = 1

OK, ERROR_INVALID_ARGUMENT, ERROR_NO_DEVICE, ERROR_CUDA, ERROR_OUT_OF_MEMORY, ERROR_NO_VOLUME, ERROR_CORRUPT_VOLUME = range(7)


class CubiquityError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cubiquity_b200 error %d: %s" % (code, message))
        self.code = code


class Camera(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("forward", C.c_double * 3), ("up", C.c_double * 3),
                ("right", C.c_double * 3), ("scale", C.c_float), ("pad", C.c_float)]


class MeshInfo(C.Structure):
    _fields_ = [("lower", C.c_float * 3), ("upper", C.c_float * 3), ("is_closed", C.c_uint32), ("is_inside_out", C.c_uint32)]


class PtParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("bounces", C.c_uint32),
                ("variant", C.c_uint32), ("include_sun", C.c_uint32), ("include_sky", C.c_uint32),
                ("add_noise", C.c_uint32), ("max_footprint", C.c_float), ("frame_id", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("band_count", C.c_uint32), ("band_index", C.c_uint32),
                ("tile_group_count", C.c_uint32), ("tile_group_index", C.c_uint32)]


def pt_params(width, height, spp=1, bounces=1, variant=VARIANT_ONE_BOUNCE, include_sun=True, include_sky=True,
              add_noise=True, max_footprint=0.0035, frame_id=0, rect=None, bands=None, tile_group=None):
    """Defaults are PathtracingDemo's (reference pathtracing_demo.h:80-84). bands = (count, index) renders only
    the 64-row bands b of the rectangle with b % count == index (round-robin tile sharding); tile_group = (count, index)
    only the 64x64-pixel tiles of that group of the GLSL viewer's progressive schedule."""
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, width, height)
    bc, bi = bands if bands is not None else (1, 0)
    tc, ti = tile_group if tile_group is not None else (0, 0)
    return PtParams(width, height, spp, bounces, variant, int(include_sun), int(include_sky), int(add_noise),
                    max_footprint, frame_id, x0, y0, x1, y1, bc, bi, tc, ti)


EXPORTS = [
    "cbq_create", "cbq_destroy", "cbq_last_error", "cbq_device_count", "cbq_synchronize",
    "cbq_upload", "cbq_update", "cbq_bake", "cbq_build_dense", "cbq_build_dense_device", "cbq_fill_sphere", "cbq_set_root", "cbq_set_colours", "cbq_get_subdags", "cbq_find_subdags",
    "cbq_download_nodes", "cbq_node_count",
    "cbq_upload_device", "cbq_update_device",
    "cbq_trace", "cbq_trace_device", "cbq_trace_compact", "cbq_trace_compact_device", "cbq_expand_hits", "cbq_camera_from_pose", "cbq_primary_rays_device", "cbq_primary_rays_tiled_device", "cbq_random_rays_device",
    "cbq_raycast_frame_device",
    "cbq_render", "cbq_render_device", "cbq_rng_points_device",
    "cbq_progressive_pass_device", "cbq_normalise_device", "cbq_blur_device",
    "cbq_mesh_analyse", "cbq_voxelize",
    "cbq_dag_load", "cbq_dag_free", "cbq_dag_save", "cbq_upload_dag", "cbq_set_log_callback",
    "cbq_shared_alloc", "cbq_shared_open", "cbq_shared_close", "cbq_shared_free", "cbq_copy_device",
    "cbq_host_alloc", "cbq_host_free", "cbq_set_option", "cbq_get_option", "cbq_get_counter", "cbq_reset_counters",
    "cbq_editable_create", "cbq_editable_destroy", "cbq_editable_checkpoint", "cbq_editable_undo", "cbq_editable_redo",
    "cbq_editable_fill_sphere", "cbq_editable_nodes", "cbq_editable_root", "cbq_editable_shared_end", "cbq_editable_sync",
]

_lib = None


def library_path():
    return _build.LIB


def load_library():
    """Loads the CUDA library, building it in-tree if needed. Raises if it cannot -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CBQ_LIBRARY") or _build.LIB      # CBQ_LIBRARY: an experimental variant for A/B runs
    if path == _build.LIB and not os.path.exists(path):
        _build.build_library()
    L = C.CDLL(path)
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    L.cbq_last_error.restype = C.c_char_p
    L.cbq_create.argtypes = [i32, C.POINTER(vp)]
    L.cbq_destroy.argtypes = [vp]
    L.cbq_destroy.restype = None
    L.cbq_synchronize.argtypes = [vp]
    L.cbq_upload.argtypes = [vp, vp, u64, u32, vp]
    L.cbq_update.argtypes = [vp, vp, u64, u64, u32]
    L.cbq_bake.argtypes = [vp, C.POINTER(u64), C.POINTER(u32)]
    L.cbq_fill_sphere.argtypes = [vp, f32, f32, f32, f32, C.c_uint8, C.POINTER(u32), C.POINTER(u64)]
    L.cbq_set_root.argtypes = [vp, u32]
    L.cbq_build_dense.argtypes = [vp, vp, u32, C.POINTER(C.c_int32), vp, C.POINTER(u64), C.POINTER(u32)]
    L.cbq_build_dense_device.argtypes = [vp, vp, u32, C.POINTER(C.c_int32), vp, C.POINTER(u64), C.POINTER(u32)]
    L.cbq_set_colours.argtypes = [vp, vp]
    L.cbq_get_subdags.argtypes = [vp, vp]
    L.cbq_find_subdags.argtypes = [vp, u64, u32, vp]
    L.cbq_download_nodes.argtypes = [vp, u64, u64, vp]
    L.cbq_node_count.argtypes = [vp, C.POINTER(u64)]
    L.cbq_trace.argtypes = [vp, vp, u64, u32, f32, vp]
    L.cbq_trace_device.argtypes = [vp, vp, u64, u32, f32, vp, vp]
    L.cbq_trace_compact.argtypes = [vp, vp, u64, u32, f32, vp]
    L.cbq_trace_compact_device.argtypes = [vp, vp, u64, u32, f32, vp, vp]
    L.cbq_expand_hits.argtypes = [vp, vp, u64, vp, i32]
    L.cbq_upload_device.argtypes = [vp, vp, u64, u32, vp, vp]
    L.cbq_update_device.argtypes = [vp, vp, u64, u64, u32, vp]
    L.cbq_camera_from_pose.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.POINTER(Camera)]
    L.cbq_primary_rays_device.argtypes = [vp, C.POINTER(Camera), u32, u32, vp, vp]
    L.cbq_primary_rays_tiled_device.argtypes = [vp, C.POINTER(Camera), u32, u32, vp, vp, vp]
    L.cbq_random_rays_device.argtypes = [vp, u64, C.POINTER(C.c_float), C.POINTER(C.c_float), u64, vp, vp]
    L.cbq_raycast_frame_device.argtypes = [vp, C.POINTER(Camera), u32, u32, u32, f32, vp, vp]
    L.cbq_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(PtParams), vp]
    L.cbq_render_device.argtypes = [vp, C.POINTER(Camera), C.POINTER(PtParams), vp, vp]
    L.cbq_rng_points_device.argtypes = [vp, vp, u64, i32, vp, vp, vp]
    L.cbq_progressive_pass_device.argtypes = [vp, C.POINTER(Camera), C.POINTER(PtParams), u32, vp, vp]
    L.cbq_normalise_device.argtypes = [vp, vp, u32, u32, vp, vp]
    L.cbq_blur_device.argtypes = [vp, vp, u32, u32, vp, i32, vp]
    L.cbq_mesh_analyse.argtypes = [vp, u64, C.POINTER(MeshInfo)]
    L.cbq_voxelize.argtypes = [vp, vp, vp, u64, C.c_uint8, C.c_uint8, i32, u32, C.POINTER(C.c_int32), vp, C.POINTER(MeshInfo), C.POINTER(u64), C.POINTER(u32)]
    L.cbq_dag_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(u32)), C.POINTER(u64), C.POINTER(u32)]
    L.cbq_dag_free.argtypes = [C.POINTER(u32)]
    L.cbq_dag_free.restype = None
    L.cbq_dag_save.argtypes = [C.c_char_p, vp, u64, u32]
    L.cbq_upload_dag.argtypes = [vp, C.c_char_p, vp]
    L.cbq_set_log_callback.argtypes = [vp]
    L.cbq_set_log_callback.restype = None
    L.cbq_dag_error.restype = C.c_char_p
    L.cbq_shared_alloc.argtypes = [vp, u64, C.POINTER(vp), vp]
    L.cbq_shared_open.argtypes = [vp, vp, u64, C.POINTER(vp)]
    L.cbq_shared_close.argtypes = [vp, vp]
    L.cbq_shared_free.argtypes = [vp, vp]
    L.cbq_copy_device.argtypes = [vp, vp, vp, u64, vp]
    L.cbq_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.cbq_host_free.argtypes = [vp]
    L.cbq_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.cbq_get_option.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    L.cbq_get_counter.argtypes = [vp, C.c_char_p, C.POINTER(u64)]
    L.cbq_reset_counters.argtypes = [vp]
    L.cbq_editable_create.argtypes = [vp, u64, u32, C.POINTER(vp)]
    L.cbq_editable_destroy.restype = None
    L.cbq_editable_destroy.argtypes = [vp]
    for f in ("cbq_editable_checkpoint", "cbq_editable_undo", "cbq_editable_redo"):
        getattr(L, f).argtypes = [vp]
    L.cbq_editable_fill_sphere.argtypes = [vp, f32, f32, f32, f32, C.c_uint8]
    L.cbq_editable_nodes.restype = C.POINTER(u32)
    L.cbq_editable_nodes.argtypes = [vp, C.POINTER(u64)]
    L.cbq_editable_root.restype = u32
    L.cbq_editable_root.argtypes = [vp]
    L.cbq_editable_shared_end.restype = u64
    L.cbq_editable_shared_end.argtypes = [vp]
    L.cbq_editable_sync.argtypes = [vp, vp, i32, vp]
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise CubiquityError(rc, load_library().cbq_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the handle that names the default stream explicitly


def _stream(stream):
    """None -> NULL (the context's own stream). 0 is torch's default stream: pass cudaStreamLegacy so the
    work is really ordered on it (a NULL would select the context's private non-blocking stream)."""
    if stream is None:
        return None
    return C.c_void_p(int(stream) if int(stream) != 0 else CUDA_STREAM_LEGACY)


def device_count():
    return int(load_library().cbq_device_count())


def find_subdags(nodes, root):
    """findSubDAGs (reference raytracing.cpp:89-97) on the host; needs no device."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
    out = np.zeros(8, dtype=SUBDAG_DTYPE)
    _check(load_library().cbq_find_subdags(_ptr(nodes), len(nodes), int(root), _ptr(out)))
    return out


def camera_from_pose(position, pitch, yaw, fov_degrees=60.0):
    """Camera::forward/right/up + fov scale (reference camera.cpp:24,40-66). Host only."""
    cam = Camera()
    pos = (C.c_double * 3)(*[float(v) for v in position])
    _check(load_library().cbq_camera_from_pose(pos, float(pitch), float(yaw), float(fov_degrees), C.byref(cam)))
    return cam


def default_camera(lower, upper):
    """The viewer's start pose for a solid object (reference viewer.cpp:71-79): centred in x, back and
    up by half the bounding-box diagonal, looking down 45 degrees."""
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    centre = (lower + upper) * 0.5
    half_diag = float(np.sqrt(((upper - lower) ** 2).sum())) * 0.5
    pi_f = float(np.float32(3.14159265358979))
    return camera_from_pose([centre[0], centre[1] - half_diag, centre[2] + half_diag], -(pi_f / 4.0), 0.0)


class PinnedArray:
    """A numpy view over cudaHostAlloc memory (cbq_host_alloc)."""

    def __init__(self, count, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(count) * self.dtype.itemsize
        p = C.c_void_p()
        _check(load_library().cbq_host_alloc(C.byref(p), self.nbytes))
        self.ptr = p
        buf = (C.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(count))

    def free(self):
        if self.ptr:
            self.array = None
            load_library().cbq_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def Scene(kind, size_log2, seed=1):
    """A procedural test / benchmark volume. The generator is not part of this library: scenes/ (repository root)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import scenes
    return scenes.Scene(kind, size_log2, seed)


class Editable:
    """Host-side copy-on-write edits of a node array (csrc/edit.cpp): the reference's checkpoint / sphere brush /
    undo / redo, producing the same array the reference would, for the edit -> delta re-upload protocol."""

    def __init__(self, nodes, root):
        self.L = load_library()
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
        h = C.c_void_p()
        _check(self.L.cbq_editable_create(_ptr(nodes), len(nodes), int(root), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.L.cbq_editable_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def checkpoint(self):
        _check(self.L.cbq_editable_checkpoint(self._h))

    def undo(self):
        _check(self.L.cbq_editable_undo(self._h))

    def redo(self):
        _check(self.L.cbq_editable_redo(self._h))

    def fill_sphere(self, x, y, z, radius, material):
        _check(self.L.cbq_editable_fill_sphere(self._h, float(x), float(y), float(z), float(radius), int(material)))

    def nodes(self):
        """A VIEW of the current array (valid until the next edit), shape (n, 8)."""
        n = C.c_uint64()
        p = self.L.cbq_editable_nodes(self._h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(int(n.value), 8))

    def root(self):
        return int(self.L.cbq_editable_root(self._h))

    def shared_end(self):
        return int(self.L.cbq_editable_shared_end(self._h))

    def sync(self, ctx, first_upload=False, colours=None):
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        _check(self.L.cbq_editable_sync(self._h, ctx._h, int(bool(first_upload)), _ptr(colours)))


def mesh_analyse(triangles):
    """cbq_mesh_analyse: Mesh::build's verdict (bounds, closed, inside-out) on (n, 9) float32 triangles. Host only."""
    tris = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 9)
    info = MeshInfo()
    _check(load_library().cbq_mesh_analyse(_ptr(tris), len(tris), C.byref(info)))
    return info


def dag_load(path):
    """cbq_dag_load: (nodes (n, 8) uint32 incl. the 256 material nodes, root) from a reference .dag file, validated."""
    L = load_library()
    p = C.POINTER(C.c_uint32)()
    n = C.c_uint64()
    root = C.c_uint32()
    rc = L.cbq_dag_load(os.fsencode(path), C.byref(p), C.byref(n), C.byref(root))
    if rc != OK:
        raise CubiquityError(rc, L.cbq_dag_error().decode("utf-8", "replace"))
    try:
        nodes = np.ctypeslib.as_array(p, shape=(int(n.value), 8)).copy()
    finally:
        L.cbq_dag_free(p)
    return nodes, int(root.value)


def dag_save(path, nodes, root):
    L = load_library()
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
    rc = L.cbq_dag_save(os.fsencode(path), _ptr(nodes), len(nodes), int(root))
    if rc != OK:
        raise CubiquityError(rc, L.cbq_dag_error().decode("utf-8", "replace"))


_log_handler = None


def set_log_callback(fn):
    """Failed calls also report their message to fn(str) (reference MessageHandlerPtr, base.h:102-107); None = off."""
    global _log_handler
    L = load_library()
    if fn is None:
        _log_handler = None
        L.cbq_set_log_callback(None)
        return
    _log_handler = C.CFUNCTYPE(None, C.c_char_p)(lambda m: fn(m.decode("utf-8", "replace")))
    L.cbq_set_log_callback(C.cast(_log_handler, C.c_void_p))


def expand_hits(rays, compact, out=None, threads=0):
    """cbq_expand_hits: COMPACT_DTYPE records + their rays -> the HIT_DTYPE records cbq_trace writes (host only)."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    compact = np.ascontiguousarray(compact, dtype=COMPACT_DTYPE)
    assert len(rays) == len(compact)
    hits = out if out is not None else np.zeros(len(rays), dtype=HIT_DTYPE)
    _check(load_library().cbq_expand_hits(_ptr(rays), _ptr(compact), len(rays), _ptr(hits), int(threads)))
    return hits


class Context:
    """One GPU. Mirrors the reference call sites: upload the Volume's node array, then intersect_volume."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        _check(self.L.cbq_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self.L.cbq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- volume ------------------------------------------------------------------------------
    def upload(self, nodes, root, colours=None):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        _check(self.L.cbq_upload(self._h, _ptr(nodes), len(nodes), int(root), _ptr(colours)))

    def update(self, nodes, dirty_begin, root):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
        _check(self.L.cbq_update(self._h, _ptr(nodes), int(dirty_begin), len(nodes), int(root)))

    def upload_device(self, d_nodes, node_count, root, colours=None, stream=None):
        """cbq_upload for a node array already in device memory (d_nodes: device pointer to node_count x 8 u32)."""
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        _check(self.L.cbq_upload_device(self._h, C.c_void_p(int(d_nodes)), int(node_count), int(root), _ptr(colours), _stream(stream)))

    def update_device(self, d_tail, dirty_begin, node_count, root, stream=None):
        """cbq_update with the dirty tail (nodes [dirty_begin, node_count), tail only) in device memory."""
        _check(self.L.cbq_update_device(self._h, C.c_void_p(int(d_tail)) if d_tail else None, int(dirty_begin), int(node_count), int(root), _stream(stream)))

    def bake(self):
        """Volume::bake (reference storage.cpp:388-395) on the device copy; returns (node_count, root) of the merged
        array (download_nodes() reads it back). Same canonical DAG as the reference's merge, different node order."""
        n = C.c_uint64()
        root = C.c_uint32()
        _check(self.L.cbq_bake(self._h, C.byref(n), C.byref(root)))
        return int(n.value), int(root.value)

    def fill_sphere(self, x, y, z, radius, material):
        """checkpoint() + fillBrush(SphereBrush) (reference viewer.cpp:165-168) on the device copy; returns
        (new_root, node_count). Earlier roots stay valid: set_root() is undo / redo."""
        root = C.c_uint32()
        n = C.c_uint64()
        _check(self.L.cbq_fill_sphere(self._h, float(x), float(y), float(z), float(radius), int(material), C.byref(root), C.byref(n)))
        return int(root.value), int(n.value)

    def set_root(self, root):
        _check(self.L.cbq_set_root(self._h, int(root)))

    def build_dense(self, voxels, origin=(0, 0, 0), colours=None, device_ptr=None, size_log2=None):
        """setVoxel for every voxel + bake (reference storage.cpp:388-438), on the device. voxels: uint8 array
        [z, y, x] with side 2^k, or a device pointer (device_ptr, size_log2). Returns (node_count, root)."""
        n = C.c_uint64()
        root = C.c_uint32()
        org = (C.c_int32 * 3)(*[int(v) for v in origin])
        col = None
        if colours is not None:
            col = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        if device_ptr is not None:
            _check(self.L.cbq_build_dense_device(self._h, C.c_void_p(int(device_ptr)), int(size_log2), org, _ptr(col) if col is not None else None,
                                                 C.byref(n), C.byref(root)))
        else:
            voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
            side = voxels.shape[0]
            k = int(side).bit_length() - 1
            if voxels.shape != (side, side, side) or (1 << k) != side:
                raise ValueError("voxels must be a cube with a power-of-two side")
            _check(self.L.cbq_build_dense(self._h, _ptr(voxels), k, org, _ptr(col) if col is not None else None, C.byref(n), C.byref(root)))
        return int(n.value), int(root.value)

    def subdags(self):
        out = np.zeros(8, dtype=SUBDAG_DTYPE)
        _check(self.L.cbq_get_subdags(self._h, _ptr(out)))
        return out

    def node_count(self):
        n = C.c_uint64()
        _check(self.L.cbq_node_count(self._h, C.byref(n)))
        return int(n.value)

    def download_nodes(self, begin=0, count=None):
        if count is None:
            count = self.node_count() - begin
        out = np.zeros((count, 8), dtype=np.uint32)
        _check(self.L.cbq_download_nodes(self._h, int(begin), int(count), _ptr(out)))
        return out

    # -- ray cast ----------------------------------------------------------------------------
    def intersect_volume(self, rays, compute_surface_properties=True, max_footprint=MAX_FOOTPRINT_DISABLED, out=None):
        """Batch of Cubiquity::intersectVolume (reference raytracing.h:72-75). Host arrays in and out."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = out if out is not None else np.zeros(len(rays), dtype=HIT_DTYPE)
        flags = TRACE_SURFACE if compute_surface_properties else 0
        _check(self.L.cbq_trace(self._h, _ptr(rays), len(rays), flags, float(max_footprint), _ptr(hits)))
        return hits

    def intersect_volume_compact(self, rays, compute_surface_properties=True, max_footprint=MAX_FOOTPRINT_DISABLED, out=None):
        """intersect_volume with 8-byte results (COMPACT_DTYPE); expand_hits() widens them to HIT_DTYPE on the host."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = out if out is not None else np.zeros(len(rays), dtype=COMPACT_DTYPE)
        flags = TRACE_SURFACE if compute_surface_properties else 0
        _check(self.L.cbq_trace_compact(self._h, _ptr(rays), len(rays), flags, float(max_footprint), _ptr(hits)))
        return hits

    def trace_compact_device(self, d_rays, n, d_hits, compute_surface_properties=True,
                             max_footprint=MAX_FOOTPRINT_DISABLED, stream=None):
        flags = TRACE_SURFACE if compute_surface_properties else 0
        _check(self.L.cbq_trace_compact_device(self._h, C.c_void_p(int(d_rays)), int(n), flags, float(max_footprint),
                                               C.c_void_p(int(d_hits)), _stream(stream)))

    def trace_device(self, d_rays, n, d_hits, compute_surface_properties=True,
                     max_footprint=MAX_FOOTPRINT_DISABLED, stream=None):
        flags = TRACE_SURFACE if compute_surface_properties else 0
        _check(self.L.cbq_trace_device(self._h, C.c_void_p(int(d_rays)), int(n), flags, float(max_footprint),
                                       C.c_void_p(int(d_hits)), _stream(stream)))

    def primary_rays_device(self, cam, width, height, d_rays, stream=None):
        _check(self.L.cbq_primary_rays_device(self._h, C.byref(cam), int(width), int(height), C.c_void_p(int(d_rays)),
                                              _stream(stream)))

    def primary_rays_tiled_device(self, cam, width, height, d_rays, d_pixel_of=None, stream=None):
        _check(self.L.cbq_primary_rays_tiled_device(self._h, C.byref(cam), int(width), int(height), C.c_void_p(int(d_rays)),
                                                    C.c_void_p(int(d_pixel_of)) if d_pixel_of else None, _stream(stream)))

    def random_rays_device(self, seed, lower, upper, n, d_rays, stream=None):
        lo = (C.c_float * 3)(*[float(v) for v in lower])
        hi = (C.c_float * 3)(*[float(v) for v in upper])
        _check(self.L.cbq_random_rays_device(self._h, int(seed), lo, hi, int(n), C.c_void_p(int(d_rays)), _stream(stream)))

    def raycast_frame_device(self, cam, width, height, d_hits, compute_surface_properties=True,
                             max_footprint=MAX_FOOTPRINT_DISABLED, stream=None):
        flags = TRACE_SURFACE if compute_surface_properties else 0
        _check(self.L.cbq_raycast_frame_device(self._h, C.byref(cam), int(width), int(height), flags,
                                               float(max_footprint), C.c_void_p(int(d_hits)),
                                               _stream(stream)))

    # -- path tracer -------------------------------------------------------------------------
    def render(self, cam, params, accum=None):
        """PathtracingDemo::raytrace (reference pathtracing_demo.cpp:214-229): adds params.spp samples."""
        if accum is None:
            accum = np.zeros((params.height, params.width, 3), dtype=np.float32)
        assert accum.dtype == np.float32 and accum.flags["C_CONTIGUOUS"]
        _check(self.L.cbq_render(self._h, C.byref(cam), C.byref(params), _ptr(accum)))
        return accum

    def render_device(self, cam, params, d_accum, stream=None):
        _check(self.L.cbq_render_device(self._h, C.byref(cam), C.byref(params), C.c_void_p(int(d_accum)),
                                        _stream(stream)))

    def voxelize(self, triangles, materials, fill, size_log2, origin, background=0, thin=False, colours=None):
        """voxelize(volume, mesh, fill, background) (reference voxelization.cpp:692-744) on the device; the result becomes the
        context's volume. Returns (node_count, root, MeshInfo)."""
        tris = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 9)
        mats = np.ascontiguousarray(materials, dtype=np.uint8)
        assert len(mats) == len(tris)
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        org = (C.c_int32 * 3)(*[int(v) for v in origin])
        info = MeshInfo()
        n = C.c_uint64()
        root = C.c_uint32()
        _check(self.L.cbq_voxelize(self._h, _ptr(tris), _ptr(mats), len(tris), int(fill), int(background), int(bool(thin)), int(size_log2), org,
                                   _ptr(colours), C.byref(info), C.byref(n), C.byref(root)))
        return int(n.value), int(root.value), info

    def upload_dag(self, path, colours=None):
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3)
        _check(self.L.cbq_upload_dag(self._h, os.fsencode(path), _ptr(colours)))

    def progressive_pass_device(self, cam, params, frame, d_rgba, stream=None):
        """GPUPathtracingViewer's progressive frame `frame`: one tile group gets params.spp more samples, added into RGBA."""
        _check(self.L.cbq_progressive_pass_device(self._h, C.byref(cam), C.byref(params), int(frame), C.c_void_p(int(d_rgba)), _stream(stream)))

    def normalise_device(self, d_rgba, width, height, d_rgb, stream=None):
        _check(self.L.cbq_normalise_device(self._h, C.c_void_p(int(d_rgba)), int(width), int(height), C.c_void_p(int(d_rgb)), _stream(stream)))

    def blur_device(self, d_rgba, width, height, d_scratch, passes=1, stream=None):
        _check(self.L.cbq_blur_device(self._h, C.c_void_p(int(d_rgba)), int(width), int(height), C.c_void_p(int(d_scratch)), int(passes), _stream(stream)))

    def rng_points_device(self, d_seeds, n, draws, d_points, d_states, stream=None):
        _check(self.L.cbq_rng_points_device(self._h, C.c_void_p(int(d_seeds)), int(n), int(draws), C.c_void_p(int(d_points)),
                                            C.c_void_p(int(d_states)), _stream(stream)))

    # -- multi-GPU: one result buffer on one GPU, addressed by every process over NVLink --------
    def shared_alloc(self, nbytes):
        """(device pointer, 64-byte IPC handle as bytes) of a zero-filled buffer other processes can open."""
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        _check(self.L.cbq_shared_alloc(self._h, int(nbytes), C.byref(p), h))
        return int(p.value), bytes(h)

    def shared_open(self, handle, nbytes):
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(bytes(handle))
        _check(self.L.cbq_shared_open(self._h, h, int(nbytes), C.byref(p)))
        return int(p.value)

    def shared_close(self, ptr):
        _check(self.L.cbq_shared_close(self._h, C.c_void_p(int(ptr))))

    def shared_free(self, ptr):
        _check(self.L.cbq_shared_free(self._h, C.c_void_p(int(ptr))))

    def copy_device(self, dst, src, nbytes, stream=None):
        _check(self.L.cbq_copy_device(self._h, C.c_void_p(int(dst)), C.c_void_p(int(src)), int(nbytes), _stream(stream)))

    # -- misc --------------------------------------------------------------------------------
    def synchronize(self):
        _check(self.L.cbq_synchronize(self._h))

    def set_option(self, key, value):
        _check(self.L.cbq_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int64()
        _check(self.L.cbq_get_option(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def counter(self, key):
        v = C.c_uint64()
        _check(self.L.cbq_get_counter(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def reset_counters(self):
        _check(self.L.cbq_reset_counters(self._h))

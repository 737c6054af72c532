"""TEST INFRASTRUCTURE ONLY. Runs the reference's voxelize() (oracle/ref_shim.cpp:ref_voxelize) in a process of its own and
saves the volume as a .dag: inside a process that has numpy's OpenBLAS loaded the reference's voxeliser crashes (it does not
in a plain process), so pyoracle.RefVolume.voxelize shells out to this script, which imports nothing but ctypes.

    python ref_voxelize_helper.py <libcbq_ref.so> <triangles.f32> <materials.u8> <fill> <background> <thin> <out.dag>
prints: is_closed is_inside_out seconds"""
import ctypes as C
import sys

so, tri_path, mat_path, fill, background, thin, out = sys.argv[1:8]
L = C.CDLL(so)
L.ref_volume_new.restype = C.c_void_p
L.ref_voxelize.restype = C.c_double
L.ref_voxelize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8, C.c_uint8, C.c_int, C.c_void_p]
L.ref_volume_save.argtypes = [C.c_void_p, C.c_char_p]
tris = open(tri_path, "rb").read()
mats = open(mat_path, "rb").read()
n = len(mats)
assert len(tris) == n * 36
v = L.ref_volume_new()
flags = (C.c_uint32 * 2)()
secs = L.ref_voxelize(v, tris, mats, n, int(fill), int(background), int(thin), flags)
L.ref_volume_save(v, out.encode())
print(flags[0], flags[1], secs)

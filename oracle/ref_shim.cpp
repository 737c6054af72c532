// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" shim around the UNMODIFIED reference library, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile). It lets the
// Python tests and bench.py's cpu_baseline / --impl reference arm drive the
// reference's own code:
//   Cubiquity::Volume            (src/library/storage.h:135-197)
//   Cubiquity::findSubDAGs       (src/library/raytracing.h:69)
//   Cubiquity::intersectVolume   (src/library/raytracing.h:72-75, six-float overload)
//   Cubiquity::traceRayRef       (src/library/raytracing.h:78-80, brute-force checker)
//   Cubiquity::fillBrush         (src/library/voxelization.h:128)
// Records use the same 24-byte ray / 40-byte hit layout as include/cubiquity_b200.h
// so results can be compared byte for byte.
#include "base.h"
#include "geometry.h"
#include "storage.h"
#include "utility.h"
#include "raytracing.h"
#include "voxelization.h"
#include "cubiquity.h"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace Cubiquity;

namespace {

struct RefVolume {
	Volume volume;
	SubDAGArray subdags;
	bool subdagsValid = false;
	void refresh() {
		subdags = findSubDAGs(Internals::getNodes(volume), Internals::getRootNodeIndex(volume));
		subdagsValid = true;
	}
};

struct RayRec { float o[3]; float d[3]; };
struct HitRec {
	uint32_t hit;
	float distance;
	uint32_t material;
	float position[3];
	float normal[3];
	uint32_t pad;
};
static_assert(sizeof(RayRec) == 24, "ray record");
static_assert(sizeof(HitRec) == 40, "hit record");

inline void storeHit(const RayVolumeIntersection& r, HitRec& h)
{
	h.hit = r.hit ? 1u : 0u;
	h.distance = static_cast<float>(r.distance); // value is a float held in a double (raytracing.cpp:305)
	h.material = r.material;
	h.position[0] = r.position.x; h.position[1] = r.position.y; h.position[2] = r.position.z;
	h.normal[0] = r.normal.x; h.normal[1] = r.normal.y; h.normal[2] = r.normal.z;
	h.pad = 0;
}

void traceRange(RefVolume* rv, const RayRec* rays, uint64_t begin, uint64_t end,
	bool surf, float maxFootprint, HitRec* out)
{
	for (uint64_t i = begin; i < end; i++) {
		const RayRec& r = rays[i];
		RayVolumeIntersection isect = intersectVolume(rv->volume, rv->subdags,
			r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], surf, maxFootprint);
		if (out) storeHit(isect, out[i]);
	}
}

} // namespace

extern "C" {

void* ref_volume_new() { return new RefVolume(); }
void ref_volume_free(void* v) { delete static_cast<RefVolume*>(v); }

int ref_volume_load(void* v, const char* path)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	bool ok = rv->volume.load(path);
	if (ok) rv->refresh();
	return ok ? 0 : 1;
}

// NB Volume::save() bakes first (storage.cpp:530-542), which relocates every node.
void ref_volume_save(void* v, const char* path)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	rv->volume.save(path);
	rv->refresh();
}

void ref_volume_set_voxel(void* v, int32_t x, int32_t y, int32_t z, uint8_t m)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	rv->volume.setVoxel(x, y, z, m);
	rv->subdagsValid = false;
}

// xyzm: n records of 4 x int32 (x, y, z, material)
void ref_volume_set_voxels(void* v, const int32_t* xyzm, uint64_t n)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	for (uint64_t i = 0; i < n; i++) {
		rv->volume.setVoxel(xyzm[4*i], xyzm[4*i+1], xyzm[4*i+2], static_cast<uint8_t>(xyzm[4*i+3]));
	}
	rv->subdagsValid = false;
}

uint8_t ref_volume_voxel(void* v, int32_t x, int32_t y, int32_t z)
{
	return static_cast<RefVolume*>(v)->volume.voxel(x, y, z);
}

void ref_volume_voxels(void* v, const int32_t* xyz, uint64_t n, uint8_t* out)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	for (uint64_t i = 0; i < n; i++) out[i] = rv->volume.voxel(xyz[3*i], xyz[3*i+1], xyz[3*i+2]);
}

void ref_volume_fill(void* v, uint8_t m)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	rv->volume.fill(m);
	rv->subdagsValid = false;
}

void ref_volume_bake(void* v)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	rv->volume.bake();
	rv->refresh();
}

void ref_volume_checkpoint(void* v) { static_cast<RefVolume*>(v)->volume.checkpoint(); static_cast<RefVolume*>(v)->subdagsValid = false; }
void ref_volume_undo(void* v) { static_cast<RefVolume*>(v)->volume.undo(); static_cast<RefVolume*>(v)->subdagsValid = false; }
void ref_volume_redo(void* v) { static_cast<RefVolume*>(v)->volume.redo(); static_cast<RefVolume*>(v)->subdagsValid = false; }

void ref_volume_fill_sphere(void* v, float x, float y, float z, float radius, uint8_t m)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	SphereBrush brush(x, y, z, radius);
	fillBrush(rv->volume, brush, m);
	rv->subdagsValid = false;
}

// Raw view of the node array exactly as the GLSL viewer uploads it
// (gpu_pathtracing_viewer.cpp:46-50): includes the 256 material nodes.
const uint32_t* ref_volume_nodes(void* v, uint64_t* nodeCountInclMaterials)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	const Internals::NodeStore& ns = Internals::getNodes(rv->volume);
	*nodeCountInclMaterials = ns.unsharedNodesEnd();
	return static_cast<const uint32_t*>(ns.rawBytesPtr());
}

uint32_t ref_volume_root(void* v) { return Internals::getRootNodeIndex(static_cast<RefVolume*>(v)->volume); }
uint32_t ref_volume_shared_end(void* v) { return Internals::getNodes(static_cast<RefVolume*>(v)->volume).sharedNodesEnd(); }

// out: 8 x 8 u32, the SubDAG struct verbatim (raytracing.h:57-65)
void ref_find_subdags(void* v, uint32_t* out)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	rv->refresh();
	static_assert(sizeof(SubDAG) == 32, "SubDAG is 32 bytes");
	std::memcpy(out, rv->subdags.data(), 8 * sizeof(SubDAG));
}

void ref_estimate_bounds(void* v, uint8_t* outside, int32_t* lowerUpper6)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	cubiquity_estimate_bounds(&rv->volume, outside, &lowerUpper6[0], &lowerUpper6[1], &lowerUpper6[2],
		&lowerUpper6[3], &lowerUpper6[4], &lowerUpper6[5]);
}

// Returns seconds spent inside the loop. threads <= 1 is the reference's own mode.
double ref_intersect_volume(void* v, const void* rays, uint64_t n, int surf, float maxFootprint,
	void* hitsOut, int threads)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	if (!rv->subdagsValid) rv->refresh();
	const RayRec* r = static_cast<const RayRec*>(rays);
	HitRec* out = static_cast<HitRec*>(hitsOut);
	auto t0 = std::chrono::steady_clock::now();
	if (threads <= 1) {
		traceRange(rv, r, 0, n, surf != 0, maxFootprint, out);
	} else {
		// intersectVolume is a pure function of const data so a static split is legal.
		std::vector<std::thread> pool;
		std::atomic<uint64_t> next(0);
		const uint64_t chunk = 4096;
		for (int t = 0; t < threads; t++) {
			pool.emplace_back([&]() {
				for (;;) {
					uint64_t b = next.fetch_add(chunk);
					if (b >= n) break;
					uint64_t e = b + chunk < n ? b + chunk : n;
					traceRange(rv, r, b, e, surf != 0, maxFootprint, out);
				}
			});
		}
		for (auto& th : pool) th.join();
	}
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// Brute-force checker (raytracing.cpp:480-535). Only hit / distance / material are defined.
void ref_trace_ray_ref(void* v, const void* rays, uint64_t n, void* hitsOut)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	const RayRec* r = static_cast<const RayRec*>(rays);
	HitRec* out = static_cast<HitRec*>(hitsOut);
	for (uint64_t i = 0; i < n; i++) {
		RayVolumeIntersection isect = traceRayRef(rv->volume,
			r[i].o[0], r[i].o[1], r[i].o[2], r[i].d[0], r[i].d[1], r[i].d[2]);
		std::memset(&out[i], 0, sizeof(HitRec));
		out[i].hit = isect.hit ? 1u : 0u;
		out[i].distance = isect.hit ? static_cast<float>(isect.distance) : 0.0f;
		out[i].material = isect.material;
	}
}

// Known-answer hashes of test_base.cpp:14-35, exposed so tests can pin the port's copies.
// The reference's mesh voxeliser: Mesh::addTriangle for every triangle, Mesh::build, voxelize (voxelization.cpp:692-823).
// flags[0] = isClosed, flags[1] = isInsideOut as Mesh::build decided them. Returns seconds spent in voxelize().
double ref_voxelize(void* v, const float* tris9, const uint8_t* mats, uint64_t n, uint8_t fill, uint8_t background, int thin, uint32_t* flags)
{
	RefVolume* rv = static_cast<RefVolume*>(v);
	Mesh mesh;
	for (uint64_t i = 0; i < n; i++) {
		const float* t = tris9 + 9 * i;
		mesh.addTriangle(Triangle(vec3f(t[0], t[1], t[2]), vec3f(t[3], t[4], t[5]), vec3f(t[6], t[7], t[8])), mats[i]);
	}
	mesh.isThin = thin != 0;
	mesh.build();
	if (flags) { flags[0] = mesh.isClosed ? 1u : 0u; flags[1] = mesh.isInsideOut ? 1u : 0u; }
	const auto t0 = std::chrono::steady_clock::now();
	voxelize(rv->volume, mesh, fill, background);
	const auto t1 = std::chrono::steady_clock::now();
	rv->subdagsValid = false;
	return std::chrono::duration<double>(t1 - t0).count();
}

uint64_t ref_bit_mix(uint64_t x) { return Internals::bit_mix(x); }
uint64_t ref_fnv1a(const void* data, int64_t len) { return Internals::fnv1a(data, len); }

} // extern "C"

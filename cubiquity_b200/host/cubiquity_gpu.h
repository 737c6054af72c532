// C++ host-side mirror of the reference's ray-cast interface, over the C ABI (include/cubiquity_b200.h).
//
// A maintainer of Cubiquity includes this header next to raytracing.h and swaps
//     Cubiquity::intersectVolume(volume, subDAGs, ox, oy, oz, dx, dy, dz, surf, maxFootprint)
// for
//     CubiquityGPU::intersectVolume(gpuVolume, ox, oy, oz, dx, dy, dz, surf, maxFootprint)
// with the same argument meaning, the same result struct and the same "never throws" contract
// (reference src/library/raytracing.h:48-55,69-76). The batch forms are what the GPU is for; the
// single-ray form exists so picking code (reference viewer.cpp:157-163) ports line for line.
//
// Header-only, no CUDA headers needed by the includer; link against libcubiquity_b200.so.
#ifndef CUBIQUITY_GPU_H
#define CUBIQUITY_GPU_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cubiquity_b200.h"

namespace CubiquityGPU {

// Field-for-field the reference's RayVolumeIntersection (raytracing.h:48-55), distance widened back to
// double exactly as the reference stores it (raytracing.cpp:305).
struct RayVolumeIntersection {
	bool hit;
	double distance;
	unsigned int material;
	float position[3];
	float normal[3];
};

typedef cbq_subdag SubDAG;                 // raytracing.h:57-65
struct SubDAGArray { SubDAG v[8]; const SubDAG& operator[](unsigned i) const { return v[i]; } };

constexpr float MAX_FOOTPRINT_DISABLED = -1.0f;   // raytracing.h:71

inline RayVolumeIntersection widen(const cbq_hit& h)
{
	RayVolumeIntersection r;
	r.hit = h.hit != 0;
	r.distance = static_cast<double>(h.distance);
	r.material = h.material;
	for (int a = 0; a < 3; a++) { r.position[a] = h.position[a]; r.normal[a] = h.normal[a]; }
	return r;
}

// The device-resident copy of a Cubiquity::Volume's node array. Owns a cbq_context.
class GpuVolume {
public:
	explicit GpuVolume(int device = 0) : mCtx(nullptr), mOk(cbq_create(device, &mCtx) == CBQ_OK) {}
	~GpuVolume() { cbq_destroy(mCtx); }
	GpuVolume(const GpuVolume&) = delete;
	GpuVolume& operator=(const GpuVolume&) = delete;

	bool ok() const { return mOk; }
	static std::string lastError() { return cbq_last_error(); }
	cbq_context* context() const { return mCtx; }

	// nodes = Internals::getNodes(volume).rawBytesPtr(), count = unsharedNodesEnd(),
	// root = Internals::getRootNodeIndex(volume)  (storage.h:87-102, storage.cpp:563-576).
	bool upload(const void* nodes, uint64_t nodeCount, uint32_t root, const float* coloursRgb = nullptr)
	{
		mSynced = nodeCount;
		return cbq_upload(mCtx, static_cast<const uint32_t*>(nodes), nodeCount, root, coloursRgb) == CBQ_OK;
	}

	// Call from onVolumeModified() (pathtracing_demo.cpp:335-341). sharedNodesEndAtLastSync is what
	// NodeStore::sharedNodesEnd() returned when the device was last brought up to date; pass the value
	// you remembered, then remember the current one.
	bool update(const void* nodes, uint64_t sharedNodesEndAtLastSync, uint64_t nodeCount, uint32_t root)
	{
		return cbq_update(mCtx, static_cast<const uint32_t*>(nodes), sharedNodesEndAtLastSync, nodeCount, root) == CBQ_OK;
	}

	// checkpoint() + fillBrush(volume, SphereBrush(centre, radius), material) (viewer.cpp:165-168) applied to the
	// device copy; root receives the new root (keep the old ones: setRoot(old) is undo, storage.cpp:373-385).
	bool fillSphere(float x, float y, float z, float radius, uint8_t material, uint32_t& root)
	{
		uint64_t count = 0;
		const bool ok = cbq_fill_sphere(mCtx, x, y, z, radius, material, &root, &count) == CBQ_OK;
		if (ok) mSynced = count;
		return ok;
	}
	bool setRoot(uint32_t root) { return cbq_set_root(mCtx, root) == CBQ_OK; }

	// Volume::bake() (storage.cpp:388-395) on the device copy. On success the device holds the merged DAG;
	// nodeCount/root describe it and download() reads it back, e.g. into a fresh Volume via the .dag layout.
	bool bake(uint64_t& nodeCount, uint32_t& root)
	{
		const bool ok = cbq_bake(mCtx, &nodeCount, &root) == CBQ_OK;
		if (ok) mSynced = nodeCount;
		return ok;
	}
	bool download(uint64_t begin, uint64_t count, void* nodesOut) { return cbq_download_nodes(mCtx, begin, count, static_cast<uint32_t*>(nodesOut)) == CBQ_OK; }

	// Volume::load(filename) + upload (storage.cpp:505-528), with the file checked against its header.
	bool load(const std::string& filename, const float* coloursRgb = nullptr) { return cbq_upload_dag(mCtx, filename.c_str(), coloursRgb) == CBQ_OK; }

	// voxelize(volume, mesh, fill, background) after mesh.build() (voxelization.cpp:692-744, 765-823): triangles = 9 floats each in
	// voxel coordinates, user order; the result becomes this volume. closed / insideOut report Mesh::build's verdict.
	bool voxelize(const std::vector<float>& triangles, const std::vector<uint8_t>& materials, uint8_t fill, bool thin,
		uint32_t sizeLog2, const int32_t origin[3], bool* closed = nullptr, bool* insideOut = nullptr, const float* coloursRgb = nullptr)
	{
		cbq_mesh_info info{};
		uint64_t count = 0; uint32_t root = 0;
		const bool ok = cbq_voxelize(mCtx, triangles.data(), materials.data(), materials.size(), fill, 0, thin ? 1 : 0, sizeLog2, origin,
			coloursRgb, &info, &count, &root) == CBQ_OK;
		if (closed) *closed = info.is_closed != 0;
		if (insideOut) *insideOut = info.is_inside_out != 0;
		if (ok) mSynced = count;
		return ok;
	}

	SubDAGArray subDAGs() const
	{
		SubDAGArray a{};
		cbq_get_subdags(mCtx, a.v);
		return a;
	}

private:
	cbq_context* mCtx;
	bool mOk;
	uint64_t mSynced = 0;
};

// findSubDAGs (raytracing.h:69) without a device.
inline bool findSubDAGs(const void* nodes, uint64_t nodeCount, uint32_t root, SubDAGArray& out)
{
	return cbq_find_subdags(static_cast<const uint32_t*>(nodes), nodeCount, root, out.v) == CBQ_OK;
}

// The six-float overload every reference call site uses (raytracing.h:72-75).
inline RayVolumeIntersection intersectVolume(const GpuVolume& volume,
	float ray_orig_x, float ray_orig_y, float ray_orig_z, float ray_dir_x, float ray_dir_y, float ray_dir_z,
	bool computeSurfaceProperties, float maxFootprint = MAX_FOOTPRINT_DISABLED)
{
	const cbq_ray r = { { ray_orig_x, ray_orig_y, ray_orig_z }, { ray_dir_x, ray_dir_y, ray_dir_z } };
	cbq_hit h{};
	if (cbq_trace(volume.context(), &r, 1, computeSurfaceProperties ? CBQ_TRACE_SURFACE : 0u, maxFootprint, &h) != CBQ_OK) {
		h = cbq_hit{};   // like the reference, report a miss rather than throw; GpuVolume::lastError() says why
	}
	return widen(h);
}

// Batch form: rays[i] -> out[i]. Returns false (and leaves `out` sized but zeroed) on error.
inline bool intersectVolume(const GpuVolume& volume, const std::vector<cbq_ray>& rays, bool computeSurfaceProperties,
	float maxFootprint, std::vector<cbq_hit>& out)
{
	out.assign(rays.size(), cbq_hit{});
	return cbq_trace(volume.context(), rays.data(), rays.size(), computeSurfaceProperties ? CBQ_TRACE_SURFACE : 0u,
		maxFootprint, out.data()) == CBQ_OK;
}

// Batch form with 8-byte results (24 + 8 bytes per ray over PCIe instead of 24 + 40): everything RayVolumeIntersection carries
// except position. expand() re-forms the full records, position = origin + dir * distance as raytracing.cpp:463-466 computes it.
inline bool intersectVolumeCompact(const GpuVolume& volume, const std::vector<cbq_ray>& rays, bool computeSurfaceProperties,
	float maxFootprint, std::vector<cbq_hit_compact>& out)
{
	out.assign(rays.size(), cbq_hit_compact{});
	return cbq_trace_compact(volume.context(), rays.data(), rays.size(), computeSurfaceProperties ? CBQ_TRACE_SURFACE : 0u,
		maxFootprint, out.data()) == CBQ_OK;
}

inline bool expand(const std::vector<cbq_ray>& rays, const std::vector<cbq_hit_compact>& compact, std::vector<cbq_hit>& out, int threads = 0)
{
	if (rays.size() != compact.size()) return false;
	out.assign(rays.size(), cbq_hit{});
	return cbq_expand_hits(rays.data(), compact.data(), rays.size(), out.data(), threads) == CBQ_OK;
}

} // namespace CubiquityGPU

#endif // CUBIQUITY_GPU_H

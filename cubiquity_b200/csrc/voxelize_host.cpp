// Host half of cbq_voxelize: the parts of the reference's voxeliser that are sequential by nature and tiny.
//   analyseMesh      Mesh::build (src/library/voxelization.cpp:765-823): bounds, and 100 winding-number samples from
//                    Box3fSampler (src/library/geometry.h:404-424: std::minstd_rand(0) + three
//                    std::uniform_real_distribution<double>) that decide isClosed / isInsideOut
//   splitTriangles   drawLargeTriangle (voxelization.cpp:488-512): triangles with a side longer than 16 are halved at the
//                    midpoint of their "longest" side, recursively, depth first, so the pieces keep the user's order
// Compiled with -ffp-contract=off. The per-voxel work is in voxelize_kernels.cu.
#include "../../include/cubiquity_b200.h"
#include "meshmath.cuh"

#include <cmath>
#include <limits>
#include <random>
#include <vector>

namespace cbq {

void meshBounds(const Tri* tris, uint64_t n, float lower[3], float upper[3])
{
	// Box::invalidate + accumulate (geometry.h:278-282, 318-322)
	float lo[3] = { std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max() };
	float hi[3] = { std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest() };
	for (uint64_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) {
		const float c[3] = { tris[i].v[k].x, tris[i].v[k].y, tris[i].v[k].z };
		for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
	}
	for (int a = 0; a < 3; a++) { lower[a] = lo[a]; upper[a] = hi[a]; }
}

void analyseMesh(const Tri* tris, uint64_t n, cbq_mesh_info* info)
{
	meshBounds(tris, n, info->lower, info->upper);
	bool allValid = true, anyPositive = false, anyNegative = false;
	const float tolerance = 0.1f;
	std::minstd_rand eng(0);
	std::uniform_real_distribution<> rx(info->lower[0], info->upper[0]), ry(info->lower[1], info->upper[1]), rz(info->lower[2], info->upper[2]);
	for (int i = 0; i < 100; i++) {
		const V3 q{ (float)rx(eng), (float)ry(eng), (float)rz(eng) };
		float sum = 0.0f;
		for (uint64_t t = 0; t < n; t++) sum += windingTerm(q, tris[t]);
		const float w = windingNormalise(sum), aw = std::abs(w);
		if (aw >= tolerance) {
			anyPositive |= w > 0.0f;
			anyNegative |= w < 0.0f;
			if (!(aw > (1.0f - tolerance))) { allValid = false; break; }
		}
	}
	info->is_closed = (allValid && !(anyNegative && anyPositive)) ? 1u : 0u;
	info->is_inside_out = (info->is_closed && anyNegative) ? 1u : 0u;
}

namespace {
void splitInto(const Tri& tri, uint8_t material, std::vector<Tri>& out, std::vector<uint8_t>& mats)
{
	auto side = [&](int i) { return length(tri.v[(i + 1) % 3] - tri.v[i]); };
	int longest = 0;
	if (side(1) > side(0)) longest = 1;
	if (side(2) > side(1)) longest = 2;          // compares with side 1, as the reference does
	if (side(longest) > 16.0f) {
		const V3 mid = (tri.v[longest] + tri.v[(longest + 1) % 3]) / 2.0f;
		for (int i = 0; i < 2; i++) {
			Tri half = tri;
			half.v[(longest + i) % 3] = mid;
			splitInto(half, material, out, mats);
		}
	} else {
		out.push_back(tri);
		mats.push_back(material);
	}
}
}

void splitTriangles(const Tri* tris, const uint8_t* materials, uint64_t n, std::vector<Tri>& out, std::vector<uint8_t>& mats)
{
	for (uint64_t i = 0; i < n; i++) splitInto(tris[i], materials[i], out, mats);
}

} // namespace cbq

extern "C" int cbq_mesh_analyse(const float* triangles, uint64_t triangle_count, cbq_mesh_info* info)
{
	if (!triangles || !info || triangle_count == 0) return CBQ_ERROR_INVALID_ARGUMENT;
	cbq::analyseMesh(reinterpret_cast<const cbq::Tri*>(triangles), triangle_count, info);
	return CBQ_OK;
}

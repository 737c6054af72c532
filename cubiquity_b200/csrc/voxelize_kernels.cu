// The reference's mesh voxeliser on the device: voxelize(volume, mesh, fill, background),
// src/library/voxelization.cpp:692-744, over a dense grid of 2^k voxels a side that cbq_build_dense then turns into the DAG.
//
// The reference works on its octree: (a) it scan-converts every triangle into a 6-separating shell of voxels with a
// checkerboard of materials (drawTriangles, thickness -1: :424-486), (b) collects every LEAF of the resulting octree that
// overlaps the shell's bounds -- single voxels next to the surface, larger and larger uniform cells away from it -- and
// classifies each by the generalized winding number at its centre (findNodes / classifyNodes, :564-644), (c) scan-converts
// the triangles again with thickness 1 and gives every non-background voxel within that distance the triangle's material,
// later triangles winning (:727-737). Here, same three steps, one kernel each:
//
//   shellKernel      one warp per (pre-split, <= 16 voxels long) triangle walks the triangle's voxel bounding box; mode 0
//                    writes the checkerboard, mode 1 records the LAST triangle within distance 1 (atomicMax of its order)
//   occupancy pyramid level j holds one byte per 2^j cell: "contains a non-background voxel". A cell is a leaf of the
//                    reference's octree iff it is empty and its parent is not -- or it is a single voxel of an occupied
//                    2x2x2 block. collectKernel lists those leaves (count pass, then emit pass).
//   classifyKernel   one thread per leaf: the flat winding-number sum over all triangles (staged through shared memory
//                    256 at a time, summed in mesh order), threshold |w| > 0.501 (isInside, :356-367)
//   fillKernel       one warp per leaf writes fill / background into its voxels
//   resolveKernel    material of the recorded triangle
//
// Bytes: per triangle ~6 k voxel tests of ~150 flops (compute-bound, tiny); classify is leaves x triangles x ~90 flops +
// 3 sqrt + 6 div + 1 atan2 (the dominant cost: compute-bound on the FP32 / SFU pipes, no HBM traffic to speak of: 36 B
// per triangle per CTA from L2); the pyramid and fill move ~2 B per voxel.
//
// The reference evaluates the winding number through a patch hierarchy (Algorithm 2 of Jacobson et al.), which is the
// same number up to float rounding; the threshold's 0.001 margin exists precisely to make the classification insensitive
// to that (voxelization.cpp:362-366). COMPILE WITH -fmad=false.
#include "cbq_internal.h"
#include "meshmath.cuh"

namespace cbq {

namespace {

struct Grid {
	uint8_t* voxels;      // [z][y][x], side S = 1 << log2
	uint32_t log2;
	int32_t ox, oy, oz;   // voxel (0,0,0) of the grid in volume coordinates
	__device__ __forceinline__ bool index(int x, int y, int z, size_t& i) const
	{
		const uint32_t S = 1u << log2;
		const uint32_t lx = (uint32_t)(x - ox), ly = (uint32_t)(y - oy), lz = (uint32_t)(z - oz);
		if (lx >= S || ly >= S || lz >= S) return false;
		i = ((size_t)lz << (2 * log2)) | ((size_t)ly << log2) | lx;
		return true;
	}
};

// mode 0: intersection-target test, write background + checkerboard + 1 (voxelization.cpp:715-719)
// mode 1: intersection-target test, record the triangle's order (an open mesh's shell, :739-743)
// mode 2: distance <= 1 and (voxel != background or thin), record the triangle's order (:727-737)
__global__ void __launch_bounds__(256)
shellKernel(const Tri* __restrict__ tris, uint32_t count, Grid g, int mode, uint8_t background, int thin, unsigned int* __restrict__ order)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < count; t += warps) {
		const Tri tri = tris[t];
		const float dil = (mode == 2) ? 1.0f : 0.5f;
		float lo[3], hi[3];
		{
			const float xs[3] = { tri.v[0].x, tri.v[1].x, tri.v[2].x }, ys[3] = { tri.v[0].y, tri.v[1].y, tri.v[2].y }, zs[3] = { tri.v[0].z, tri.v[1].z, tri.v[2].z };
			lo[0] = fminf(fminf(xs[0], xs[1]), xs[2]) - dil; hi[0] = fmaxf(fmaxf(xs[0], xs[1]), xs[2]) + dil;
			lo[1] = fminf(fminf(ys[0], ys[1]), ys[2]) - dil; hi[1] = fmaxf(fmaxf(ys[0], ys[1]), ys[2]) + dil;
			lo[2] = fminf(fminf(zs[0], zs[1]), zs[2]) - dil; hi[2] = fmaxf(fmaxf(zs[0], zs[1]), zs[2]) + dil;
		}
		// Shrink to the integer positions inside the float bounds (:436-439)
		const int x0 = (int)ceilf(lo[0]), y0 = (int)ceilf(lo[1]), z0 = (int)ceilf(lo[2]);
		const int x1 = (int)floorf(hi[0]), y1 = (int)floorf(hi[1]), z1 = (int)floorf(hi[2]);
		if (x1 < x0 || y1 < y0 || z1 < z0) continue;
		const uint32_t nx = (uint32_t)(x1 - x0 + 1), ny = (uint32_t)(y1 - y0 + 1), nz = (uint32_t)(z1 - z0 + 1);
		const uint64_t n = (uint64_t)nx * ny * nz;
		for (uint64_t i = lane; i < n; i += 32) {
			const int x = x0 + (int)(i % nx), y = y0 + (int)((i / nx) % ny), z = z0 + (int)(i / ((uint64_t)nx * ny));
			size_t at;
			if (!g.index(x, y, z, at)) continue;
			if (mode == 2) {
				if (!(pointTriangleDistance(V3{ (float)x, (float)y, (float)z }, tri) <= 1.0f)) continue;
				if (g.voxels[at] != background || thin) atomicMax(order + at, t + 1u);
			} else {
				if (!touchesIntersectionTarget(x, y, z, tri)) continue;
				if (mode == 0) g.voxels[at] = (uint8_t)(background + (uint8_t)((x & 1) ^ (y & 1) ^ (z & 1)) + 1u);
				else atomicMax(order + at, t + 1u);
			}
		}
	}
}

__global__ void __launch_bounds__(256)
resolveKernel(uint8_t* __restrict__ voxels, unsigned int* __restrict__ order, const uint8_t* __restrict__ materials, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned int o = order[i];
		if (o) { voxels[i] = materials[o - 1u]; order[i] = 0u; }
	}
}

// Level-1 occupancy from the voxels, level j + 1 from level j: a cell is occupied if it holds a non-background voxel.
__global__ void __launch_bounds__(256)
occupancyKernel(const uint8_t* __restrict__ below, uint32_t belowLog2, uint8_t background, int fromVoxels, uint8_t* __restrict__ out, int* __restrict__ bounds)
{
	const uint32_t log2 = belowLog2 - 1u;
	const size_t n = (size_t)1 << (3 * log2);
	const uint32_t S = 1u << log2, B = 1u << belowLog2;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t x = (uint32_t)(i & (S - 1u)), y = (uint32_t)((i >> log2) & (S - 1u)), z = (uint32_t)(i >> (2 * log2));
		uint8_t any = 0;
		for (uint32_t c = 0; c < 8; c++) {
			const uint32_t cx = 2 * x + (c & 1u), cy = 2 * y + ((c >> 1) & 1u), cz = 2 * z + (c >> 2);
			const uint8_t v = below[((size_t)cz * B + cy) * B + cx];
			const bool occupied = fromVoxels ? (v != background) : (v != 0);
			if (occupied && fromVoxels) {
				// bounds of the non-background voxels (computeBounds(volume, background), utility.cpp:45-95), grid-local
				atomicMin(bounds + 0, (int)cx); atomicMin(bounds + 1, (int)cy); atomicMin(bounds + 2, (int)cz);
				atomicMax(bounds + 3, (int)cx); atomicMax(bounds + 4, (int)cy); atomicMax(bounds + 5, (int)cz);
			}
			any |= occupied ? 1 : 0;
		}
		out[i] = any;
	}
}

struct Leaf { uint32_t x, y, z, log2; };   // grid-local lower corner and size of an octree leaf to classify

// The leaves of level `log2` (cells of 2^log2 voxels): children of an OCCUPIED parent that are themselves empty, or -- at
// level 0 -- any voxel of an occupied 2x2x2 block; only those that overlap the bounds (NodeFinder, voxelization.cpp:564-595).
// `out` == nullptr: count only.
__global__ void __launch_bounds__(256)
collectKernel(const uint8_t* __restrict__ self /* occupancy at this level; nullptr at level 0 */, const uint8_t* __restrict__ parent /* nullptr: see below */,
	uint32_t gridLog2, uint32_t log2, const int* __restrict__ bounds, unsigned long long* __restrict__ counter, Leaf* __restrict__ out)
{
	const uint32_t cellsLog2 = gridLog2 - log2;
	const size_t n = (size_t)1 << (3 * cellsLog2);
	const uint32_t C = 1u << cellsLog2, P = C >> 1;
	const unsigned lane = threadIdx.x & 31u;
	for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += (size_t)gridDim.x * blockDim.x) {
		const size_t i = base + lane;
		bool take = false;
		uint32_t x = 0, y = 0, z = 0;
		if (i < n) {
			x = (uint32_t)(i & (C - 1u)); y = (uint32_t)((i >> cellsLog2) & (C - 1u)); z = (uint32_t)(i >> (2 * cellsLog2));
			// parent == nullptr: the grid's eight half-side cubes when the grid straddles the octree's octant planes (origin a
			// multiple of S/2 but not of S). Each then has a different octree parent whose other seven children lie outside
			// the grid and are empty, so the parent is occupied exactly when the cube is -- an empty cube is never a leaf here.
			const bool parentOccupied = parent ? (parent[((size_t)(z >> 1) * P + (y >> 1)) * P + (x >> 1)] != 0) : (self[i] != 0);
			const bool leaf = parentOccupied && (self == nullptr || self[i] == 0);
			if (leaf) {
				const int s = 1 << log2;
				const int lx = (int)(x << log2), ly = (int)(y << log2), lz = (int)(z << log2);
				take = !(lx + s - 1 < bounds[0] || lx > bounds[3] || ly + s - 1 < bounds[1] || ly > bounds[4] || lz + s - 1 < bounds[2] || lz > bounds[5]);
			}
		}
		const unsigned mask = __ballot_sync(0xffffffffu, take);
		if (mask == 0u) continue;
		unsigned long long at = 0;
		if (lane == (unsigned)(__ffs(mask) - 1)) at = atomicAdd(counter, (unsigned long long)__popc(mask));
		at = __shfl_sync(0xffffffffu, at, __ffs(mask) - 1);
		if (take && out) out[at + __popc(mask & ((1u << lane) - 1u))] = Leaf{ x << log2, y << log2, z << log2, log2 };
	}
}

// isInside(centre of the leaf) with the flat winding-number sum, in mesh order.
__global__ void __launch_bounds__(256)
classifyKernel(const Leaf* __restrict__ leaves, uint64_t count, const Tri* __restrict__ tris, uint32_t triCount, int ox, int oy, int oz, uint8_t* __restrict__ inside)
{
	__shared__ Tri tile[256];
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	V3 q{ 0.0f, 0.0f, 0.0f };
	if (i < count) {
		const Leaf l = leaves[i];
		const int s = 1 << l.log2;
		// centre = vec3f(lower + upper) * 0.5f with integer corners (voxelization.cpp:588)
		const int lx = ox + (int)l.x, ly = oy + (int)l.y, lz = oz + (int)l.z;
		q = V3{ (float)(lx + (lx + s - 1)) * 0.5f, (float)(ly + (ly + s - 1)) * 0.5f, (float)(lz + (lz + s - 1)) * 0.5f };
	}
	float sum = 0.0f;
	for (uint32_t base = 0; base < triCount; base += 256) {
		__syncthreads();
		if (base + threadIdx.x < triCount) tile[threadIdx.x] = tris[base + threadIdx.x];
		__syncthreads();
		const uint32_t m = min(256u, triCount - base);
		if (i < count) for (uint32_t k = 0; k < m; k++) sum += windingTerm(q, tile[k]);
	}
	if (i < count) inside[i] = windingInside(windingNormalise(sum)) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
fillKernel(const Leaf* __restrict__ leaves, const uint8_t* __restrict__ inside, uint64_t count, uint8_t* __restrict__ voxels, uint32_t gridLog2, uint8_t fill, uint8_t background)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += warps) {
		const Leaf l = leaves[w];
		const uint8_t value = inside[w] ? fill : background;      // setNodeChild(index, childId, fill or background), :659-664
		const uint64_t n = (uint64_t)1 << (3 * l.log2);
		const uint32_t s = 1u << l.log2;
		for (uint64_t k = lane; k < n; k += 32) {
			const uint32_t dx = (uint32_t)(k & (s - 1u)), dy = (uint32_t)((k >> l.log2) & (s - 1u)), dz = (uint32_t)(k >> (2 * l.log2));
			voxels[((size_t)(l.z + dz) << (2 * gridLog2)) | ((size_t)(l.y + dy) << gridLog2) | (l.x + dx)] = value;
		}
	}
}

int blocksFor(uint64_t items, int smCount, int perSm = 8) { uint64_t b = (items + 255) / 256; const uint64_t cap = (uint64_t)smCount * perSm; if (b > cap) b = cap; return (int)(b ? b : 1); }

} // namespace

cudaError_t launchShell(const void* tris, uint32_t count, uint8_t* voxels, uint32_t gridLog2, const int32_t origin[3], int mode, uint8_t background, int thin,
	unsigned int* order, int smCount, cudaStream_t stream)
{
	if (count == 0) return cudaSuccess;
	const Grid g{ voxels, gridLog2, origin[0], origin[1], origin[2] };
	shellKernel<<<blocksFor((uint64_t)count * 32, smCount), 256, 0, stream>>>(static_cast<const Tri*>(tris), count, g, mode, background, thin, order);
	return cudaGetLastError();
}

cudaError_t launchResolve(uint8_t* voxels, unsigned int* order, const uint8_t* materials, size_t n, int smCount, cudaStream_t stream)
{
	resolveKernel<<<blocksFor(n, smCount), 256, 0, stream>>>(voxels, order, materials, n);
	return cudaGetLastError();
}

cudaError_t launchOccupancy(const uint8_t* below, uint32_t belowLog2, uint8_t background, int fromVoxels, uint8_t* out, int* bounds, int smCount, cudaStream_t stream)
{
	occupancyKernel<<<blocksFor((uint64_t)1 << (3 * (belowLog2 - 1)), smCount), 256, 0, stream>>>(below, belowLog2, background, fromVoxels, out, bounds);
	return cudaGetLastError();
}

cudaError_t launchCollect(const uint8_t* self, const uint8_t* parent, uint32_t gridLog2, uint32_t log2, const int* bounds, unsigned long long* counter, void* out,
	int smCount, cudaStream_t stream)
{
	collectKernel<<<blocksFor((uint64_t)1 << (3 * (gridLog2 - log2)), smCount), 256, 0, stream>>>(self, parent, gridLog2, log2, bounds, counter, static_cast<Leaf*>(out));
	return cudaGetLastError();
}

cudaError_t launchClassify(const void* leaves, uint64_t count, const void* tris, uint32_t triCount, const int32_t origin[3], uint8_t* inside, cudaStream_t stream)
{
	if (count == 0) return cudaSuccess;
	classifyKernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(static_cast<const Leaf*>(leaves), count, static_cast<const Tri*>(tris), triCount,
		origin[0], origin[1], origin[2], inside);
	return cudaGetLastError();
}

cudaError_t launchFill(const void* leaves, const uint8_t* inside, uint64_t count, uint8_t* voxels, uint32_t gridLog2, uint8_t fill, uint8_t background, int smCount, cudaStream_t stream)
{
	if (count == 0) return cudaSuccess;
	fillKernel<<<blocksFor(count * 32, smCount), 256, 0, stream>>>(static_cast<const Leaf*>(leaves), inside, count, voxels, gridLog2, fill, background);
	return cudaGetLastError();
}

} // namespace cbq

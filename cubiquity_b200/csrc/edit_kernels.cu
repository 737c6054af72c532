// Device-side sphere brush: the reference's runtime edit (Viewer::onMouseButtonDown -> fillBrush with a SphereBrush,
// reference src/library/voxelization.cpp:825-915, voxelization.h:91-127, viewer.cpp:152-172) applied to the node
// array where it lives, in HBM, instead of on the host followed by a PCIe copy of the dirty tail.
//
// The reference recurses depth first and copies a node on write only once it knows a child changed
// (NodeStore::setNodeChild, storage.cpp:152-167). Here the same recursion runs breadth first, one launch per tree
// level, one thread per (node, child slot): a child cube the brush box does not touch is skipped, a cube whose eight
// corners are inside the brush becomes the material, anything else is copied to the tail of the array (atomic
// bump allocator) and queued for the next level. Copies are therefore made BEFORE it is known whether anything
// below changes; a second sweep, bottom-up, puts the original child back wherever the copy ended up identical to
// its source -- the reference's `if (fresh != old)` -- so the tree below the new root is EXACTLY the reference's
// (tests/test_gpu_edit.py compares unfolded-tree signatures and ray hits). The orphaned copies stay in the tail
// as garbage until the next bake. Existing nodes are never written: every older root (undo history) stays valid.
#include "cbq_internal.h"

namespace cbq {

namespace {

struct BrushItem {
	uint32_t node;          // writable copy made by the level above
	uint32_t parent;        // node whose child slot points at `node` ...
	uint32_t slot;          // ... which slot ...
	uint32_t old;           // ... and what was there before
	int32_t lx, ly, lz;     // lower corner of the node's cube
	uint32_t pad;
};

// state[] words (device): the allocator, the queue and the per-level ranges of the queue.
enum { kNodeTail = 0, kItemTail = 1, kOverflow = 2, kRoot = 4, kLevelBegin = 16, kLevelEnd = 49, kStateWords = 96 };

struct Sphere {
	float cx, cy, cz, radiusSquared;
	float lo[3], hi[3];
};

// SphereBrush::contains (voxelization.h:113-121): squared distance truncated to an integer, compared as float.
__device__ __forceinline__ bool contains(const Sphere& b, float x, float y, float z)
{
	const float dx = x - b.cx, dy = y - b.cy, dz = z - b.cz;
	const long long distSq = (long long)(dx * dx + dy * dy + dz * dz);
	return (float)distSq < b.radiusSquared;
}

// One tree level: thread-slots first, first + stride, ... each take one (item, child slot) pair.
__device__ __forceinline__ void brushLevelRange(uint32_t* nodes, uint32_t capacity, const Sphere& b, uint32_t mat, int height, BrushItem* items,
	uint32_t itemCapacity, unsigned int* state, uint32_t first, uint32_t stride)
{
	const uint32_t begin = state[kLevelBegin + height], count = state[kLevelEnd + height] - begin;
	const uint32_t childHeight = (uint32_t)height - 1u;
	const uint32_t side = 1u << childHeight;
	for (uint32_t t = first; t < count * 8u; t += stride) {
		const BrushItem it = items[begin + (t >> 3)];
		const uint32_t slot = t & 7u;
		const uint32_t cx = slot & 1u, cy = (slot >> 1) & 1u, cz = slot >> 2;
		const int32_t x0 = (int32_t)((uint32_t)it.lx + side * cx), y0 = (int32_t)((uint32_t)it.ly + side * cy), z0 = (int32_t)((uint32_t)it.lz + side * cz);
		const int32_t x1 = (int32_t)((uint32_t)x0 + (side - 1u)), y1 = (int32_t)((uint32_t)y0 + (side - 1u)), z1 = (int32_t)((uint32_t)z0 + (side - 1u));
		const float fl[3] = { (float)x0, (float)y0, (float)z0 }, fu[3] = { (float)x1, (float)y1, (float)z1 };
		bool apart = false;
#pragma unroll
		for (int a = 0; a < 3; a++) if (b.hi[a] < fl[a] || b.lo[a] > fu[a]) apart = true;   // overlaps(Box3f, Box3f), geometry.h:347-357
		if (apart) continue;
		bool inside = true;
#pragma unroll
		for (int k = 0; k < 8; k++)
			if (!contains(b, (k & 4) ? fu[0] : fl[0], (k & 2) ? fu[1] : fl[1], (k & 1) ? fu[2] : fl[2])) inside = false;
		uint32_t* cell = nodes + (size_t)it.node * 8 + slot;
		const uint32_t old = *cell;
		if (old == mat) continue;
		if (height >= 2 && !inside) {
			// Reserve the queue entry only once the node exists: every entry below the queue's tail must have been
			// written, or the next level would walk whatever the work space held before.
			const uint32_t fresh = atomicAdd(&state[kNodeTail], 1u);
			if (fresh >= capacity) { atomicOr(&state[kOverflow], 1u); continue; }
			const uint32_t at = atomicAdd(&state[kItemTail], 1u);
			if (at >= itemCapacity) { atomicOr(&state[kOverflow], 2u); continue; }
			uint4* dst = reinterpret_cast<uint4*>(nodes) + (size_t)fresh * 2;
			if (old < kMaterialCount) { dst[0] = make_uint4(old, old, old, old); dst[1] = dst[0]; }
			else { const uint4* src = reinterpret_cast<const uint4*>(nodes) + (size_t)old * 2; dst[0] = src[0]; dst[1] = src[1]; }
			*cell = fresh;
			BrushItem n; n.node = fresh; n.parent = it.node; n.slot = slot; n.old = old; n.lx = x0; n.ly = y0; n.lz = z0; n.pad = 0;
			items[at] = n;
		} else if (inside) {
			*cell = mat;
		}
	}
}

__global__ void brushLevel(uint32_t* nodes, uint32_t capacity, Sphere b, uint32_t mat, int height, BrushItem* items, uint32_t itemCapacity, unsigned int* state)
{
	brushLevelRange(nodes, capacity, b, mat, height, items, itemCapacity, state, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// Between levels: what was queued during the level is the next level's range.
__device__ __forceinline__ void advanceLevel(unsigned int* state, int height, uint32_t itemCapacity)
{
	const uint32_t tail = state[kItemTail] < itemCapacity ? state[kItemTail] : itemCapacity;
	state[kLevelBegin + height - 1] = state[kLevelEnd + height];
	state[kLevelEnd + height - 1] = tail;
}

__global__ void brushAdvance(unsigned int* state, int height, uint32_t itemCapacity) { advanceLevel(state, height, itemCapacity); }

// Bottom-up: a copy that ended up identical to its source was not needed; point the parent back at the original
// (the reference's `if (fresh != old)`, voxelization.cpp fillBrush / storage.cpp:152-167).
__device__ __forceinline__ void brushRevertRange(uint32_t* nodes, int height, const BrushItem* items, const unsigned int* state, uint32_t first, uint32_t stride)
{
	const uint32_t begin = state[kLevelBegin + height], count = state[kLevelEnd + height] - begin;
	for (uint32_t t = first; t < count; t += stride) {
		const BrushItem it = items[begin + t];
		const uint4* mine = reinterpret_cast<const uint4*>(nodes) + (size_t)it.node * 2;
		const uint4 a = mine[0], c = mine[1];
		uint4 sa, sc;
		if (it.old < kMaterialCount) { sa = make_uint4(it.old, it.old, it.old, it.old); sc = sa; }
		else { const uint4* src = reinterpret_cast<const uint4*>(nodes) + (size_t)it.old * 2; sa = src[0]; sc = src[1]; }
		const bool same = a.x == sa.x && a.y == sa.y && a.z == sa.z && a.w == sa.w && c.x == sc.x && c.y == sc.y && c.z == sc.z && c.w == sc.w;
		if (same) nodes[(size_t)it.parent * 8 + it.slot] = it.old;
	}
}

__global__ void brushRevert(uint32_t* nodes, int height, const BrushItem* items, const unsigned int* state)
{
	brushRevertRange(nodes, height, items, state, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// The writable copy of the root (Volume::checkpoint / cloneRoot, storage.cpp:298-303, 358-371) and the first item.
__device__ __forceinline__ void beginStroke(uint32_t* nodes, uint32_t capacity, uint32_t root, BrushItem* items, unsigned int* state)
{
	const uint32_t fresh = atomicAdd(&state[kNodeTail], 1u);
	state[kLevelBegin + 32] = 0;
	if (fresh >= capacity) { atomicOr(&state[kOverflow], 1u); state[kLevelEnd + 32] = 0; return; }
	for (int c = 0; c < 8; c++) nodes[(size_t)fresh * 8 + c] = root < kMaterialCount ? root : nodes[(size_t)root * 8 + c];
	BrushItem it; it.node = fresh; it.parent = 0; it.slot = 0; it.old = root; it.lx = it.ly = it.lz = (int32_t)0x80000000u; it.pad = 0;
	items[0] = it;
	state[kItemTail] = 1;
	state[kLevelEnd + 32] = 1;
	state[kRoot] = fresh;
}

// The top of the tree holds a handful of items per level (the brush box bounds them): heights 32 .. lowest are
// walked by ONE block, a barrier between levels instead of a launch.
__global__ void __launch_bounds__(256) brushTop(uint32_t* nodes, uint32_t capacity, uint32_t root, Sphere b, uint32_t mat, int lowest, BrushItem* items,
	uint32_t itemCapacity, unsigned int* state)
{
	if (threadIdx.x == 0) beginStroke(nodes, capacity, root, items, state);
	__syncthreads();
	for (int height = 32; height >= lowest; height--) {
		brushLevelRange(nodes, capacity, b, mat, height, items, itemCapacity, state, threadIdx.x, blockDim.x);
		__syncthreads();
		if (threadIdx.x == 0) advanceLevel(state, height, itemCapacity);
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) brushRevertTop(uint32_t* nodes, int lowest, const BrushItem* items, const unsigned int* state)
{
	for (int height = lowest; height <= 31; height++) {
		brushRevertRange(nodes, height, items, state, threadIdx.x, blockDim.x);
		__syncthreads();
	}
}

} // namespace

// Enqueue one brush stroke. state (kStateWords words, device, zeroed except [0] = current node count); afterwards
// [0] = new node count, [2] = overflow flags (1 node array full, 2 work list full), [4] = new root.
// items: work list of itemCapacity 32-byte entries.
cudaError_t launchFillSphere(uint32_t* nodes, uint32_t capacity, uint32_t root, float x, float y, float z, float radius, uint32_t material,
	void* itemList, uint32_t itemCapacity, unsigned int* state, int smCount, cudaStream_t stream, uint64_t* launches)
{
	Sphere b;
	b.cx = x; b.cy = y; b.cz = z; b.radiusSquared = radius * radius;
	b.lo[0] = x - radius; b.lo[1] = y - radius; b.lo[2] = z - radius;
	b.hi[0] = x + radius; b.hi[1] = y + radius; b.hi[2] = z + radius;
	BrushItem* items = static_cast<BrushItem*>(itemList);
	auto blocksFor = [&](int height, int perItem) {
		// The brush box bounds how many cubes of a level can be in play; the top of the tree holds a handful.
		const double across = height >= 31 ? 2.0 : 2.0 * radius / (double)(1u << height) + 2.0;
		uint64_t blocks = (uint64_t)(across * across * across * perItem / 256.0) + 1;
		if (blocks > (uint64_t)smCount * 8) blocks = (uint64_t)smCount * 8;
		return (unsigned)blocks;
	};
	// Heights whose cubes are so large that at most ~64 of them can meet the brush box go to the one-block kernels.
	int lowest = 32;
	while (lowest > 1) {
		const double across = 2.0 * radius / (double)(1u << (lowest - 1)) + 2.0;
		if (across * across * across > 64.0) break;
		lowest--;
	}
	brushTop<<<1, 256, 0, stream>>>(nodes, capacity, root, b, material, lowest, items, itemCapacity, state);
	uint64_t count = 1;
	for (int height = lowest - 1; height >= 1; height--) {
		brushLevel<<<blocksFor(height, 8), 256, 0, stream>>>(nodes, capacity, b, material, height, items, itemCapacity, state);
		brushAdvance<<<1, 1, 0, stream>>>(state, height, itemCapacity);
		count += 2;
	}
	for (int height = 1; height < lowest && height <= 31; height++) {
		brushRevert<<<blocksFor(height, 1), 256, 0, stream>>>(nodes, height, items, state);
		count++;
	}
	if (lowest <= 31) { brushRevertTop<<<1, 256, 0, stream>>>(nodes, lowest, items, state); count++; }
	if (launches) *launches += count;
	return cudaGetLastError();
}

size_t fillSphereStateBytes() { return kStateWords * sizeof(unsigned int); }

} // namespace cbq

// Internal host-side declarations shared by api.cu and the kernel translation units.
#pragma once
#include <vector>

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/cubiquity_b200.h"
#include "traverse.cuh"

namespace cbq {

// Layout of the ONE linear device buffer a volume lives in (cbq_upload):
//   [0, 256)        VolumeHeader
//   [256, 512)      SubDag[8]
//   [512, 4608)     float4 colours[256]   (rgb + pad, so a material colour is one 16-byte load)
//   [4608, ...)     nodes, 32 bytes each, 128-byte aligned start => every node is one L2 sector
constexpr size_t kHeaderOffset = 0;
constexpr size_t kSubDagOffset = 256;
constexpr size_t kColourOffset = 512;
constexpr size_t kNodeOffset = 4608;

struct VolumeHeader {
	uint32_t magic;        // 'CBQ1'
	uint32_t version;
	uint64_t nodeCount;    // including the 256 material nodes
	uint64_t nodeCapacity;
	uint32_t rootIndex;
	uint32_t maxSubDagHeight;
	uint64_t generation;   // bumped by every upload / update
	uint8_t pad[256 - 40];
};
static_assert(sizeof(VolumeHeader) == 256, "header is 256 bytes");

// Row l of a banded rectangle -> image row: bands of 64 rows counted from y0, this share owns every
// bandCount-th band starting at bandIndex. bandCount <= 1: identity.
CBQ_HD uint32_t bandedRow(uint32_t y0, uint32_t localRow, uint32_t bandCount, uint32_t bandIndex)
{
	if (bandCount <= 1u) return y0 + localRow;
	return y0 + ((localRow >> 6) * bandCount + bandIndex) * 64u + (localRow & 63u);
}
inline uint32_t bandedRowCount(uint32_t rectH, uint32_t bandCount, uint32_t bandIndex)
{
	if (bandCount <= 1u) return rectH;
	uint32_t rows = 0;
	for (uint32_t b = bandIndex; b * 64u < rectH; b += bandCount) rows += (rectH - b * 64u < 64u) ? (rectH - b * 64u) : 64u;
	return rows;
}

// Which image pixel a rectangle-local pixel (lx, ly) is. Three layouts share the path tracer's kernels:
//   plain rectangle    x = x0 + lx, y = y0 + ly
//   64-row bands       every bandCount-th band of the rectangle's rows (tile sharding over GPUs)
//   tile list          `tiles` != nullptr: the local image is a column of 64x64-pixel tiles, tile t = ly / 64 sits at
//                      image tile (tiles[t] & 0xffff, tiles[t] >> 16) -- the progressive renderer's tile groups
//                      (glsl/pathtracing.frag:789-803). Tiles may overhang the image: such pixels are invalid.
struct PixelMap {
	uint32_t x0, y0, bandCount, bandIndex;
	const uint32_t* tiles;
	uint32_t imgW, imgH;
};
CBQ_HD bool pixelAt(const PixelMap& m, uint32_t lx, uint32_t ly, uint32_t& x, uint32_t& y)
{
	if (m.tiles == nullptr) { x = m.x0 + lx; y = bandedRow(m.y0, ly, m.bandCount, m.bandIndex); return true; }
	const uint32_t tile = m.tiles[ly >> 6];
	x = (tile & 0xffffu) * 64u + lx;
	y = (tile >> 16) * 64u + (ly & 63u);
	return x < m.imgW && y < m.imgH;
}

struct LaunchConfig {
	int blockThreads;      // threads per CTA
	int blocksPerSm;       // resident CTAs per SM the grid is sized for
	int smCount;
	int refillThreshold;   // idle lanes in a warp that trigger a refill from the ray queue (1..32)
	int refillQuantum;     // tickets are dealt in whole groups of this many (1, 2, 4, 8, 16 or 32)
	int stackLevels;       // entries per lane in the shared-memory stack (max sub-DAG height + 1)
	int sampleGroup;       // wavefront path tracer: samples of a pixel traced together (1..16)
	int secondaryRefill;   // wavefront path tracer: refill threshold of the shadow- and bounce-ray casts (1..32)
};

// What a kernel needs to know about the uploaded volume (device pointers into the volume buffer).
struct VolumeView {
	const uint32_t* nodes;                 // 8 words per node, the reference's layout
	const SubDag* subdags;
};

struct TraceArgs {
	VolumeView volume;
	const Ray* rays;             // nullptr => generate from the camera
	Hit* hits;
	uint64_t count;
	float maxFootprint;
	unsigned long long* queue;   // [0] ticket counter, [1] finished CTAs; both zero, re-armed by the kernel itself
	unsigned long long* abandoned; // device counter, incremented per abandoned ray
	// camera source (rays == nullptr): image size and the pixel rectangle to trace (0 = whole image)
	cbq_camera camera;
	uint32_t width, height;
	uint32_t x0, y0, rectW, rectH;   // rectH counts the rows actually traced (owned rows when banded)
	uint32_t bandCount, bandIndex;   // 64-row band interleave (cbq_pt_params::band_count / band_index)
	const uint32_t* tiles;           // camera source: tile-list layout (PixelMap), rectW = 64, rectH = 64 x tiles
	// optional: batch size read on the device (count = *countPtr * countScale), flag-only results
	const unsigned long long* countPtr;
	uint32_t countScale;
	uint8_t* flags;                  // alone: one byte per ray instead of a record (shadow rays); with `pathHits`: both
	uint4* pathHits;                 // path tracer: 16-byte records {position, material | normal bits | hit} (PathHitSink)
	const uint32_t* flagIndex;       // flag-only results go to flags[flagIndex[slot]]
	cbq_hit_compact* compact;        // != nullptr: 8-byte results (cbq_trace_compact) instead of `hits`
	bool remoteResults;              // `compact` lives in another GPU's memory: coalesce the stores per warp (ParkedCompactSink)
	uint32_t untileWidth;            // != 0: rays are in 8x4-tile order of an image this wide; write hits row-major
	// optional cost feedback (refillThreshold == 32 only): deal the 32-ray tickets in this order / record their cost
	const uint32_t* ticketOrder;     // permutation of [0, ceil(count / 32)), or nullptr
	uint32_t* ticketCost;            // [ceil(count / 32)], or nullptr
};

cudaError_t launchTrace(const TraceArgs& a, bool surface, const LaunchConfig& cfg, cudaStream_t stream);
cudaError_t launchOrderTickets(const uint32_t* cost, uint32_t tickets, uint32_t* hist /* ceil(tickets / 1024) x 256 words */, uint32_t* order, cudaStream_t stream);
cudaError_t launchPrimaryRays(const cbq_camera& cam, uint32_t width, uint32_t height, Ray* rays, cudaStream_t stream, int tiled = 0, uint32_t* pixelOf = nullptr);
cudaError_t launchRandomRays(uint64_t seed, const float lower[3], const float upper[3], uint64_t n, Ray* rays, cudaStream_t stream);

struct RenderArgs {
	VolumeView volume;
	const float4* colours;
	const uint32_t* tiles;           // != nullptr: render this list of 64x64-pixel image tiles (device array) ...
	uint32_t tileCount;              // ... instead of the rectangle / bands of `params`
	cbq_camera camera;
	cbq_pt_params params;
	float* accum;
	unsigned long long* abandoned;
};

// Wavefront path tracer (wavefront_kernels.cu): per-bounce kernels around the ray-cast kernel.
struct WavefrontBuffers {
	size_t pathCapacity = 0;       // PATHS (pixels x samples per group) of the largest wave so far
	size_t pixelCapacity = 0;      // pixels of the largest rectangle so far
	uint4* hits = nullptr;         // surface hits of the current depth, 16 bytes each (PathHitSink) [paths]
	Ray* rays[2] = { nullptr, nullptr };        // bounce rays, ping-pong           [paths]
	uint32_t* pixel[2] = { nullptr, nullptr };  // path id of each compacted slot   [paths]
	uint32_t* rng[2] = { nullptr, nullptr };    // RNG state of each path           [paths]
	float4* hist[2] = { nullptr, nullptr };     // {c[k].rgb, D[k]} of each path, plane k at [k * pathCapacity]  [6 * paths]
	Ray* shadowRays = nullptr;     // sun + sky shadow rays of the live paths      [2 * paths]
	uint8_t* shadowFlags = nullptr;//                                              [2 * paths]
	uint8_t* hitFlags = nullptr;   // hit or miss of each surface ray of the current depth [paths]
	Ray* sunRays = nullptr;        // depth 0: one sun ray per lit pixel, compacted [pixels]
	uint32_t* sunPixel = nullptr;  // the pixel each of them belongs to            [pixels]
	uint8_t* sunFlags = nullptr;   // their results, by pixel                      [pixels]
	uint32_t* pixelHash = nullptr; // per-pixel part of the RNG seed               [pixels]
	uint8_t* onImage = nullptr;    // the pixel exists (tile lists may overhang)   [pixels]
	float4* radiance = nullptr;    // finished radiance of each path id            [paths]
	unsigned long long* counters = nullptr;     // [0..5] live paths per depth, [7] depth-0 sun rays, [8..13] run tickets of the shade kernels
};
// Viewer passes (viewer_kernels.cu): progressive accumulation, normalise, edge-stopping blur.
cudaError_t launchProgressiveAdd(float* scratchRgb, float* rgba, uint32_t width, uint32_t height, uint32_t groupCount, uint32_t groupIndex, float samples, cudaStream_t stream);
cudaError_t launchNormalise(const float* rgba, uint32_t width, uint32_t height, float* rgb, cudaStream_t stream);
cudaError_t launchBlur(const float* in, float* out, uint32_t width, uint32_t height, int vertical, cudaStream_t stream);
cudaError_t launchRngPoints(const uint32_t* seeds, uint64_t n, int draws, float* points, uint32_t* states, cudaStream_t stream);
int wavefrontReserve(WavefrontBuffers& b, size_t paths, size_t pixels);       // cudaError_t as int
void wavefrontRelease(WavefrontBuffers& b);
// nextQueue hands out zeroed ticket counters for the trace launches.
typedef int (*QueueFn)(void* user, cudaStream_t stream, unsigned long long** out);
cudaError_t launchRenderWavefront(const RenderArgs& a, WavefrontBuffers& b, const LaunchConfig& cfg, cudaStream_t stream,
	QueueFn nextQueue, void* user, uint64_t* launches);

// Device-side bake (bake_kernels.cu).
size_t bakeScratchBytes(uint64_t nodeCount, uint64_t* tableSlots);
cudaError_t launchBake(const uint32_t* nodes, uint64_t nodeCount, uint32_t root, uint8_t* scratch, uint64_t tableSlots, uint32_t* out,
	unsigned long long* results, int smCount, cudaStream_t stream, uint64_t* launches, uint32_t levels = 33);
// findSubDAGs on a device array; root is read from *rootPtr when rootPtr != nullptr. *status |= 1 on a runaway chain.
// *worst = max(*worst, largest child word of nodes [0, count)): the device-side half of the child-index check.
cudaError_t launchMaxChild(const uint32_t* nodes, uint64_t count, uint32_t* worst, int smCount, cudaStream_t stream);
cudaError_t launchSubdags(const uint32_t* nodes, uint32_t nodeCount, uint32_t root, const unsigned long long* rootPtr, SubDag* out, uint32_t* status, cudaStream_t stream);

// Dense voxel grid -> complete (un-merged) octree below a height-32 root; launchBake then merges it.
uint64_t denseNodeCount(uint32_t sizeLog2);   // upper bound: what to allocate
// Host-built chains from a grid's eight half-side cubes up to the height-32 root (bake_kernels.cu).
uint32_t topTrie(uint32_t sizeLog2, const int32_t origin[3], const uint32_t cubes[8], uint32_t base, std::vector<uint32_t>& words);
// Brick-wise build: re-indexed append of a brick's merged nodes to the collection.
cudaError_t launchAppendNodes(const uint32_t* src, uint64_t count, uint32_t delta, uint32_t* dst, int smCount, cudaStream_t stream);
cudaError_t launchBuildDense(const uint8_t* voxels, uint32_t sizeLog2, const int32_t origin[3], uint32_t* nodes, uint32_t* root, uint64_t* nodeCount,
	int smCount, cudaStream_t stream, uint64_t* launches, bool chain = true);

// Mesh voxeliser (voxelize_kernels.cu; host helpers in voxelize_host.cpp). Leaf records are 16 bytes.
cudaError_t launchShell(const void* tris, uint32_t count, uint8_t* voxels, uint32_t gridLog2, const int32_t origin[3], int mode, uint8_t background, int thin,
	unsigned int* order, int smCount, cudaStream_t stream);
cudaError_t launchResolve(uint8_t* voxels, unsigned int* order, const uint8_t* materials, size_t n, int smCount, cudaStream_t stream);
cudaError_t launchOccupancy(const uint8_t* below, uint32_t belowLog2, uint8_t background, int fromVoxels, uint8_t* out, int* bounds, int smCount, cudaStream_t stream);
cudaError_t launchCollect(const uint8_t* self, const uint8_t* parent, uint32_t gridLog2, uint32_t log2, const int* bounds, unsigned long long* counter, void* out,
	int smCount, cudaStream_t stream);
cudaError_t launchClassify(const void* leaves, uint64_t count, const void* tris, uint32_t triCount, const int32_t origin[3], uint8_t* inside, cudaStream_t stream);
cudaError_t launchFill(const void* leaves, const uint8_t* inside, uint64_t count, uint8_t* voxels, uint32_t gridLog2, uint8_t fill, uint8_t background, int smCount, cudaStream_t stream);

// Device-side sphere brush (edit_kernels.cu). Work-list entries are 32 bytes; state is fillSphereStateBytes() bytes.
cudaError_t launchFillSphere(uint32_t* nodes, uint32_t capacity, uint32_t root, float x, float y, float z, float radius, uint32_t material,
	void* itemList, uint32_t itemCapacity, unsigned int* state, int smCount, cudaStream_t stream, uint64_t* launches);
size_t fillSphereStateBytes();

} // namespace cbq

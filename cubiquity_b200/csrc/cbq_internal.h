// Internal host-side declarations shared by api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
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

struct LaunchConfig {
	int blockThreads;      // threads per CTA
	int blocksPerSm;       // resident CTAs per SM the grid is sized for
	int smCount;
	int refillThreshold;   // idle lanes in a warp that trigger a refill from the ray queue (1..32)
	int kernel;            // 0 persistent queue kernel, 1 one-thread-per-ray
	int stackLevels;       // entries per lane in the shared-memory stack (max sub-DAG height + 1)
};

struct TraceArgs {
	const uint32_t* nodes;
	const SubDag* subdags;       // device pointer into the volume buffer
	const Ray* rays;             // nullptr => generate from the camera
	Hit* hits;
	uint64_t count;
	float maxFootprint;
	unsigned long long* queue;   // one zeroed 64-bit ticket counter for this launch
	unsigned long long* abandoned; // device counter, incremented per abandoned ray
	// camera source (rays == nullptr)
	cbq_camera camera;
	uint32_t width, height;
};

cudaError_t launchTrace(const TraceArgs& a, bool surface, const LaunchConfig& cfg, cudaStream_t stream);
cudaError_t launchPrimaryRays(const cbq_camera& cam, uint32_t width, uint32_t height, Ray* rays, cudaStream_t stream);

struct RenderArgs {
	const uint32_t* nodes;
	const SubDag* subdags;
	const float4* colours;
	cbq_camera camera;
	cbq_pt_params params;
	float* accum;
	unsigned long long* queue;
	unsigned long long* abandoned;
};
cudaError_t launchRender(const RenderArgs& a, const LaunchConfig& cfg, cudaStream_t stream);

} // namespace cbq

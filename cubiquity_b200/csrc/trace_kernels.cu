// sm_100a ray-cast kernels: batches of intersectVolume (reference src/library/raytracing.cpp:397-478).
//
// COMPILE WITH -fmad=false (see traverse.cuh for the arithmetic contract).
//
// tracePersistent -- the production kernel.
//   * Persistent threads: the grid is (SM count x resident CTAs per SM); every warp loops, taking
//     tickets from ONE global atomic counter (the ray queue). A warp claims kChunk consecutive rays
//     per atomic and deals them to its lanes, so the counter sees 1/kChunk of the traffic.
//   * Dynamic refill: the per-ray work is a flat sequence of `steps` (traverse.cuh). After every
//     kStepsPerRound steps the warp votes (__ballot_sync); if at least `refillThreshold` lanes have
//     retired their ray, those lanes are compacted onto the next tickets (rank = popc of the lower
//     idle lanes) and start new rays while their neighbours keep walking. That is the
//     warp-vote/ballot compaction of the north star, applied continuously instead of per bounce.
//   * Per-ray short stack in SHARED memory, one column per thread ([level][thread]: conflict-free),
//     sized to the deepest sub-DAG of the uploaded volume (13 entries for 4096^3) instead of the
//     reference's 33-entry local array (raytracing.cpp:251).
//   * Node words are fetched through the read-only path (ld.global.nc). A sibling step re-reads a
//     different word of the SAME 32-byte node, i.e. the same L1 sector.
//   * Rays come either from a buffer (24-byte records) or straight from the camera in 8x4-pixel
//     tiles per warp (Morton-like locality for primary rays) -- cbq_raycast_frame_device.
#include "cbq_internal.h"
#include "shading.cuh"

namespace cbq {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
#ifndef CBQ_CHUNK
#define CBQ_CHUNK 32
#endif
constexpr int kChunk = CBQ_CHUNK;     // rays claimed per atomic ticket (small: ~14 claims per warp per frame keeps the tail short)
#ifndef CBQ_STEPS_PER_ROUND
#define CBQ_STEPS_PER_ROUND 8
#endif
constexpr int kStepsPerRound = CBQ_STEPS_PER_ROUND;     // traversal steps between two refill votes

// Node words through the read-only path. The node array starts 128-byte aligned (cbq_internal.h), so
// node * 32 + base never carries into the word offset: one IMAD.WIDE + one LEA instead of a 64-bit
// scaled add.
struct GlobalNodes {
	const uint32_t* __restrict__ base;
	__device__ __forceinline__ const uint32_t* address(uint32_t node, uint32_t slot) const
	{
		const uint64_t a = reinterpret_cast<uint64_t>(base) + (uint64_t)node * 32u;
		const uint32_t lo = (uint32_t)a + (slot << 2);
		return reinterpret_cast<const uint32_t*>((a & 0xffffffff00000000ull) | lo);
	}
	__device__ __forceinline__ uint32_t child(uint32_t node, uint32_t slot) const { return __ldg(address(node, slot)); }
	__device__ __forceinline__ void prefetch(uint32_t) const {}
};

// One column of a [levels][blockDim.x] array in shared memory (conflict-free: consecutive threads, consecutive
// words). `written` has one bit per level stored to since the current sub-DAG was entered: a pop to a level that
// was never pushed (only possible when NaNs or collapsed float planes defeat the `tExit < lastExit` guard,
// raytracing.cpp:285) then reads 0, the value the oracle's zero-initialised array holds, instead of what an
// earlier ray left behind. (Zero-filling the column on entry instead was measured: +1.9 % warp instructions and
// 62 registers instead of 56, profiles/r02_analysis.md.)
struct SharedStack {
	uint32_t column;      // shared-window address of this thread's level-0 slot
	uint32_t strideBytes; // blockDim.x * 4
	uint32_t written;
	__device__ __forceinline__ void store(int h, uint32_t n)
	{
		asm volatile("st.shared.u32 [%0], %1;" :: "r"(column + (uint32_t)h * strideBytes), "r"(n));
		written |= 1u << h;
	}
	__device__ __forceinline__ uint32_t load(int h) const
	{
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(column + (uint32_t)h * strideBytes));
		return ((written >> h) & 1u) ? v : 0u;
	}
	__device__ __forceinline__ void clear(int) { written = 0u; }
};

__device__ __forceinline__ void loadRay(const Ray* __restrict__ rays, uint64_t i, Ray& r)
{
	// 24-byte records are 8-byte aligned: three 64-bit read-only loads.
	const float2* p = reinterpret_cast<const float2*>(rays + i);
	const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
	r.o[0] = a.x; r.o[1] = a.y; r.o[2] = b.x; r.d[0] = b.y; r.d[1] = c.x; r.d[2] = c.y;
}

__device__ __forceinline__ void storeHit(Hit* __restrict__ hits, uint64_t i, const Hit& h)
{
	// 40-byte records are 8-byte aligned: five 64-bit stores.
	uint2* p = reinterpret_cast<uint2*>(hits + i);
	p[0] = make_uint2(h.hit, __float_as_uint(h.distance));
	p[1] = make_uint2(h.material, __float_as_uint(h.position[0]));
	p[2] = make_uint2(__float_as_uint(h.position[1]), __float_as_uint(h.position[2]));
	p[3] = make_uint2(__float_as_uint(h.normal[0]), __float_as_uint(h.normal[1]));
	p[4] = make_uint2(__float_as_uint(h.normal[2]), h.status);
}

// Ray sources. `ticket` is the position in the queue; `slot` is where the result goes.
struct BufferSource {
	const Ray* __restrict__ rays;
	__device__ __forceinline__ void fetch(uint64_t ticket, Ray& r, uint64_t& slot) const { loadRay(rays, ticket, r); slot = ticket; }
};

struct CameraSource {
	cbq_camera cam;
	uint32_t width, height;            // the image the camera is defined for
	uint32_t x0, y0, rectW, rectH;     // the pixel rectangle traced; results are indexed rect-locally
	PixelMap map;                      // rectangle-local pixel -> image pixel (bands, tile lists)
	uint32_t tilesX;
	// Tickets enumerate 8x4-pixel tiles row by row, 32 tickets per tile, so the 32 lanes of a warp
	// start on one compact tile. Tiles overhanging the rectangle produce slot = ~0 (skipped).
	__device__ __forceinline__ void fetch(uint64_t ticket, Ray& r, uint64_t& slot) const
	{
		const uint32_t tile = (uint32_t)(ticket >> 5), within = (uint32_t)ticket & 31u;
		const uint32_t tx = tile % tilesX, ty = tile / tilesX;
		const uint32_t x = tx * 8u + (within & 7u), y = ty * 4u + (within >> 3);
		uint32_t px, py;
		if (x < rectW && y < rectH && pixelAt(map, x, y, px, py)) {
			cameraRay(cam, (int)px, (int)py, (int)width, (int)height, r);
			slot = (uint64_t)y * rectW + x;
		} else {
			slot = ~0ull;
		}
	}
};

// Result sinks. FullSink writes the 40-byte record; FlagSink one byte (hit or not), which is all a
// shadow ray needs (gatherLighting only tests .hit, reference pathtracing_demo.cpp:96,111).
struct FullSink {
	static constexpr bool kParked = false;
	Hit* __restrict__ hits;
	static constexpr bool kNeedsPosition = true;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const { storeHit(hits, slot, h); }
};
// Rays arrive in 8x4-tile order (primaryRays, tiled); results go back to row-major pixel order.
struct TiledSink {
	static constexpr bool kParked = false;
	Hit* __restrict__ hits;
	uint32_t width, tilesX;
	static constexpr bool kNeedsPosition = true;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const
	{
		const uint32_t tile = (uint32_t)(slot >> 5), within = (uint32_t)slot & 31u;
		const uint32_t x = (tile % tilesX) * 8u + (within & 7u), y = (tile / tilesX) * 4u + (within >> 3);
		storeHit(hits, (uint64_t)y * width + x, h);
	}
};
// 8-byte results (cbq_hit_compact): no position, so the ray is not fetched again and one 64-bit store replaces five.
__device__ __forceinline__ uint2 compactRecord(const Hit& h)
{
	const uint32_t nx = __float_as_uint(h.normal[0]), ny = __float_as_uint(h.normal[1]), nz = __float_as_uint(h.normal[2]);
	const uint32_t normal = ((nx << 1) != 0u ? 1u : 0u) | (nx >> 31) << 1 | ((ny << 1) != 0u ? 4u : 0u) | (ny >> 31) << 3 |
	                        ((nz << 1) != 0u ? 16u : 0u) | (nz >> 31) << 5;
	const uint32_t code = (h.material & 0xffu) | normal << 8 | (h.hit ? 1u << 14 : 0u) | (h.status ? 1u << 15 : 0u);
	return make_uint2(__float_as_uint(h.distance), code);
}
struct CompactSink {
	cbq_hit_compact* __restrict__ hits;
	static constexpr bool kNeedsPosition = false;
	static constexpr bool kParked = false;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const { reinterpret_cast<uint2*>(hits)[slot] = compactRecord(h); }
};
// The same records for a buffer in ANOTHER GPU's memory (multi-GPU gather over NVLink, cbq_shared_open): a finished ray
// parks its 8 bytes in two registers and the warp stores them when it takes its next rays -- for tile-ordered batches
// all 32 lanes at once, consecutive slots, i.e. two full 128-byte NVLink writes instead of thirty-two 8-byte ones.
// (Stored one by one as rays finish, eight GPUs' results saturate the receiver's link: 0.55 scaling efficiency.)
struct ParkedCompactSink {
	cbq_hit_compact* __restrict__ hits;
	static constexpr bool kNeedsPosition = false;
	static constexpr bool kParked = true;
	__device__ __forceinline__ void write(uint64_t slot, uint2 record) const { reinterpret_cast<uint2*>(hits)[slot] = record; }
};
struct FlagSink {
	static constexpr bool kParked = false;
	uint8_t* __restrict__ flags;
	static constexpr bool kNeedsPosition = false;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const { flags[slot] = (uint8_t)h.hit; }
};
// Path tracer, surface rays: what the shade kernel needs of a hit in ONE 16-byte word -- position and compactRecord's code
// (material, the six normal bits, hit) -- plus a byte per ray that says hit or miss, so that the shade kernel can count its
// survivors from one coalesced byte per path before it touches the records (wavefront_kernels.cu, pass 1). One 128-bit
// store per finished ray instead of five 64-bit ones, one coalesced 128-bit load per path on the other side.
struct PathHitSink {
	static constexpr bool kParked = false;
	uint4* __restrict__ records;
	uint8_t* __restrict__ flags;
	static constexpr bool kNeedsPosition = true;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const
	{
		records[slot] = make_uint4(__float_as_uint(h.position[0]), __float_as_uint(h.position[1]), __float_as_uint(h.position[2]), compactRecord(h).y);
		flags[slot] = (uint8_t)(h.hit != 0u);
	}
};
// Path tracer, depth-0 sun rays (one per lit pixel, compacted): the flag goes to the ray's PIXEL.
struct ScatterFlagSink {
	static constexpr bool kParked = false;
	uint8_t* __restrict__ flags;
	const uint32_t* __restrict__ index;
	static constexpr bool kNeedsPosition = false;
	__device__ __forceinline__ void write(uint64_t slot, const Hit& h) const { flags[index[slot]] = (uint8_t)h.hit; }
};

#ifndef CBQ_TRACE_MIN_BLOCKS
#define CBQ_TRACE_MIN_BLOCKS 4
#endif

template <bool kSurface, bool kLodOff, bool kDeviceCount, typename Source, typename Sink>
__global__ void __launch_bounds__(256, CBQ_TRACE_MIN_BLOCKS)
tracePersistent(const uint32_t* __restrict__ nodeBase, const SubDag* __restrict__ subdagsGlobal,
	Source source, Sink sink, const uint64_t hostCount, const unsigned long long* __restrict__ countPtr, uint32_t countScale,
	float maxFootprint, int refillThreshold, int refillQuantum,
	unsigned long long* __restrict__ queue, unsigned long long* __restrict__ abandoned,
	const uint32_t* __restrict__ ticketOrder, uint32_t* __restrict__ ticketCost)
{
	// The batch size may live on the device (wavefront path tracer: the number of surviving paths is
	// produced by the previous kernel), so no host round trip is needed between bounces. The plain
	// instantiation keeps it a kernel parameter (constant bank, no registers).
	const uint64_t count = kDeviceCount ? (uint64_t)(*countPtr) * countScale : hostCount;
	extern __shared__ uint32_t stackMem[];
	__shared__ SubDag subdags[8];
	for (uint32_t i = threadIdx.x; i < 64u; i += blockDim.x) reinterpret_cast<uint32_t*>(subdags)[i] = reinterpret_cast<const uint32_t*>(subdagsGlobal)[i];
	__syncthreads();

	const GlobalNodes nodes{ nodeBase };
	// The two address terms are made opaque so that they stay in registers: left alone, the compiler re-derives them
	// (tid, shared-window base, blockDim: 9 extra instructions) at every push and pop.
	uint32_t stackColumn = (uint32_t)__cvta_generic_to_shared(stackMem + threadIdx.x), stackStride = blockDim.x * 4u;
	asm volatile("" : "+r"(stackColumn), "+r"(stackStride));
	SharedStack stack{ stackColumn, stackStride, 0u };
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lowerLanes = (1u << lane) - 1u;

	RayState s;
	s.phase = kPhaseIdle;
	uint64_t slot = 0, ticketOfRay = 0;
	// Warp-uniform window of claimed tickets.
	uint64_t chunkNext = 0, chunkEnd = 0;
	bool drained = false;
	// Cost feedback (ticketCost != nullptr, only with refillThreshold == 32, i.e. one 32-ray ticket per warp at a
	// time): rounds of the loop below this warp spent on each ticket. The next launch over the same batch deals
	// the tickets longest first (ticketOrder), which removes most of the end-of-kernel tail.
	uint32_t rounds = 0;      // since the current ticket was claimed

	// Sink::kParked: the result of the lane's finished ray, waiting to be stored (see ParkedCompactSink).
	uint2 parked = make_uint2(0u, 0u);
	bool hasParked = false;

	for (;;) {
		const unsigned idle = __ballot_sync(kFullMask, s.phase == kPhaseIdle);
		if (Sink::kParked && hasParked && (drained || idle == kFullMask || __popc(idle) >= refillThreshold)) {
			if constexpr (Sink::kParked) sink.write(slot, parked);
			hasParked = false;
		}
		if (idle != 0u && !drained && (__popc(idle) >= refillThreshold || idle == kFullMask)) {
			// Only whole groups of `refillQuantum` tickets are dealt (a power of two <= 32): with tile-ordered
			// rays a quantum of 16 keeps every refill an aligned 8x2-pixel half tile, so the warp's ticket
			// window never drifts off the tile grid.
			int want = __popc(idle) & ~(refillQuantum - 1);
			int myRank = __popc(idle & lowerLanes);
			// Deal from the current chunk, claiming a new one when it runs dry.
			while (want > 0) {
				if (chunkNext == chunkEnd) {
					unsigned long long base = 0;
					if (lane == 0) base = atomicAdd(queue, (unsigned long long)kChunk);
					base = __shfl_sync(kFullMask, base, 0);
					if (base >= count) { drained = true; break; }
					if (ticketOrder) base = (unsigned long long)ticketOrder[base / kChunk] * kChunk;   // a permutation of the tickets
					if (ticketCost) {
						// chunkEnd still describes the ticket this warp has just finished (0 = none yet)
						if (lane == 0 && chunkEnd != 0) ticketCost[(chunkEnd - 1) / kChunk] = rounds;
						rounds = 0;
					}
					chunkNext = base;
					chunkEnd = (base + kChunk < count) ? base + kChunk : count;
				}
				const int avail = (int)(chunkEnd - chunkNext);
				const int take = want < avail ? want : avail;
				if (s.phase == kPhaseIdle && myRank >= 0 && myRank < take) {
					Ray r;
					ticketOfRay = chunkNext + (uint64_t)myRank;
					source.fetch(ticketOfRay, r, slot);
					if (slot != ~0ull) beginRay(s, r);
					myRank = -1;    // served
				} else if (myRank >= take) {
					myRank -= take;
				}
				chunkNext += (uint64_t)take;
				want -= take;
			}
		}
		if (__ballot_sync(kFullMask, s.phase != kPhaseIdle) == 0u) {
			if (drained) break;
			continue;
		}
		rounds++;

#pragma unroll 1
		for (int k = 0; k < kStepsPerRound; k++) {
			if (s.phase == kPhaseIdle) continue;
			Hit out;
			StepResult res;
			if (s.phase == kPhaseOctant) res = stepOctant(s, subdags, stack);
			else res = stepEsvo<kLodOff>(s, fetchNext(s, nodes), nodes, stack, maxFootprint, kSurface, out);
			if (res != kStepContinue) {
				if (res == kStepHit) {
					if (!kSurface) { out.material = 0; out.normal[0] = out.normal[1] = out.normal[2] = 0.0f; }
					out.status = 0;
					if (Sink::kNeedsPosition) {
						Ray r; uint64_t again;
						source.fetch(ticketOfRay, r, again);   // the un-reflected ray, for position = o + d * t
						finishHit(out, r);
					}
				} else {
					clearHit(out);
					if (res == kStepAbandoned) { out.status = CBQ_HIT_ABANDONED; atomicAdd(abandoned, 1ull); }
				}
				if constexpr (Sink::kParked) { parked = compactRecord(out); hasParked = true; }
				else sink.write(slot, out);
				s.phase = kPhaseIdle;
			}
		}
	}
	if (ticketCost && lane == 0 && chunkEnd != 0) ticketCost[(chunkEnd - 1) / kChunk] = rounds;
	// The last CTA out re-arms the ticket counter for the next launch on this stream (queue[1] counts finished CTAs).
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(queue + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) { queue[0] = 0ull; queue[1] = 0ull; __threadfence(); }
	}
}

// Tickets by descending cost: counting sort on min(cost, 255) over 1024-ticket slices, one block per slice.
// Pass 1 leaves each slice's histogram in `hist` ([slices][256]); pass 2 turns the histograms into this slice's
// first output position per bucket and scatters. Order within a bucket is arbitrary: it is a scheduling hint.
__global__ void __launch_bounds__(1024)
ticketHistogram(const uint32_t* __restrict__ cost, uint32_t tickets, uint32_t* __restrict__ hist)
{
	__shared__ unsigned int bucket[256];
	if (threadIdx.x < 256) bucket[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t t = blockIdx.x * 1024u + threadIdx.x;
	if (t < tickets) atomicAdd(&bucket[255u - min(cost[t], 255u)], 1u);
	__syncthreads();
	if (threadIdx.x < 256) hist[blockIdx.x * 256u + threadIdx.x] = bucket[threadIdx.x];
}

__global__ void __launch_bounds__(1024)
ticketScatter(const uint32_t* __restrict__ cost, uint32_t tickets, const uint32_t* __restrict__ hist, uint32_t slices, uint32_t* __restrict__ order)
{
	__shared__ unsigned int total[256], cursor[256];
	if (threadIdx.x < 256) {
		unsigned int all = 0, before = 0;
		for (uint32_t sl = 0; sl < slices; sl++) {
			const unsigned int c = hist[sl * 256u + threadIdx.x];
			all += c;
			if (sl < blockIdx.x) before += c;
		}
		total[threadIdx.x] = all;
		cursor[threadIdx.x] = before;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int run = 0;
		for (int b = 0; b < 256; b++) { const unsigned int c = total[b]; cursor[b] += run; run += c; }
	}
	__syncthreads();
	const uint32_t t = blockIdx.x * 1024u + threadIdx.x;
	if (t < tickets) order[atomicAdd(&cursor[255u - min(cost[t], 255u)], 1u)] = t;
}

// tiled != 0 (needs width % 8 == 0 and height % 4 == 0): ray i belongs to pixel (tile i / 32, lane i % 32) of
// the 8x4-pixel tile enumeration, so the 32 consecutive rays a warp claims from a buffer form one compact
// tile ("Morton-ordered primary rays"); `pixelOf` (nullable) receives y * width + x of every ray.
__global__ void __launch_bounds__(256)
primaryRays(cbq_camera cam, uint32_t width, uint32_t height, int tiled, Ray* __restrict__ rays, uint32_t* __restrict__ pixelOf)
{
	const uint64_t total = (uint64_t)width * height;
	const uint32_t tilesX = width / 8u;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
		if (tiled) {
			const uint32_t tile = (uint32_t)(i >> 5), within = (uint32_t)i & 31u;
			x = (tile % tilesX) * 8u + (within & 7u);
			y = (tile / tilesX) * 4u + (within >> 3);
		}
		if (pixelOf) pixelOf[i] = y * width + x;
		Ray r;
		cameraRay(cam, (int)x, (int)y, (int)width, (int)height, r);
		float2* p = reinterpret_cast<float2*>(rays + i);
		p[0] = make_float2(r.o[0], r.o[1]); p[1] = make_float2(r.o[2], r.d[0]); p[2] = make_float2(r.d[1], r.d[2]);
	}
}

// BASELINE config 3: collision-query style rays, generated on the device from a counter-based hash so
// that 100 M of them need no host memory. Ray i, draw k: u = top 24 bits of mix64(seed + (8 i + k + 1) * phi).
// origin = lower + u * (upper - lower); direction = a point drawn in the unit ball by rejection (draws
// 3.., at most 5 tries, then +z), normalised. cubiquity_b200/rays.py:counter_rays is the same in numpy.
__device__ __forceinline__ float counterUniform(uint64_t seed, uint64_t i, uint32_t k)
{
	uint64_t x = seed + (8ull * i + k + 1ull) * 0x9E3779B97F4A7C15ull;
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256)
randomRays(uint64_t seed, float lx, float ly, float lz, float ex, float ey, float ez, uint64_t n, Ray* __restrict__ rays)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		Ray r;
		r.o[0] = lx + counterUniform(seed, i, 0) * ex;
		r.o[1] = ly + counterUniform(seed, i, 1) * ey;
		r.o[2] = lz + counterUniform(seed, i, 2) * ez;
		float dx = 0.0f, dy = 0.0f, dz = 1.0f;
		// One try per counter pair: (3,4,5) then the same hash at i + n, i + 2n ... keeps 8 draws per ray.
		for (uint32_t t = 0; t < 5; t++) {
			const float x = counterUniform(seed, i + t * n, 3) * 2.0f - 1.0f;
			const float y = counterUniform(seed, i + t * n, 4) * 2.0f - 1.0f;
			const float z = counterUniform(seed, i + t * n, 5) * 2.0f - 1.0f;
			const float r2 = (x * x + y * y) + z * z;
			if (r2 < 1.0f && r2 > 1e-12f) {
				const float len = sqrtf(r2);
				dx = x / len; dy = y / len; dz = z / len;
				break;
			}
		}
		r.d[0] = dx; r.d[1] = dy; r.d[2] = dz;
		float2* p = reinterpret_cast<float2*>(rays + i);
		p[0] = make_float2(r.o[0], r.o[1]); p[1] = make_float2(r.o[2], r.d[0]); p[2] = make_float2(r.d[1], r.d[2]);
	}
}

template <typename Kernel, typename Source, typename Sink>
cudaError_t launchKernel(Kernel kernel, const TraceArgs& a, const Source& src, const Sink& sink, uint64_t tickets, const LaunchConfig& cfg, cudaStream_t stream)
{
	const size_t smem = (size_t)cfg.stackLevels * (size_t)cfg.blockThreads * sizeof(uint32_t);
	cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int grid = cfg.smCount * cfg.blocksPerSm;
	const uint64_t needed = (tickets + (uint64_t)cfg.blockThreads - 1) / (uint64_t)cfg.blockThreads;
	if (!a.countPtr && (uint64_t)grid > needed) grid = (int)(needed ? needed : 1);
	kernel<<<grid, cfg.blockThreads, smem, stream>>>(a.volume.nodes, a.volume.subdags, src, sink, tickets,
		a.countPtr, a.countScale, a.maxFootprint, cfg.refillThreshold, cfg.refillQuantum > 0 ? cfg.refillQuantum : 1,
		a.queue, a.abandoned, a.ticketOrder, a.ticketCost);
	return cudaGetLastError();
}

template <bool kSurface, bool kLodOff, typename Source, typename Sink>
cudaError_t launchPersistentImpl(const TraceArgs& a, const Source& src, const Sink& sink, uint64_t tickets, const LaunchConfig& cfg, cudaStream_t stream)
{
	if (a.countPtr) return launchKernel(tracePersistent<kSurface, kLodOff, true, Source, Sink>, a, src, sink, tickets, cfg, stream);
	return launchKernel(tracePersistent<kSurface, kLodOff, false, Source, Sink>, a, src, sink, tickets, cfg, stream);
}

template <bool kSurface, typename Source, typename Sink>
cudaError_t launchPersistentLod(const TraceArgs& a, const Source& src, const Sink& sink, uint64_t tickets, const LaunchConfig& cfg, cudaStream_t stream)
{
	// maxFootprint == -1 exactly (MAX_FOOTPRINT_DISABLED) selects the division-free LOD test.
	if (a.maxFootprint == CBQ_MAX_FOOTPRINT_DISABLED) return launchPersistentImpl<kSurface, true>(a, src, sink, tickets, cfg, stream);
	return launchPersistentImpl<kSurface, false>(a, src, sink, tickets, cfg, stream);
}

template <typename Source>
cudaError_t launchPersistent(const TraceArgs& a, bool surface, const Source& src, uint64_t tickets, const LaunchConfig& cfg, cudaStream_t stream)
{
	if (a.flags && a.pathHits) return launchPersistentLod<true>(a, src, PathHitSink{ a.pathHits, a.flags }, tickets, cfg, stream);   // path tracer, surface rays
	if (a.flags) {
		// Flag-only results are for shadow rays, which never ask for surface properties.
		if (a.flagIndex) return launchPersistentLod<false>(a, src, ScatterFlagSink{ a.flags, a.flagIndex }, tickets, cfg, stream);
		return launchPersistentLod<false>(a, src, FlagSink{ a.flags }, tickets, cfg, stream);
	}
	if (a.compact && a.remoteResults) {
		const ParkedCompactSink sink{ a.compact };
		return surface ? launchPersistentLod<true>(a, src, sink, tickets, cfg, stream) : launchPersistentLod<false>(a, src, sink, tickets, cfg, stream);
	}
	if (a.compact) {
		const CompactSink sink{ a.compact };
		return surface ? launchPersistentLod<true>(a, src, sink, tickets, cfg, stream) : launchPersistentLod<false>(a, src, sink, tickets, cfg, stream);
	}
	if (a.untileWidth) {
		const TiledSink sink{ a.hits, a.untileWidth, a.untileWidth / 8u };
		return surface ? launchPersistentLod<true>(a, src, sink, tickets, cfg, stream) : launchPersistentLod<false>(a, src, sink, tickets, cfg, stream);
	}
	return surface ? launchPersistentLod<true>(a, src, FullSink{ a.hits }, tickets, cfg, stream)
	               : launchPersistentLod<false>(a, src, FullSink{ a.hits }, tickets, cfg, stream);
}

} // namespace

cudaError_t launchTrace(const TraceArgs& a, bool surface, const LaunchConfig& cfg, cudaStream_t stream)
{
	if (a.rays == nullptr) {
		CameraSource src;
		src.cam = a.camera; src.width = a.width; src.height = a.height;
		src.x0 = a.x0; src.y0 = a.y0; src.rectW = a.rectW ? a.rectW : a.width; src.rectH = a.rectH ? a.rectH : a.height;
		src.map = PixelMap{ a.x0, a.y0, a.bandCount, a.bandIndex, a.tiles, a.width, a.height };
		src.tilesX = (src.rectW + 7u) / 8u;
		const uint64_t tickets = (uint64_t)src.tilesX * ((src.rectH + 3u) / 4u) * 32u;
		return launchPersistent(a, surface, src, tickets, cfg, stream);
	}
	BufferSource src{ a.rays };
	return launchPersistent(a, surface, src, a.count, cfg, stream);
}

cudaError_t launchOrderTickets(const uint32_t* cost, uint32_t tickets, uint32_t* hist, uint32_t* order, cudaStream_t stream)
{
	const uint32_t slices = (tickets + 1023u) / 1024u;            // hist holds slices x 256 words
	ticketHistogram<<<slices, 1024, 0, stream>>>(cost, tickets, hist);
	ticketScatter<<<slices, 1024, 0, stream>>>(cost, tickets, hist, slices, order);
	return cudaGetLastError();
}

cudaError_t launchRandomRays(uint64_t seed, const float lower[3], const float upper[3], uint64_t n, Ray* rays, cudaStream_t stream)
{
	uint64_t blocks = (n + 255) / 256;
	if (blocks > 148u * 16u) blocks = 148u * 16u;
	if (blocks == 0) blocks = 1;
	randomRays<<<(int)blocks, 256, 0, stream>>>(seed, lower[0], lower[1], lower[2], upper[0] - lower[0], upper[1] - lower[1], upper[2] - lower[2], n, rays);
	return cudaGetLastError();
}

cudaError_t launchPrimaryRays(const cbq_camera& cam, uint32_t width, uint32_t height, Ray* rays, cudaStream_t stream, int tiled, uint32_t* pixelOf)
{
	const uint64_t total = (uint64_t)width * height;
	uint64_t blocks = (total + 255) / 256;
	if (blocks > 148u * 8u) blocks = 148u * 8u;
	if (blocks == 0) blocks = 1;
	primaryRays<<<(int)blocks, 256, 0, stream>>>(cam, width, height, tiled, rays, pixelOf);
	return cudaGetLastError();
}

} // namespace cbq

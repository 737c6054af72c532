// sm_100a wavefront path tracer: PathtracingDemo::raytrace and its bounce loop
// (reference src/application/commands/view/pathtracing_demo.cpp:81-229) as per-bounce kernels around
// the ray-cast kernel of trace_kernels.cu.
//
// COMPILE WITH -fmad=false.
//
// One "wave" = a GROUP of up to `sample_group` samples of every pixel of the requested rectangle, traced
// together so that the deeper, thinner bounces still fill the machine. What the samples of a pixel have in
// common is traced ONCE per wave: the primary ray, and the depth-0 sun shadow ray (same primary hit, fixed sun
// direction: one ray per lit pixel instead of one per path, and those rays are as coherent as primary rays).
// Each path writes its radiance to its own slot and a final kernel adds the group's samples to the image IN
// SAMPLE ORDER, so the result is bit-identical to tracing the samples one after another. Per depth d:
//
//   trace   surface rays of depth d  (d = 0: straight from the camera, 8x4-pixel tiles; d > 0: the
//           compacted bounce-ray buffer)                         tracePersistent<surface, PathHitSink> -> 16-byte hits + hit/miss bytes
//   shade   one thread per path: first D[d-1] from the shadow flags of the previous depth (gatherLighting's sum);
//           miss -> fold the path's radiance into its slot; hit -> c[d] = surfaceColour, draw the sky and bounce
//           directions from the path's RNG stream, COMPACT the survivors (slots reserved with one atomicAdd per run of
//           512 paths, ranks by warp ballot), and emit for each survivor its shadow rays and its bounce ray
//   trace   the shadow rays, flag-only results                   tracePersistent<!surface, FlagSink>
// and after the last depth one `lightKernel` folds the paths that are still alive.
//
// A path's history -- {c[k].rgb, D[k]} for every depth it has been lit at -- travels WITH the path: plane k of the
// history buffer is indexed by the path's compacted slot and copied forward at each compaction (16 bytes per level,
// coalesced), instead of sitting in per-pixel planes that thin, compacted batches would read and write one 32-byte
// sector per value (profiles/r02_analysis.md, "Path tracer"). D[k] holds the sun term until the flags are in.
//
// The number of survivors never visits the host: shade leaves it in a device counter and the next
// trace / shade kernels read their batch size from there, so a wave is a fixed sequence of launches
// on one stream. Paths are independent and each pixel has exactly one path per sample, so the order
// in which compaction packs them cannot change any result: for the recursive variant images are
// bit-identical to the oracle.
//
// Why no sort of the secondary rays: the ray-cast kernel is issue-bound, not memory-bound, and with
// mid-flight lane refill it traces incoherent rays as fast as coherent ones (4.5 vs 4.4 Grays/s,
// profiles/r01_analysis.md), so a sort would only add its own cost.
#include "cbq_internal.h"
#include "shading.cuh"

namespace cbq {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxDepth = 6;   // bounces <= 5

struct WaveParams {
	cbq_camera cam;
	cbq_pt_params p;
	uint32_t rectW, rectH;
	uint32_t pixels;          // rectW * rectH
	uint32_t paths;           // pixels * samples in this group; path id = sample * pixels + pixel
	uint32_t sampleIndex;     // frame_id + index of the group's first sample
	int lastDepth;            // paths alive after the lighting of this depth end here
	uint32_t shadowsPerPath;  // include_sun + include_sky
	uint32_t sunShared;       // the depth-0 sun ray is traced once per pixel (include_sun)
	PixelMap map;             // rectangle-local pixel -> image pixel
};

// What the shade and light kernels read and write. `cur` is the batch being shaded, `nxt` its survivors.
struct PathSet {
	uint32_t* pixel;          // path id of each slot
	uint32_t* rng;            // RNG state
	float4* hist;             // [kMaxDepth][capacity]: {c[k].r, c[k].g, c[k].b, D[k]}
	Ray* rays;                // the bounce rays that lead to the next depth
};
struct ShadeIo {
	const float4* colours;
	const uint4* hits;                       // {position, material | normal bits << 8 | hit << 14} per surface ray
	PathSet cur, nxt;
	size_t capacity;                         // slots per history plane
	const unsigned long long* pathCount;     // batch size (nullptr at depth 0: w.paths)
	unsigned long long* liveCount;           // survivors
	unsigned long long* tickets;             // runs handed out so far (zero at launch)
	Ray* shadowRays;                         // written for this depth ...
	const uint8_t* shadowFlags;              // ... read for the previous one
	const uint8_t* hitFlags;                 // one byte per surface ray of this depth: hit or miss
	const uint32_t* pixelHash;               // depth 0: per-pixel part of the RNG seed
	const uint8_t* onImage;                  // depth 0: the pixel exists (tile lists may overhang the image)
	unsigned long long* sunCount;            // depth-0 sun rays: one per lit pixel, compacted ...
	Ray* sunRays;
	uint32_t* sunPixel;                      // ... with the pixel each belongs to,
	const uint8_t* sunFlags;                 // which is where its flag goes [pixels]
	float4* radiance;                        // finished radiance of each path id
};

// A 24-byte ray as three 8-byte stores (rays sit at multiples of 24 bytes in 256-byte aligned buffers).
__device__ __forceinline__ void storeRay(Ray* __restrict__ rays, uint64_t index, float ox, float oy, float oz, float dx, float dy, float dz)
{
	float2* p = reinterpret_cast<float2*>(rays + index);
	p[0] = make_float2(ox, oy); p[1] = make_float2(oz, dx); p[2] = make_float2(dy, dz);
}

// gatherLighting's sum for depth kDepth - 1 (pathtracing_demo.cpp:96-99, :111-114) from that depth's shadow flags.
template <int kDepth>
__device__ __forceinline__ float directLight(const WaveParams& w, const ShadeIo& io, uint64_t slot, uint32_t pathId, float sunTerm)
{
	float d = 0.0f;
	if (kDepth == 1 && w.sunShared) {
		if (!io.sunFlags[pathId % w.pixels]) d += sunTerm;
		if (w.p.include_sky) { if (!io.shadowFlags[slot]) d += 1.5f; }
	} else {
		uint32_t k = 0;
		if (w.p.include_sun) { if (!io.shadowFlags[slot * w.shadowsPerPath + k]) d += sunTerm; k++; }
		if (w.p.include_sky) { if (!io.shadowFlags[slot * w.shadowsPerPath + k]) d += 1.5f; }
	}
	return d;
}

// Fold a finished path back to front (pathtracing_demo.cpp:143, :182-185) from its history {c[k].rgb, D[k]}, k < kLevels.
template <int kLevels>
__device__ __forceinline__ float4 foldPath(int variant, const float4 (&h)[kMaxDepth], float r, float g, float b)
{
	if (variant == CBQ_VARIANT_RECURSIVE) {
#pragma unroll
		for (int k = kLevels - 1; k >= 0; k--) {
			r = h[k].x * (h[k].w + r);
			g = h[k].y * (h[k].w + g);
			b = h[k].z * (h[k].w + b);
		}
	} else if (kLevels > 0) {
		float ir = 0.0f, ig = 0.0f, ib = 0.0f;
		if (kLevels > 1) { ir = h[1].x * h[1].w; ig = h[1].y * h[1].w; ib = h[1].z * h[1].w; }
		r = h[0].x * (h[0].w + ir);
		g = h[0].y * (h[0].w + ig);
		b = h[0].z * (h[0].w + ib);
		const float gamma = (float)(1.0 / 2.2);
		r = powf(r, gamma); g = powf(g, gamma); b = powf(b, gamma);
	}
	return make_float4(r, g, b, 0.0f);
}

// mImage[...] += pixel (pathtracing_demo.cpp:224), the group's samples in order.
__global__ void __launch_bounds__(256)
accumulateKernel(WaveParams w, const float4* __restrict__ radiance, float* __restrict__ accum)
{
	const uint32_t samples = w.paths / w.pixels;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < w.pixels; i += gridDim.x * blockDim.x) {
		uint32_t x, y;
		if (!pixelAt(w.map, i % w.rectW, i / w.rectW, x, y)) continue;        // a tile overhanging the image
		float* px = accum + 3ull * ((uint64_t)y * w.p.width + x);
		float r = px[0], g = px[1], b = px[2];
		for (uint32_t s = 0; s < samples; s++) {
			const float4 v = radiance[(size_t)s * w.pixels + i];
			r += v.x; g += v.y; b += v.z;
		}
		px[0] = r; px[1] = g; px[2] = b;
	}
}

// Per pixel, once per wave: hashRay(primary ray) of the per-sample RNG seed (glsl/pathtracing.frag:770-780); the
// sample index is mixed in per path. Pixels of a tile that overhangs the image get no ray and no path.
__global__ void __launch_bounds__(256)
seedKernel(WaveParams w, uint32_t* __restrict__ pixelHash, uint8_t* __restrict__ onImage)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < w.pixels; i += gridDim.x * blockDim.x) {
		uint32_t x, y, hash = 0;
		const bool valid = pixelAt(w.map, i % w.rectW, i / w.rectW, x, y);
		if (valid) {
			Ray r;
			cameraRay(w.cam, (int)x, (int)y, (int)w.p.width, (int)w.p.height, r);
			hash = rayHash(r);
		}
		pixelHash[i] = hash;
		onImage[i] = valid ? 1 : 0;
	}
}

// Is path i of the batch a survivor? One byte per path: the surface-ray trace leaves it next to the record.
template <int kDepth>
__device__ __forceinline__ bool survives(const WaveParams& w, const ShadeIo& io, uint64_t i)
{
	if (kDepth == 0) {
		const uint32_t pix = (uint32_t)(i % w.pixels);
		return io.onImage[pix] != 0 && io.hitFlags[pix] != 0;
	}
	return io.hitFlags[i] != 0;
}

// A lane reading or writing its own 24- or 40-byte structure costs the load/store unit one sector request per lane and
// word, and that -- not DRAM, not instruction issue -- is what bounds this kernel (profiles/r02_analysis.md, "Path tracer").
// Hits therefore arrive as one 16-byte word per path (PathHitSink), and the survivors' rays, whose slots are consecutive,
// leave through a per-warp shared-memory stage: whole 8-byte words, consecutive lanes, consecutive addresses.
constexpr int kStageWords = 32 * 2 * 3 + 32 * 3;   // uint2 words per warp: 32 x 2 shadow rays + 32 bounce rays

__device__ __forceinline__ void stageRay(uint2* __restrict__ stage, uint32_t index, float ox, float oy, float oz, float dx, float dy, float dz)
{
	stage[index * 3u + 0u] = make_uint2(__float_as_uint(ox), __float_as_uint(oy));
	stage[index * 3u + 1u] = make_uint2(__float_as_uint(oz), __float_as_uint(dx));
	stage[index * 3u + 2u] = make_uint2(__float_as_uint(dy), __float_as_uint(dz));
}

template <int kDepth>
__global__ void __launch_bounds__(256)
shadeKernel(WaveParams w, ShadeIo io)
{
	__shared__ float ballLut[kBallLutSize];
	__shared__ uint2 stageAll[8][kStageWords];
	fillBallLut(ballLut);
	uint2* const stage = stageAll[threadIdx.x >> 5];
	const uint64_t count = (kDepth == 0) ? (uint64_t)w.paths : (uint64_t)(*io.pathCount);
	const unsigned lane = threadIdx.x & 31u;
	const unsigned below = (1u << lane) - 1u;
	float sunX, sunY, sunZ;
	sunDirection(sunX, sunY, sunZ);
	const bool sharedSun = (kDepth == 0) && w.sunShared;
	const uint32_t perPath = sharedSun ? (w.shadowsPerPath - 1u) : w.shadowsPerPath;     // shadow rays a survivor adds to io.shadowRays
	const bool bounces = (kDepth != w.lastDepth) && ((w.p.variant == CBQ_VARIANT_RECURSIVE) || kDepth == 0);
	// A warp takes RUNS of kRun consecutive paths from a ticket counter until the batch is used up (sky and terrain
	// cost very different amounts of work per path).
	constexpr uint64_t kRun = 512;
	for (;;) {
		unsigned long long ticket = 0;
		if (lane == 0) ticket = atomicAdd(io.tickets, 1ull);
		ticket = __shfl_sync(kFullMask, ticket, 0);
		const uint64_t begin = ticket * kRun;
		if (begin >= count) break;
		const uint64_t end = (begin + kRun < count) ? begin + kRun : count;

		// ---- pass 1: count the run's survivors (and, at depth 0, its sun rays: the first sample of each lit pixel) and
		// reserve their slots with ONE atomic per run, from one byte per path.
		uint32_t liveTotal = 0, sunTotal = 0;
#pragma unroll 16
		for (uint64_t i = begin + lane; i < end; i += 32u) {
			const bool live = survives<kDepth>(w, io, i);
			liveTotal += live ? 1u : 0u;
			sunTotal += (sharedSun && live && i < (uint64_t)w.pixels) ? 1u : 0u;
		}
		liveTotal = __reduce_add_sync(kFullMask, liveTotal);
		if (sharedSun) sunTotal = __reduce_add_sync(kFullMask, sunTotal);
		unsigned long long slot = 0, sunSlot = 0;
		if (lane == 0) {
			if (liveTotal) slot = atomicAdd(io.liveCount, (unsigned long long)liveTotal);
			if (sunTotal) sunSlot = atomicAdd(io.sunCount, (unsigned long long)sunTotal);
		}
		slot = __shfl_sync(kFullMask, slot, 0);
		if (sharedSun) sunSlot = __shfl_sync(kFullMask, sunSlot, 0);

		// ---- pass 2: light the previous depth, fold the paths that end here, shade and spawn the survivors
		for (uint64_t base = begin; base < end; base += 32u) {
			const uint64_t i = base + lane;
			bool live = false, finished = false;
			uint32_t pixel = 0, rng = 0;
			float4 hst[kMaxDepth];       // only [0, kDepth) is touched
			Hit h;
			h.hit = 0;
			bool onImage = false;
			if (i < end) {
				pixel = (kDepth == 0) ? (uint32_t)i : io.cur.pixel[i];   // the path id
				onImage = true;
				if (kDepth == 0) {
					const uint32_t pix = (uint32_t)(i % w.pixels);
					onImage = io.onImage[pix] != 0;     // a pixel of a tile that overhangs the image: no ray was cast, nothing to fold
					rng = io.pixelHash[pix] ^ fmix32(w.sampleIndex + pixel / w.pixels);
				} else {
					rng = io.cur.rng[i];
#pragma unroll
					for (int k = 0; k < kDepth; k++) hst[k] = io.cur.hist[(size_t)k * io.capacity + i];
					hst[kDepth > 0 ? kDepth - 1 : 0].w = directLight<kDepth>(w, io, i, pixel, hst[kDepth > 0 ? kDepth - 1 : 0].w);   // D[kDepth - 1] is complete now
				}
			}
			if (onImage) {
				// depth 0: every sample of a pixel shares the pixel's primary hit
				const uint4 rec = io.hits[(kDepth == 0) ? (i % w.pixels) : i];
				h.hit = (rec.w >> 14) & 1u; h.material = rec.w & 0xffu;
				h.position[0] = __uint_as_float(rec.x); h.position[1] = __uint_as_float(rec.y); h.position[2] = __uint_as_float(rec.z);
#pragma unroll
				for (int a = 0; a < 3; a++) {     // two bits per component: non-zero, sign (-0.0f included)
					const uint32_t two = (rec.w >> (8 + 2 * a)) & 3u;
					h.normal[a] = __uint_as_float(((two & 1u) ? 0x3f800000u : 0u) | ((two & 2u) ? 0x80000000u : 0u));
				}
				live = h.hit != 0;
				finished = !live;
			}
			const unsigned liveMask = __ballot_sync(kFullMask, live);
			const uint32_t rank = (uint32_t)__popc(liveMask & below), survivors = (uint32_t)__popc(liveMask);
			const uint64_t j = slot + rank;
			unsigned sunMask = 0u;
			if (sharedSun && base < (uint64_t)w.pixels) {
				const uint64_t left = (uint64_t)w.pixels - base;
				sunMask = left >= 32u ? liveMask : (liveMask & ((1u << (unsigned)left) - 1u));
			}
			const uint64_t sunJ = sunSlot + (uint64_t)__popc(sunMask & below);

			if (finished) {
				// a missed bounce adds nothing in traceSingleRay (:168-180); everywhere else a miss is the sky (:124, :152)
				const bool dark = (w.p.variant == CBQ_VARIANT_ONE_BOUNCE) && kDepth == 1;
				io.radiance[pixel] = foldPath<kDepth>((int)w.p.variant, hst, dark ? 0.0f : 0.8f, dark ? 0.0f : 0.8f, dark ? 0.0f : 1.0f);
			}
			if (live) {
				float cr, cg, cb;
				surfaceColour(io.colours, h.material, h.position, w.p.add_noise != 0, cr, cg, cb);   // :45-60
				const float nx = h.normal[0], ny = h.normal[1], nz = h.normal[2];
				const float sunTerm = w.p.include_sun ? 0.1f * maxStd(dot3(sunX, sunY, sunZ, nx, ny, nz), 0.0f) : 0.0f;

				// the path's history moves to its new slot; D[kDepth] waits for its flags
#pragma unroll
				for (int k = 0; k < kDepth; k++) io.nxt.hist[(size_t)k * io.capacity + j] = hst[k];
				io.nxt.hist[(size_t)kDepth * io.capacity + j] = make_float4(cr, cg, cb, sunTerm);

				// gatherLighting (:81-118): both shadow rays leave from position + normal * 0.001
				const float sx = h.position[0] + nx * 0.001f, sy = h.position[1] + ny * 0.001f, sz = h.position[2] + nz * 0.001f;
				uint32_t k = 0;
				if (w.p.include_sun) {
					if (sharedSun) {
						if ((sunMask >> lane) & 1u) {
							storeRay(io.sunRays, sunJ, sx, sy, sz, sunX, sunY, sunZ);
							io.sunPixel[sunJ] = (uint32_t)i;
						}
					} else {
						stageRay(stage, rank * perPath + k++, sx, sy, sz, sunX, sunY, sunZ);
					}
				}
				if (w.p.include_sky) {
					float rx, ry, rz;
					unitBallPoint(ballLut, rng, rx, ry, rz);
					const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
					const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
					stageRay(stage, rank * perPath + k++, sx, sy, sz, vx / len, vy / len, vz / len);
				}
				// The bounce direction is drawn whenever the reference draws it: always in the recursive variant
				// (before its depth check, :138-141 then :122), only at depth 0 in traceSingleRay (:165).
				if ((w.p.variant == CBQ_VARIANT_RECURSIVE) || kDepth == 0) {
					float rx, ry, rz;
					unitBallPoint(ballLut, rng, rx, ry, rz);
					if (bounces) {
						const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
						const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
						stageRay(stage + 32 * 2 * 3, rank, h.position[0] + (nx * 0.01f), h.position[1] + (ny * 0.01f), h.position[2] + (nz * 0.01f), vx / len, vy / len, vz / len);
					}
				}
				io.nxt.pixel[j] = pixel;
				io.nxt.rng[j] = rng;
			}
			__syncwarp();
			// the survivors' rays leave for their consecutive slots
			{
				uint2* __restrict__ dst = reinterpret_cast<uint2*>(io.shadowRays + slot * perPath);
				for (uint32_t t = lane; t < survivors * perPath * 3u; t += 32u) dst[t] = stage[t];
			}
			if (bounces) {
				uint2* __restrict__ dst = reinterpret_cast<uint2*>(io.nxt.rays + slot);
				for (uint32_t t = lane; t < survivors * 3u; t += 32u) dst[t] = stage[32 * 2 * 3 + t];
			}
			__syncwarp();
			slot += survivors;
			sunSlot += (uint64_t)__popc(sunMask);
		}
	}
}

// The paths still alive after the last depth: its D from the flags, then the fold. kLevels = lastDepth + 1; `io.cur`
// is the set the last shade kernel wrote.
template <int kLevels>
__global__ void __launch_bounds__(256)
lightKernel(WaveParams w, ShadeIo io)
{
	const uint64_t count = (uint64_t)(*io.pathCount);
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t pixel = io.cur.pixel[j];
		float4 hst[kMaxDepth];
#pragma unroll
		for (int k = 0; k < kLevels; k++) hst[k] = io.cur.hist[(size_t)k * io.capacity + j];
		hst[kLevels - 1].w = directLight<kLevels>(w, io, j, pixel, hst[kLevels - 1].w);
		io.radiance[pixel] = foldPath<kLevels>((int)w.p.variant, hst, 0.0f, 0.0f, 0.0f);
	}
}

// Self-test hook (cbq_rng_points_device): the path tracer's RNG for caller-chosen seeds.
__global__ void __launch_bounds__(256)
rngPointsKernel(const uint32_t* __restrict__ seeds, uint64_t n, int draws, float* __restrict__ points, uint32_t* __restrict__ states)
{
	__shared__ float ballLut[kBallLutSize];
	fillBallLut(ballLut);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t rng = seeds[i];
		for (int d = 0; d < draws; d++) {
			float x, y, z;
			unitBallPoint(ballLut, rng, x, y, z);
			float* p = points + 3 * (i * (uint64_t)draws + d);
			p[0] = x; p[1] = y; p[2] = z;
		}
		states[i] = rng;
	}
}

template <typename T>
cudaError_t grow(T*& p, size_t count)
{
	if (p) cudaFree(p);
	p = nullptr;
	return cudaMalloc(&p, count * sizeof(T));
}

} // namespace

cudaError_t launchRngPoints(const uint32_t* seeds, uint64_t n, int draws, float* points, uint32_t* states, cudaStream_t stream)
{
	uint64_t blocks = (n + 255) / 256;
	if (blocks > 148u * 8u) blocks = 148u * 8u;
	if (blocks == 0) blocks = 1;
	rngPointsKernel<<<(int)blocks, 256, 0, stream>>>(seeds, n, draws, points, states);
	return cudaGetLastError();
}

int wavefrontReserve(WavefrontBuffers& b, size_t paths, size_t pixels)
{
	if (paths <= b.pathCapacity && pixels <= b.pixelCapacity) return (int)cudaSuccess;
	if (paths < b.pathCapacity) paths = b.pathCapacity;
	if (pixels < b.pixelCapacity) pixels = b.pixelCapacity;
	cudaError_t e;
	// Every buffer is freed before it is re-allocated: on a failure half-way nothing may be left looking usable.
#define CBQ_TRY(x) do { e = (x); if (e != cudaSuccess) { wavefrontRelease(b); return (int)e; } } while (0)
	CBQ_TRY(grow(b.hits, paths));
	for (int i = 0; i < 2; i++) {
		CBQ_TRY(grow(b.rays[i], paths)); CBQ_TRY(grow(b.pixel[i], paths)); CBQ_TRY(grow(b.rng[i], paths));
		CBQ_TRY(grow(b.hist[i], (size_t)kMaxDepth * paths));
	}
	CBQ_TRY(grow(b.shadowRays, 2 * paths));
	CBQ_TRY(grow(b.shadowFlags, 2 * paths));
	CBQ_TRY(grow(b.hitFlags, paths));
	CBQ_TRY(grow(b.sunRays, pixels));
	CBQ_TRY(grow(b.sunFlags, pixels));
	CBQ_TRY(grow(b.sunPixel, pixels));
	CBQ_TRY(grow(b.pixelHash, pixels));
	CBQ_TRY(grow(b.onImage, pixels));
	CBQ_TRY(grow(b.radiance, paths));
	if (!b.counters) CBQ_TRY(cudaMalloc(&b.counters, 16 * sizeof(unsigned long long)));
#undef CBQ_TRY
	b.pathCapacity = paths;
	b.pixelCapacity = pixels;
	return (int)cudaSuccess;
}

void wavefrontRelease(WavefrontBuffers& b)
{
	cudaFree(b.hits);
	for (int i = 0; i < 2; i++) { cudaFree(b.rays[i]); cudaFree(b.pixel[i]); cudaFree(b.rng[i]); cudaFree(b.hist[i]); }
	cudaFree(b.shadowRays); cudaFree(b.shadowFlags); cudaFree(b.hitFlags); cudaFree(b.sunRays); cudaFree(b.sunFlags); cudaFree(b.sunPixel); cudaFree(b.pixelHash); cudaFree(b.onImage);
	cudaFree(b.radiance); cudaFree(b.counters);
	b = WavefrontBuffers();
}

namespace {

// One resident wave exactly: every CTA gets the same share of the batch, so a partial second wave would run at a
// fraction of the machine for as long as the first.
template <int kDepth>
void launchShadeAt(int smCount, cudaStream_t stream, const WaveParams& w, const ShadeIo& io)
{
	static int perSm = 0;
	if (perSm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, shadeKernel<kDepth>, 256, 0) != cudaSuccess || perSm <= 0)) perSm = 4;
	shadeKernel<kDepth><<<smCount * perSm, 256, 0, stream>>>(w, io);
}

void launchShade(int depth, int smCount, cudaStream_t stream, const WaveParams& w, const ShadeIo& io)
{
	switch (depth) {
	case 0: launchShadeAt<0>(smCount, stream, w, io); break;
	case 1: launchShadeAt<1>(smCount, stream, w, io); break;
	case 2: launchShadeAt<2>(smCount, stream, w, io); break;
	case 3: launchShadeAt<3>(smCount, stream, w, io); break;
	case 4: launchShadeAt<4>(smCount, stream, w, io); break;
	default: launchShadeAt<5>(smCount, stream, w, io); break;
	}
}

void launchLight(int levels, int grid, cudaStream_t stream, const WaveParams& w, const ShadeIo& io)
{
	switch (levels) {
	case 1: lightKernel<1><<<grid, 256, 0, stream>>>(w, io); break;
	case 2: lightKernel<2><<<grid, 256, 0, stream>>>(w, io); break;
	case 3: lightKernel<3><<<grid, 256, 0, stream>>>(w, io); break;
	case 4: lightKernel<4><<<grid, 256, 0, stream>>>(w, io); break;
	case 5: lightKernel<5><<<grid, 256, 0, stream>>>(w, io); break;
	default: lightKernel<6><<<grid, 256, 0, stream>>>(w, io); break;
	}
}

} // namespace

cudaError_t launchRenderWavefront(const RenderArgs& a, WavefrontBuffers& b, const LaunchConfig& cfg, cudaStream_t stream,
	QueueFn nextQueue, void* user, uint64_t* launches)
{
	const cbq_pt_params& p = a.params;
	WaveParams w;
	w.cam = a.camera; w.p = p;
	w.rectW = p.x1 - p.x0; w.rectH = bandedRowCount(p.y1 - p.y0, p.band_count, p.band_index);
	w.map = PixelMap{ p.x0, p.y0, p.band_count, p.band_index, nullptr, p.width, p.height };
	if (a.tiles) {
		w.rectW = 64; w.rectH = 64 * a.tileCount;
		w.map.tiles = a.tiles;
	}
	if (w.rectW == 0 || w.rectH == 0) return cudaSuccess;
	w.pixels = w.rectW * w.rectH;
	w.shadowsPerPath = (p.include_sun ? 1u : 0u) + (p.include_sky ? 1u : 0u);
	w.sunShared = p.include_sun ? 1u : 0u;
	// Depth at which surviving paths stop: traceSingleRay lights depth 1 and stops (:173-180);
	// traceSingleRayRecurse stops when depth + 1 > bounces (:122).
	w.lastDepth = (p.variant == CBQ_VARIANT_ONE_BOUNCE) ? 1 : (int)p.bounces;
	if (w.lastDepth >= kMaxDepth) return cudaErrorInvalidValue;

	const int shadeGrid = cfg.smCount * 8;
	LaunchConfig surfaceCfg = cfg, shadowCfg = cfg;
	const int secondaryRefill = cfg.secondaryRefill > 0 ? cfg.secondaryRefill : cfg.refillThreshold;
	shadowCfg.refillThreshold = secondaryRefill;
	unsigned long long* const sunCount = b.counters + 7;
	cudaError_t e;
	const uint32_t kGroup = (uint32_t)(cfg.sampleGroup > 0 ? cfg.sampleGroup : 1);
	for (uint32_t s0 = 0; s0 < p.spp; s0 += kGroup) {
		const uint32_t group = (p.spp - s0 < kGroup) ? (p.spp - s0) : kGroup;
		w.sampleIndex = p.frame_id + s0;
		w.paths = w.pixels * group;
		if ((size_t)w.paths > b.pathCapacity || (size_t)w.pixels > b.pixelCapacity) return cudaErrorInvalidValue;
		e = cudaMemsetAsync(b.counters, 0, 16 * sizeof(unsigned long long), stream);
		if (e != cudaSuccess) return e;
		if (s0 == 0) {
			seedKernel<<<shadeGrid, 256, 0, stream>>>(w, b.pixelHash, b.onImage);     // the same for every wave of the call
			e = cudaGetLastError();
			if (e != cudaSuccess) return e;
			*launches += 1;
		}
		for (int d = 0; d <= w.lastDepth; d++) {
			const int cur = d & 1, nxt = cur ^ 1;   // path sets ping-pong: depth d reads [cur], writes [nxt]
			// ---- surface rays of this depth
			TraceArgs t;
			memset(&t, 0, sizeof(t));
			t.volume = a.volume; t.pathHits = b.hits; t.flags = b.hitFlags; t.maxFootprint = p.max_footprint; t.abandoned = a.abandoned;
			if (nextQueue(user, stream, &t.queue) != 0) return cudaErrorUnknown;
			if (d == 0) {
				// one primary ray per PIXEL: the samples of the group share it
				t.rays = nullptr; t.camera = a.camera; t.width = p.width; t.height = p.height;
				t.x0 = p.x0; t.y0 = p.y0; t.rectW = w.rectW; t.rectH = w.rectH; t.count = w.pixels;
				t.bandCount = p.band_count; t.bandIndex = p.band_index; t.tiles = a.tiles;
				surfaceCfg.refillThreshold = 32;           // coherent tiles
			} else {
				t.rays = b.rays[cur]; t.countPtr = b.counters + (d - 1); t.countScale = 1; t.count = w.paths;
				surfaceCfg.refillThreshold = secondaryRefill;
			}
			e = launchTrace(t, true, surfaceCfg, stream);
			if (e != cudaSuccess) return e;
			// ---- light the previous depth, shade + compact + spawn this one
			ShadeIo io;
			io.colours = a.colours; io.hits = b.hits;
			io.cur = PathSet{ b.pixel[cur], b.rng[cur], b.hist[cur], b.rays[cur] };
			io.nxt = PathSet{ b.pixel[nxt], b.rng[nxt], b.hist[nxt], b.rays[nxt] };
			io.capacity = b.pathCapacity;
			io.pathCount = d ? b.counters + (d - 1) : nullptr;
			io.liveCount = b.counters + d;
			io.tickets = b.counters + 8 + d;
			io.shadowRays = b.shadowRays; io.shadowFlags = b.shadowFlags;
			io.hitFlags = b.hitFlags; io.pixelHash = b.pixelHash; io.onImage = b.onImage;
			io.sunCount = sunCount; io.sunRays = b.sunRays; io.sunPixel = b.sunPixel; io.sunFlags = b.sunFlags;
			io.radiance = b.radiance;
			launchShade(d, cfg.smCount, stream, w, io);
			e = cudaGetLastError();
			if (e != cudaSuccess) return e;
			*launches += 2;
			// ---- shadow rays (flag-only results). Depth 0: the sun rays once per lit pixel, whole tiles of neighbours at a time.
			const bool sharedSun = (d == 0) && w.sunShared;
			const uint32_t perPath = sharedSun ? (w.shadowsPerPath - 1u) : w.shadowsPerPath;
			if (sharedSun) {
				TraceArgs sh;
				memset(&sh, 0, sizeof(sh));
				sh.volume = a.volume; sh.rays = b.sunRays; sh.flags = b.sunFlags; sh.flagIndex = b.sunPixel;
				sh.maxFootprint = p.max_footprint; sh.abandoned = a.abandoned;
				sh.countPtr = sunCount; sh.countScale = 1; sh.count = w.pixels;
				if (nextQueue(user, stream, &sh.queue) != 0) return cudaErrorUnknown;
				LaunchConfig sunCfg = cfg;
				sunCfg.refillThreshold = 32;
				e = launchTrace(sh, false, sunCfg, stream);
				if (e != cudaSuccess) return e;
				*launches += 1;
			}
			if (perPath) {
				TraceArgs sh;
				memset(&sh, 0, sizeof(sh));
				sh.volume = a.volume; sh.rays = b.shadowRays; sh.flags = b.shadowFlags;
				sh.maxFootprint = p.max_footprint; sh.abandoned = a.abandoned;
				sh.countPtr = b.counters + d; sh.countScale = perPath; sh.count = (uint64_t)w.paths * perPath;
				if (nextQueue(user, stream, &sh.queue) != 0) return cudaErrorUnknown;
				e = launchTrace(sh, false, shadowCfg, stream);
				if (e != cudaSuccess) return e;
				*launches += 1;
			}
			if (d == w.lastDepth) {
				io.cur = io.nxt;
				io.pathCount = b.counters + d;
				launchLight(d + 1, shadeGrid, stream, w, io);
				e = cudaGetLastError();
				if (e != cudaSuccess) return e;
				*launches += 1;
			}
		}
		accumulateKernel<<<shadeGrid, 256, 0, stream>>>(w, b.radiance, a.accum);
		e = cudaGetLastError();
		if (e != cudaSuccess) return e;
		*launches += 1;
	}
	return cudaSuccess;
}

} // namespace cbq

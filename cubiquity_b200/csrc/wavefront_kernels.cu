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
//           compacted bounce-ray buffer)                         tracePersistent<surface>  -> 40-byte hits
//   shade   one thread per path: first D[d-1] from the shadow flags of the previous depth (gatherLighting's sum);
//           miss -> fold the path's radiance into its slot; hit -> c[d] = surfaceColour, draw the sky and bounce
//           directions from the path's RNG stream, COMPACT the survivors with a warp ballot + one atomicAdd
//           per warp, and emit for each survivor its shadow rays and its bounce ray
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
	const Hit* hits;
	PathSet cur, nxt;
	size_t capacity;                         // slots per history plane
	const unsigned long long* pathCount;     // batch size (nullptr at depth 0: w.paths)
	unsigned long long* liveCount;           // survivors
	Ray* shadowRays;                         // written for this depth ...
	const uint8_t* shadowFlags;              // ... read for the previous one
	unsigned long long* sunCount;            // depth-0 sun rays: one per lit pixel
	Ray* sunRays;
	uint32_t* sunSlot;                       // [pixels]: which of them is this pixel's
	const uint8_t* sunFlags;
	float4* radiance;                        // finished radiance of each path id
};

// gatherLighting's sum for depth kDepth - 1 (pathtracing_demo.cpp:96-99, :111-114) from that depth's shadow flags.
template <int kDepth>
__device__ __forceinline__ float directLight(const WaveParams& w, const ShadeIo& io, uint64_t slot, uint32_t pathId, float sunTerm)
{
	float d = 0.0f;
	if (kDepth == 1 && w.sunShared) {
		if (!io.sunFlags[io.sunSlot[pathId % w.pixels]]) d += sunTerm;
		if (w.p.include_sky) { if (!io.shadowFlags[slot]) d += 1.5f; }
	} else {
		uint32_t k = 0;
		if (w.p.include_sun) { if (!io.shadowFlags[slot * w.shadowsPerPath + k]) d += sunTerm; k++; }
		if (w.p.include_sky) { if (!io.shadowFlags[slot * w.shadowsPerPath + k]) d += 1.5f; }
	}
	return d;
}

// Fold a finished path back to front (pathtracing_demo.cpp:143, :182-185). `lastD` is D[kLevels - 1], which the
// history plane does not hold yet.
template <int kLevels>
__device__ __forceinline__ float4 foldPath(int variant, const float4* __restrict__ hist, size_t capacity, uint64_t slot, float lastD,
	float r, float g, float b)
{
	if (variant == CBQ_VARIANT_RECURSIVE) {
#pragma unroll
		for (int k = kLevels - 1; k >= 0; k--) {
			const float4 v = hist[(size_t)k * capacity + slot];
			const float d = (k == kLevels - 1) ? lastD : v.w;
			r = v.x * (d + r);
			g = v.y * (d + g);
			b = v.z * (d + b);
		}
	} else if (kLevels > 0) {
		float ir = 0.0f, ig = 0.0f, ib = 0.0f;
		if (kLevels > 1) {
			const float4 v1 = hist[capacity + slot];
			const float d1 = (kLevels == 2) ? lastD : v1.w;
			ir = v1.x * d1; ig = v1.y * d1; ib = v1.z * d1;
		}
		const float4 v0 = hist[slot];
		const float d0 = (kLevels == 1) ? lastD : v0.w;
		r = v0.x * (d0 + ir);
		g = v0.y * (d0 + ig);
		b = v0.z * (d0 + ib);
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

template <int kDepth>
__global__ void __launch_bounds__(256)
shadeKernel(WaveParams w, ShadeIo io)
{
	const uint64_t count = (kDepth == 0) ? (uint64_t)w.paths : (uint64_t)(*io.pathCount);
	const unsigned lane = threadIdx.x & 31u;
	float sunX, sunY, sunZ;
	sunDirection(sunX, sunY, sunZ);
	// Whole warps iterate together so the ballots below are convergent.
	const uint64_t warpStride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += warpStride) {
		const uint64_t i = base + lane;
		bool live = false, finished = false;
		uint32_t pixel = 0, rng = 0;
		float lastD = 0.0f;
		Hit h;
		h.hit = 0;
		if (i < count) {
			pixel = (kDepth == 0) ? (uint32_t)i : io.cur.pixel[i];   // the path id
			// depth 0: every sample of a pixel shares the pixel's primary hit
			const uint2* hp = reinterpret_cast<const uint2*>(io.hits + ((kDepth == 0) ? (i % w.pixels) : i));
			const uint2 a = hp[0], b = hp[1], c = hp[2], d = hp[3], e = hp[4];
			h.hit = a.x; h.distance = __uint_as_float(a.y); h.material = b.x;
			h.position[0] = __uint_as_float(b.y); h.position[1] = __uint_as_float(c.x); h.position[2] = __uint_as_float(c.y);
			h.normal[0] = __uint_as_float(d.x); h.normal[1] = __uint_as_float(d.y); h.normal[2] = __uint_as_float(e.x);
			bool onImage = true;
			if (kDepth == 0) {
				const uint32_t pix = pixel % w.pixels, sample = pixel / w.pixels;
				uint32_t x, y;
				onImage = pixelAt(w.map, pix % w.rectW, pix / w.rectW, x, y);
				if (onImage) {
					Ray r;
					cameraRay(w.cam, (int)x, (int)y, (int)w.p.width, (int)w.p.height, r);
					rng = pixelSeed(r, w.sampleIndex + sample);
				}
			} else {
				rng = io.cur.rng[i];
				lastD = directLight<kDepth>(w, io, i, pixel, io.cur.hist[(size_t)(kDepth - 1) * io.capacity + i].w);
			}
			// a pixel of a tile that overhangs the image: no ray was cast, no hit record exists, nothing to fold
			live = onImage && h.hit != 0;
			finished = onImage && h.hit == 0;
		}
		// ---- compaction of the survivors: warp ballot, one atomic per warp, rank by popc. The atomics are issued here
		// and their results picked up after the arithmetic below: every warp of the grid queues on the same counter.
		const unsigned liveMask = __ballot_sync(kFullMask, live);
		const int leader = (liveMask != 0u) ? (__ffs(liveMask) - 1) : 0;
		unsigned long long slotBase = 0, sunBase = 0;
		if (liveMask != 0u && lane == (unsigned)leader) slotBase = atomicAdd(io.liveCount, (unsigned long long)__popc(liveMask));
		unsigned sunMask = 0u;
		if (kDepth == 0 && w.sunShared) {
			sunMask = __ballot_sync(kFullMask, live && i < (uint64_t)w.pixels);      // the pixel's first sample casts its sun ray
			if (sunMask != 0u && lane == (unsigned)(__ffs(sunMask) - 1)) sunBase = atomicAdd(io.sunCount, (unsigned long long)__popc(sunMask));
		}

		if (finished) {
			// a missed bounce adds nothing in traceSingleRay (:168-180); everywhere else a miss is the sky (:124, :152)
			const bool dark = (w.p.variant == CBQ_VARIANT_ONE_BOUNCE) && kDepth == 1;
			io.radiance[pixel] = foldPath<kDepth>((int)w.p.variant, io.cur.hist, io.capacity, i, lastD, dark ? 0.0f : 0.8f, dark ? 0.0f : 0.8f, dark ? 0.0f : 1.0f);
		}

		float cr = 0.0f, cg = 0.0f, cb = 0.0f, sunTerm = 0.0f;
		float skyX = 0.0f, skyY = 0.0f, skyZ = 0.0f, bncX = 0.0f, bncY = 0.0f, bncZ = 0.0f;
		const float nx = h.normal[0], ny = h.normal[1], nz = h.normal[2];
		if (live) {
			surfaceColour(io.colours, h.material, h.position, w.p.add_noise != 0, cr, cg, cb);   // :45-60
			if (w.p.include_sun) sunTerm = 0.1f * maxStd(dot3(sunX, sunY, sunZ, nx, ny, nz), 0.0f);
			if (w.p.include_sky) {
				float rx, ry, rz;
				unitBallPoint(rng, rx, ry, rz);
				const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
				const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
				skyX = vx / len; skyY = vy / len; skyZ = vz / len;
			}
			// The bounce direction is drawn whenever the reference draws it: always in the recursive variant
			// (before its depth check, :138-141 then :122), only at depth 0 in traceSingleRay (:165).
			if ((w.p.variant == CBQ_VARIANT_RECURSIVE) || kDepth == 0) {
				float rx, ry, rz;
				unitBallPoint(rng, rx, ry, rz);
				const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
				const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
				bncX = vx / len; bncY = vy / len; bncZ = vz / len;
			}
		}

		slotBase = __shfl_sync(kFullMask, slotBase, leader);
		if (kDepth == 0 && w.sunShared) sunBase = __shfl_sync(kFullMask, sunBase, (sunMask != 0u) ? (__ffs(sunMask) - 1) : 0);
		if (!live) continue;
		const uint64_t j = slotBase + (uint64_t)__popc(liveMask & ((1u << lane) - 1u));

		// the path's history moves to its new slot; D[kDepth - 1] is complete now, D[kDepth] waits for its flags
#pragma unroll
		for (int k = 0; k < kDepth; k++) {
			float4 v = io.cur.hist[(size_t)k * io.capacity + i];
			if (k == kDepth - 1) v.w = lastD;
			io.nxt.hist[(size_t)k * io.capacity + j] = v;
		}
		io.nxt.hist[(size_t)kDepth * io.capacity + j] = make_float4(cr, cg, cb, sunTerm);

		// gatherLighting (:81-118): both shadow rays leave from position + normal * 0.001
		const float sx = h.position[0] + nx * 0.001f, sy = h.position[1] + ny * 0.001f, sz = h.position[2] + nz * 0.001f;
		const bool sharedSun = (kDepth == 0) && w.sunShared;
		const uint32_t perPath = sharedSun ? (w.shadowsPerPath - 1u) : w.shadowsPerPath;
		uint32_t k = 0;
		if (w.p.include_sun) {
			if (sharedSun) {
				if (i < (uint64_t)w.pixels) {
					const uint64_t s = sunBase + (uint64_t)__popc(sunMask & ((1u << lane) - 1u));
					Ray& r = io.sunRays[s];
					r.o[0] = sx; r.o[1] = sy; r.o[2] = sz; r.d[0] = sunX; r.d[1] = sunY; r.d[2] = sunZ;
					io.sunSlot[i] = (uint32_t)s;
				}
			} else {
				Ray& r = io.shadowRays[j * perPath + k++];
				r.o[0] = sx; r.o[1] = sy; r.o[2] = sz; r.d[0] = sunX; r.d[1] = sunY; r.d[2] = sunZ;
			}
		}
		if (w.p.include_sky) {
			Ray& r = io.shadowRays[j * perPath + k++];
			r.o[0] = sx; r.o[1] = sy; r.o[2] = sz; r.d[0] = skyX; r.d[1] = skyY; r.d[2] = skyZ;
		}
		if (kDepth != w.lastDepth && ((w.p.variant == CBQ_VARIANT_RECURSIVE) || kDepth == 0)) {
			Ray& r = io.nxt.rays[j];
			r.o[0] = h.position[0] + (nx * 0.01f); r.o[1] = h.position[1] + (ny * 0.01f); r.o[2] = h.position[2] + (nz * 0.01f);
			r.d[0] = bncX; r.d[1] = bncY; r.d[2] = bncZ;
		}
		io.nxt.pixel[j] = pixel;
		io.nxt.rng[j] = rng;
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
		const float d = directLight<kLevels>(w, io, j, pixel, io.cur.hist[(size_t)(kLevels - 1) * io.capacity + j].w);
		io.radiance[pixel] = foldPath<kLevels>((int)w.p.variant, io.cur.hist, io.capacity, j, d, 0.0f, 0.0f, 0.0f);
	}
}

// Self-test hook (cbq_rng_points_device): the path tracer's RNG for caller-chosen seeds.
__global__ void __launch_bounds__(256)
rngPointsKernel(const uint32_t* __restrict__ seeds, uint64_t n, int draws, float* __restrict__ points, uint32_t* __restrict__ states)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t rng = seeds[i];
		for (int d = 0; d < draws; d++) {
			float x, y, z;
			unitBallPoint(rng, x, y, z);
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
	CBQ_TRY(grow(b.sunRays, pixels));
	CBQ_TRY(grow(b.sunFlags, pixels));
	CBQ_TRY(grow(b.sunSlot, pixels));
	CBQ_TRY(grow(b.radiance, paths));
	if (!b.counters) CBQ_TRY(cudaMalloc(&b.counters, 8 * sizeof(unsigned long long)));
#undef CBQ_TRY
	b.pathCapacity = paths;
	b.pixelCapacity = pixels;
	return (int)cudaSuccess;
}

void wavefrontRelease(WavefrontBuffers& b)
{
	cudaFree(b.hits);
	for (int i = 0; i < 2; i++) { cudaFree(b.rays[i]); cudaFree(b.pixel[i]); cudaFree(b.rng[i]); cudaFree(b.hist[i]); }
	cudaFree(b.shadowRays); cudaFree(b.shadowFlags); cudaFree(b.sunRays); cudaFree(b.sunFlags); cudaFree(b.sunSlot);
	cudaFree(b.radiance); cudaFree(b.counters);
	b = WavefrontBuffers();
}

namespace {

void launchShade(int depth, int grid, cudaStream_t stream, const WaveParams& w, const ShadeIo& io)
{
	switch (depth) {
	case 0: shadeKernel<0><<<grid, 256, 0, stream>>>(w, io); break;
	case 1: shadeKernel<1><<<grid, 256, 0, stream>>>(w, io); break;
	case 2: shadeKernel<2><<<grid, 256, 0, stream>>>(w, io); break;
	case 3: shadeKernel<3><<<grid, 256, 0, stream>>>(w, io); break;
	case 4: shadeKernel<4><<<grid, 256, 0, stream>>>(w, io); break;
	default: shadeKernel<5><<<grid, 256, 0, stream>>>(w, io); break;
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
	unsigned long long* const sunCount = b.counters + 7;
	cudaError_t e;
	const uint32_t kGroup = (uint32_t)(cfg.sampleGroup > 0 ? cfg.sampleGroup : 1);
	for (uint32_t s0 = 0; s0 < p.spp; s0 += kGroup) {
		const uint32_t group = (p.spp - s0 < kGroup) ? (p.spp - s0) : kGroup;
		w.sampleIndex = p.frame_id + s0;
		w.paths = w.pixels * group;
		if ((size_t)w.paths > b.pathCapacity || (size_t)w.pixels > b.pixelCapacity) return cudaErrorInvalidValue;
		e = cudaMemsetAsync(b.counters, 0, 8 * sizeof(unsigned long long), stream);
		if (e != cudaSuccess) return e;
		for (int d = 0; d <= w.lastDepth; d++) {
			const int cur = d & 1, nxt = cur ^ 1;   // path sets ping-pong: depth d reads [cur], writes [nxt]
			// ---- surface rays of this depth
			TraceArgs t;
			memset(&t, 0, sizeof(t));
			t.volume = a.volume; t.hits = b.hits; t.maxFootprint = p.max_footprint; t.abandoned = a.abandoned;
			if (nextQueue(user, stream, &t.queue) != 0) return cudaErrorUnknown;
			if (d == 0) {
				// one primary ray per PIXEL: the samples of the group share it
				t.rays = nullptr; t.camera = a.camera; t.width = p.width; t.height = p.height;
				t.x0 = p.x0; t.y0 = p.y0; t.rectW = w.rectW; t.rectH = w.rectH; t.count = w.pixels;
				t.bandCount = p.band_count; t.bandIndex = p.band_index; t.tiles = a.tiles;
				surfaceCfg.refillThreshold = 32;           // coherent tiles
			} else {
				t.rays = b.rays[cur]; t.countPtr = b.counters + (d - 1); t.countScale = 1; t.count = w.paths;
				surfaceCfg.refillThreshold = cfg.refillThreshold;
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
			io.shadowRays = b.shadowRays; io.shadowFlags = b.shadowFlags;
			io.sunCount = sunCount; io.sunRays = b.sunRays; io.sunSlot = b.sunSlot; io.sunFlags = b.sunFlags;
			io.radiance = b.radiance;
			launchShade(d, shadeGrid, stream, w, io);
			e = cudaGetLastError();
			if (e != cudaSuccess) return e;
			*launches += 2;
			// ---- shadow rays (flag-only results). Depth 0: the sun rays once per lit pixel, whole tiles of neighbours at a time.
			const bool sharedSun = (d == 0) && w.sunShared;
			const uint32_t perPath = sharedSun ? (w.shadowsPerPath - 1u) : w.shadowsPerPath;
			if (sharedSun) {
				TraceArgs sh;
				memset(&sh, 0, sizeof(sh));
				sh.volume = a.volume; sh.rays = b.sunRays; sh.flags = b.sunFlags;
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

// sm_100a wavefront path tracer: PathtracingDemo::raytrace and its bounce loop
// (reference src/application/commands/view/pathtracing_demo.cpp:81-229) as per-bounce kernels around
// the ray-cast kernel of trace_kernels.cu.
//
// COMPILE WITH -fmad=false.
//
// One "wave" = a GROUP of up to `sample_group` (default 4) samples of every pixel of the requested rectangle, traced
// together so that the deeper, thinner bounces still fill the machine; the samples of a group share
// one primary-ray trace (same pixel, same ray). Each path writes its radiance to its own slot and a
// final kernel adds the group's samples to the image IN SAMPLE ORDER, so the result is bit-identical
// to tracing the samples one after another. Per depth d:
//
//   trace   surface rays of depth d  (d = 0: straight from the camera, 8x4-pixel tiles; d > 0: the
//           compacted bounce-ray buffer)                         tracePersistent<surface>  -> 40-byte hits
//   shade   one thread per path: miss -> fold the path's radiance and add it to the pixel; hit ->
//           c[d] = surfaceColour, draw the sky and bounce directions from the path's RNG stream,
//           COMPACT the survivors with a warp ballot + one atomicAdd per warp, and emit for each
//           survivor its sun and sky shadow rays and its bounce ray
//   trace   the shadow rays, flag-only results                   tracePersistent<!surface, FlagSink>
//   light   D[d] from the two flags; at the last depth fold and accumulate
//
// The number of survivors never visits the host: shade leaves it in a device counter and the next
// trace / light kernels read their batch size from there, so a wave is a fixed sequence of launches
// on one stream. Paths are independent and each pixel has exactly one path per wave, so the order
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
	int depth;
	int lastDepth;            // paths alive after the lighting of this depth end here
	uint32_t shadowsPerPath;  // include_sun + include_sky
	PixelMap map;             // rectangle-local pixel -> image pixel
};

// Fold a finished path back to front (pathtracing_demo.cpp:143, :182-185) into its radiance slot.
__device__ __forceinline__ void finishPath(const WaveParams& w, const float* __restrict__ colour, const float* __restrict__ direct,
	float* __restrict__ radiance, uint32_t pixel /* path id */, int levels, float r, float g, float b)
{
	const size_t n = w.paths;
	if (w.p.variant == CBQ_VARIANT_RECURSIVE) {
		for (int k = levels - 1; k >= 0; k--) {
			const float d = direct[(size_t)k * n + pixel];
			r = colour[((size_t)k * 3 + 0) * n + pixel] * (d + r);
			g = colour[((size_t)k * 3 + 1) * n + pixel] * (d + g);
			b = colour[((size_t)k * 3 + 2) * n + pixel] * (d + b);
		}
	} else if (levels > 0) {
		float ir = 0.0f, ig = 0.0f, ib = 0.0f;
		if (levels > 1) {
			const float d1 = direct[n + pixel];
			ir = colour[(3 + 0) * n + pixel] * d1; ig = colour[(3 + 1) * n + pixel] * d1; ib = colour[(3 + 2) * n + pixel] * d1;
		}
		const float d0 = direct[pixel];
		r = colour[0 * n + pixel] * (d0 + ir);
		g = colour[1 * n + pixel] * (d0 + ig);
		b = colour[2 * n + pixel] * (d0 + ib);
		const float gamma = (float)(1.0 / 2.2);
		r = powf(r, gamma); g = powf(g, gamma); b = powf(b, gamma);
	}
	radiance[pixel] = r; radiance[n + pixel] = g; radiance[2 * n + pixel] = b;
}

// mImage[...] += pixel (pathtracing_demo.cpp:224), the group's samples in order.
__global__ void __launch_bounds__(256)
accumulateKernel(WaveParams w, const float* __restrict__ radiance, float* __restrict__ accum)
{
	const size_t n = w.paths;
	const uint32_t samples = w.paths / w.pixels;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < w.pixels; i += gridDim.x * blockDim.x) {
		uint32_t x, y;
		if (!pixelAt(w.map, i % w.rectW, i / w.rectW, x, y)) continue;        // a tile overhanging the image
		float* px = accum + 3ull * ((uint64_t)y * w.p.width + x);
		float r = px[0], g = px[1], b = px[2];
		for (uint32_t s = 0; s < samples; s++) {
			const size_t id = (size_t)s * w.pixels + i;
			r += radiance[id]; g += radiance[n + id]; b += radiance[2 * n + id];
		}
		px[0] = r; px[1] = g; px[2] = b;
	}
}

__global__ void __launch_bounds__(256)
shadeKernel(WaveParams w, const float4* __restrict__ colours, const Hit* __restrict__ hits,
	const uint32_t* __restrict__ pathPixel, const uint32_t* __restrict__ pathRng, const unsigned long long* __restrict__ pathCount,
	float* __restrict__ colour, float* __restrict__ direct, float* __restrict__ accum /* radiance slots */,
	unsigned long long* __restrict__ liveCount, uint32_t* __restrict__ livePixel, uint32_t* __restrict__ liveRng,
	float* __restrict__ sunTerm, Ray* __restrict__ shadowRays, Ray* __restrict__ bounceRays)
{
	const uint64_t count = (w.depth == 0) ? (uint64_t)w.paths : (uint64_t)(*pathCount);
	const unsigned lane = threadIdx.x & 31u;
	float sunX, sunY, sunZ;
	sunDirection(sunX, sunY, sunZ);
	// Whole warps iterate together so the ballots below are convergent.
	const uint64_t warpStride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += warpStride) {
		const uint64_t i = base + lane;
		bool live = false;
		uint32_t pixel = 0, rng = 0;
		Hit h;
		if (i < count) {
			pixel = (w.depth == 0) ? (uint32_t)i : pathPixel[i];   // the path id
			// depth 0: every sample of a pixel shares the pixel's primary hit
			const uint2* hp = reinterpret_cast<const uint2*>(hits + ((w.depth == 0) ? (i % w.pixels) : i));
			const uint2 a = hp[0], b = hp[1], c = hp[2], d = hp[3], e = hp[4];
			h.hit = a.x; h.distance = __uint_as_float(a.y); h.material = b.x;
			h.position[0] = __uint_as_float(b.y); h.position[1] = __uint_as_float(c.x); h.position[2] = __uint_as_float(c.y);
			h.normal[0] = __uint_as_float(d.x); h.normal[1] = __uint_as_float(d.y); h.normal[2] = __uint_as_float(e.x);
			bool onImage = true;
			if (w.depth == 0) {
				const uint32_t pix = pixel % w.pixels, sample = pixel / w.pixels;
				uint32_t x, y;
				onImage = pixelAt(w.map, pix % w.rectW, pix / w.rectW, x, y);
				if (onImage) {
					Ray r;
					cameraRay(w.cam, (int)x, (int)y, (int)w.p.width, (int)w.p.height, r);
					rng = pixelSeed(r, w.sampleIndex + sample);
				}
			} else {
				rng = pathRng[i];
			}
			if (!onImage) {
				// a pixel of a tile that overhangs the image: no ray was cast, no hit record exists, nothing to fold
			} else if (h.hit) {
				live = true;
			} else if (w.p.variant == CBQ_VARIANT_ONE_BOUNCE && w.depth == 1) {
				finishPath(w, colour, direct, accum, pixel, 1, 0.0f, 0.0f, 0.0f);      // a missed bounce adds nothing (:168-180)
			} else {
				finishPath(w, colour, direct, accum, pixel, w.depth, 0.8f, 0.8f, 1.0f); // sky (:124,152)
			}
		}
		// ---- compaction of the survivors: warp ballot, one atomic per warp, rank by popc
		const unsigned liveMask = __ballot_sync(kFullMask, live);
		unsigned long long slotBase = 0;
		if (liveMask != 0u && lane == (unsigned)(__ffs(liveMask) - 1)) slotBase = atomicAdd(liveCount, (unsigned long long)__popc(liveMask));
		slotBase = __shfl_sync(kFullMask, slotBase, (liveMask != 0u) ? (__ffs(liveMask) - 1) : 0);
		if (!live) continue;
		const uint64_t j = slotBase + (uint64_t)__popc(liveMask & ((1u << lane) - 1u));

		// surfaceColour (pathtracing_demo.cpp:45-60)
		float cr, cg, cb;
		surfaceColour(colours, h.material, h.position, w.p.add_noise != 0, cr, cg, cb);
		const size_t n = w.paths;
		colour[((size_t)w.depth * 3 + 0) * n + pixel] = cr;
		colour[((size_t)w.depth * 3 + 1) * n + pixel] = cg;
		colour[((size_t)w.depth * 3 + 2) * n + pixel] = cb;

		// gatherLighting (:81-118): both shadow rays leave from position + normal * 0.001
		const float nx = h.normal[0], ny = h.normal[1], nz = h.normal[2];
		const float sx = h.position[0] + nx * 0.001f, sy = h.position[1] + ny * 0.001f, sz = h.position[2] + nz * 0.001f;
		uint32_t k = 0;
		if (w.p.include_sun) {
			sunTerm[j] = 0.1f * maxStd(dot3(sunX, sunY, sunZ, nx, ny, nz), 0.0f);
			Ray& r = shadowRays[j * w.shadowsPerPath + k++];
			r.o[0] = sx; r.o[1] = sy; r.o[2] = sz; r.d[0] = sunX; r.d[1] = sunY; r.d[2] = sunZ;
		}
		if (w.p.include_sky) {
			float rx, ry, rz;
			unitBallPoint(rng, rx, ry, rz);
			const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
			const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
			Ray& r = shadowRays[j * w.shadowsPerPath + k++];
			r.o[0] = sx; r.o[1] = sy; r.o[2] = sz; r.d[0] = vx / len; r.d[1] = vy / len; r.d[2] = vz / len;
		}
		// The bounce direction is drawn whenever the reference draws it: always in the recursive variant
		// (before its depth check, :138-141 then :122), only at depth 0 in traceSingleRay (:165).
		const bool drawBounce = (w.p.variant == CBQ_VARIANT_RECURSIVE) || (w.depth == 0);
		if (drawBounce) {
			float rx, ry, rz;
			unitBallPoint(rng, rx, ry, rz);
			if (w.depth != w.lastDepth) {
				const float vx = nx + rx, vy = ny + ry, vz = nz + rz;
				const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
				Ray& r = bounceRays[j];
				r.o[0] = h.position[0] + (nx * 0.01f); r.o[1] = h.position[1] + (ny * 0.01f); r.o[2] = h.position[2] + (nz * 0.01f);
				r.d[0] = vx / len; r.d[1] = vy / len; r.d[2] = vz / len;
			}
		}
		livePixel[j] = pixel;
		liveRng[j] = rng;
	}
}

__global__ void __launch_bounds__(256)
lightKernel(WaveParams w, const unsigned long long* __restrict__ liveCount, const uint32_t* __restrict__ livePixel,
	const float* __restrict__ sunTerm, const uint8_t* __restrict__ shadowFlags,
	const float* __restrict__ colour, float* __restrict__ direct, float* __restrict__ accum)
{
	const uint64_t count = (uint64_t)(*liveCount);
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t pixel = livePixel[j];
		float d = 0.0f;
		uint32_t k = 0;
		if (w.p.include_sun) { if (!shadowFlags[j * w.shadowsPerPath + k]) d += sunTerm[j]; k++; }   // :96-99
		if (w.p.include_sky) { if (!shadowFlags[j * w.shadowsPerPath + k]) d += 1.5f; }              // :111-114
		direct[(size_t)w.depth * w.paths + pixel] = d;
		if (w.depth == w.lastDepth) finishPath(w, colour, direct, accum, pixel, w.depth + 1, 0.0f, 0.0f, 0.0f);
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

int wavefrontReserve(WavefrontBuffers& b, size_t pixels /* paths */)
{
	if (pixels <= b.pixelCapacity) return (int)cudaSuccess;
	cudaError_t e;
	// Every buffer is freed before it is re-allocated: on a failure half-way nothing may be left looking usable.
#define CBQ_TRY(x) do { e = (x); if (e != cudaSuccess) { wavefrontRelease(b); return (int)e; } } while (0)
	CBQ_TRY(grow(b.hits, pixels));
	for (int i = 0; i < 2; i++) { CBQ_TRY(grow(b.rays[i], pixels)); CBQ_TRY(grow(b.pixel[i], pixels)); CBQ_TRY(grow(b.rng[i], pixels)); }
	CBQ_TRY(grow(b.sunTerm, pixels));
	CBQ_TRY(grow(b.shadowRays, 2 * pixels));
	CBQ_TRY(grow(b.shadowFlags, 2 * pixels));
	CBQ_TRY(grow(b.colour, (size_t)kMaxDepth * 3 * pixels));
	CBQ_TRY(grow(b.direct, (size_t)kMaxDepth * pixels));
	CBQ_TRY(grow(b.radiance, 3 * pixels));
	if (!b.counters) CBQ_TRY(cudaMalloc(&b.counters, 8 * sizeof(unsigned long long)));
#undef CBQ_TRY
	b.pixelCapacity = pixels;
	return (int)cudaSuccess;
}

void wavefrontRelease(WavefrontBuffers& b)
{
	cudaFree(b.hits);
	for (int i = 0; i < 2; i++) { cudaFree(b.rays[i]); cudaFree(b.pixel[i]); cudaFree(b.rng[i]); }
	cudaFree(b.sunTerm); cudaFree(b.shadowRays); cudaFree(b.shadowFlags); cudaFree(b.colour); cudaFree(b.direct); cudaFree(b.radiance); cudaFree(b.counters);
	b = WavefrontBuffers();
}

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
	// Depth at which surviving paths stop: traceSingleRay lights depth 1 and stops (:173-180);
	// traceSingleRayRecurse stops when depth + 1 > bounces (:122).
	w.lastDepth = (p.variant == CBQ_VARIANT_ONE_BOUNCE) ? 1 : (int)p.bounces;

	const int shadeGrid = cfg.smCount * 8;
	LaunchConfig surfaceCfg = cfg, shadowCfg = cfg;
	cudaError_t e;
	const uint32_t kGroup = (uint32_t)(cfg.sampleGroup > 0 ? cfg.sampleGroup : 1);
	for (uint32_t s0 = 0; s0 < p.spp; s0 += kGroup) {
		const uint32_t group = (p.spp - s0 < kGroup) ? (p.spp - s0) : kGroup;
		w.sampleIndex = p.frame_id + s0;
		w.paths = w.pixels * group;
		e = cudaMemsetAsync(b.counters, 0, 8 * sizeof(unsigned long long), stream);
		if (e != cudaSuccess) return e;
		for (int d = 0; d <= w.lastDepth; d++) {
			w.depth = d;
			const int cur = d & 1, nxt = cur ^ 1;   // path buffers ping-pong: depth d reads [cur], writes [nxt]
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
			// ---- shade + compact + spawn
			shadeKernel<<<shadeGrid, 256, 0, stream>>>(w, a.colours, b.hits, b.pixel[cur], b.rng[cur], d ? b.counters + (d - 1) : nullptr,
				b.colour, b.direct, b.radiance, b.counters + d, b.pixel[nxt], b.rng[nxt], b.sunTerm, b.shadowRays, b.rays[nxt]);
			e = cudaGetLastError();
			if (e != cudaSuccess) return e;
			*launches += 2;
			// ---- shadow rays (flag-only results)
			if (w.shadowsPerPath) {
				TraceArgs sh;
				memset(&sh, 0, sizeof(sh));
				sh.volume = a.volume; sh.rays = b.shadowRays; sh.flags = b.shadowFlags;
				sh.maxFootprint = p.max_footprint; sh.abandoned = a.abandoned;
				sh.countPtr = b.counters + d; sh.countScale = w.shadowsPerPath; sh.count = (uint64_t)w.paths * w.shadowsPerPath;
				if (nextQueue(user, stream, &sh.queue) != 0) return cudaErrorUnknown;
				e = launchTrace(sh, false, shadowCfg, stream);
				if (e != cudaSuccess) return e;
				*launches += 1;
			}
			lightKernel<<<shadeGrid, 256, 0, stream>>>(w, b.counters + d, b.pixel[nxt], b.sunTerm, b.shadowFlags, b.colour, b.direct, b.radiance);
			e = cudaGetLastError();
			if (e != cudaSuccess) return e;
			*launches += 1;
		}
		accumulateKernel<<<shadeGrid, 256, 0, stream>>>(w, b.radiance, a.accum);
		e = cudaGetLastError();
		if (e != cudaSuccess) return e;
		*launches += 1;
	}
	return cudaSuccess;
}

} // namespace cbq

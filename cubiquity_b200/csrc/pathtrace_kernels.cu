// sm_100a path tracer: PathtracingDemo::raytrace and its bounce loop
// (reference src/application/commands/view/pathtracing_demo.cpp:33-229).
//
// COMPILE WITH -fmad=false.
//
// One persistent kernel. A lane's unit of work is a PIXEL of the requested rectangle (tickets
// enumerate 8x4-pixel tiles so a warp starts on one compact tile); the lane walks that pixel's
// samples in order, and each sample is a little state machine of rays
//     surface ray -> [sun shadow ray] -> [sky shadow ray] -> bounce surface ray -> ...
// every one of which is traced with the SAME flat step loop as the ray-cast kernel
// (traverse.cuh). Lanes in different stages, depths and octants therefore still share one
// instruction stream, a lane whose pixel is finished is refilled from the ticket queue by warp
// vote, and no ray is ever written to memory between bounces.
//
// Radiance is evaluated exactly like the reference's recursion: per depth k the surface colour
// c[k] and direct light D[k] are remembered and, when the path ends, folded back to front
// (L = c[k] * (D[k] + L)), which is the order traceSingleRayRecurse multiplies in
// (pathtracing_demo.cpp:143). Samples of one pixel are added in sample order. With -fmad=false
// that makes variant 1 (recursive) bit-identical to the oracle; variant 0 applies the
// reference's per-sample powf gamma (pathtracing_demo.cpp:184-185), which is not bit-portable
// between libm and CUDA, so that variant is compared with a tolerance.
//
// RNG: the reference's CPU tracer has ONE process-global stream (pathtracing_demo.cpp:33) and is
// therefore order dependent. Like its GLSL sibling (glsl/pathtracing.frag:770-780,816) we seed per
// pixel and sample: state = hashRay(primary ray) ^ fmix32(frame_id + s), advanced with the CPU
// tracer's (u32)bit_mix (pathtracing_demo.cpp:67; base.cpp:72-77).
#include "cbq_internal.h"
#include "shading.cuh"

namespace cbq {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kChunk = 32;            // pixels (one 8x4 tile) claimed per atomic ticket
constexpr int kStepsPerRound = 8;
constexpr int kMaxDepth = 6;          // bounces <= 5 => depths 0..5

// GlobalNodes and SharedStack: see trace_kernels.cu for the commentary.
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

struct SharedStack {
	uint32_t column;
	uint32_t strideBytes;
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
	__device__ __forceinline__ void reset() { written = 0u; }
};

enum RayKind : int { kKindSurface = 0, kKindSun = 1, kKindSky = 2 };

// What one lane is working on.
struct PathState {
	uint32_t x, y;          // pixel
	uint32_t sample;        // current sample index within the call
	uint32_t rng;
	int depth;              // depth of the surface hit being lit / of the surface ray in flight
	int kind;
	Ray ray;                // the un-reflected ray in flight (for position = o + d * t)
	float px, py, pz;       // surface point being lit
	float nx, ny, nz;       // its normal
	float ar, ag, ab;       // running pixel sum
};

#ifndef CBQ_RENDER_MIN_BLOCKS
#define CBQ_RENDER_MIN_BLOCKS 2
#endif

__global__ void __launch_bounds__(256, CBQ_RENDER_MIN_BLOCKS)
renderPersistent(const uint32_t* __restrict__ nodeBase, const SubDag* __restrict__ subdagsGlobal,
	const float4* __restrict__ colours, cbq_camera cam, cbq_pt_params p, uint32_t rectH, float* __restrict__ accum, int refillThreshold,
	unsigned long long* __restrict__ queue, unsigned long long* __restrict__ abandoned)
{
	extern __shared__ uint32_t stackMem[];
	__shared__ SubDag subdags[8];
	if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(subdags)[threadIdx.x] = reinterpret_cast<const uint32_t*>(subdagsGlobal)[threadIdx.x];
	__syncthreads();

	const GlobalNodes nodes{ nodeBase };
	uint32_t stackColumn = (uint32_t)__cvta_generic_to_shared(stackMem + threadIdx.x), stackStride = blockDim.x * 4u;
	asm volatile("" : "+r"(stackColumn), "+r"(stackStride));     // keep both in registers (see tracePersistent)
	SharedStack stack{ stackColumn, stackStride, 0u };
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lowerLanes = (1u << lane) - 1u;
	uint64_t chunkNext = 0, chunkEnd = 0;   // warp-uniform window of claimed tickets

	const uint32_t rectW = p.x1 - p.x0;
	const uint32_t tilesX = (rectW + 7u) / 8u, tilesY = (rectH + 3u) / 4u;   // rectH = owned rows (band interleave)
	const uint64_t tickets = (uint64_t)tilesX * tilesY * 32u;

	// normalize(vec3(1, -2, 10)) (pathtracing_demo.cpp:89): IEEE sqrt and divides, same bits as the host.
	float sunX, sunY, sunZ;
	sunDirection(sunX, sunY, sunZ);

	RayState s;
	s.phase = kPhaseIdle;
	PathState t;
	float colStack[kMaxDepth][3], dirStack[kMaxDepth][3];   // c[k], D[k]
	bool drained = false;

	// Start sample t.sample of the lane's pixel: primary ray + seed.
	auto startSample = [&]() {
		cameraRay(cam, (int)t.x, (int)t.y, (int)p.width, (int)p.height, t.ray);
		t.rng = pixelSeed(t.ray, p.frame_id + t.sample);
		t.depth = 0;
		t.kind = kKindSurface;
		beginRay(s, t.ray);
	};

	// Sample finished with radiance (r, g, b): add it, move to the next sample or retire the pixel.
	auto endSample = [&](float r, float g, float b) {
		t.ar += r; t.ag += g; t.ab += b;
		t.sample++;
		if (t.sample < p.spp) {
			startSample();
		} else {
			float* px = accum + 3ull * ((uint64_t)t.y * p.width + t.x);
			px[0] = t.ar; px[1] = t.ag; px[2] = t.ab;
			s.phase = kPhaseIdle;
		}
	};

	// The path ended below depth `levels - 1` with incoming radiance (r, g, b): fold back to front.
	auto finishPath = [&](int levels, float r, float g, float b) {
		if (p.variant == CBQ_VARIANT_RECURSIVE) {
			for (int k = levels - 1; k >= 0; k--) {
				r = colStack[k][0] * (dirStack[k][0] + r);
				g = colStack[k][1] * (dirStack[k][1] + g);
				b = colStack[k][2] * (dirStack[k][2] + b);
			}
		} else if (levels > 0) {
			// traceSingleRay (pathtracing_demo.cpp:150-190): one indirect bounce, gamma per sample.
			float ir = 0.0f, ig = 0.0f, ib = 0.0f;
			if (levels > 1) { ir = colStack[1][0] * dirStack[1][0]; ig = colStack[1][1] * dirStack[1][1]; ib = colStack[1][2] * dirStack[1][2]; }
			r = colStack[0][0] * (dirStack[0][0] + ir);
			g = colStack[0][1] * (dirStack[0][1] + ig);
			b = colStack[0][2] * (dirStack[0][2] + ib);
			const float gamma = (float)(1.0 / 2.2);
			r = powf(r, gamma); g = powf(g, gamma); b = powf(b, gamma);
		}
		endSample(r, g, b);
	};

	auto spawn = [&](int kind, float ox, float oy, float oz, float dx, float dy, float dz) {
		t.kind = kind;
		t.ray.o[0] = ox; t.ray.o[1] = oy; t.ray.o[2] = oz;
		t.ray.d[0] = dx; t.ray.d[1] = dy; t.ray.d[2] = dz;
		beginRay(s, t.ray);
	};

	// After the lighting of the hit at t.depth is complete: bounce, or end the path.
	auto afterLighting = [&]() {
		if (p.variant == CBQ_VARIANT_ONE_BOUNCE && t.depth >= 1) { finishPath(2, 0.0f, 0.0f, 0.0f); return; }
		// normalize(normal + randomPointInUnitSphere()), from position + normal * 0.01 (pathtracing_demo.cpp:138-139,165-166)
		float rx, ry, rz;
		unitBallPoint(t.rng, rx, ry, rz);
		const float vx = t.nx + rx, vy = t.ny + ry, vz = t.nz + rz;
		const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
		const int next = t.depth + 1;
		if (p.variant == CBQ_VARIANT_RECURSIVE && (uint32_t)next > p.bounces) { finishPath(next, 0.0f, 0.0f, 0.0f); return; } // :122
		t.depth = next;
		spawn(kKindSurface, t.px + (t.nx * 0.01f), t.py + (t.ny * 0.01f), t.pz + (t.nz * 0.01f), vx / len, vy / len, vz / len);
	};

	auto skyStage = [&]() {
		if (p.include_sky) {
			float rx, ry, rz;
			unitBallPoint(t.rng, rx, ry, rz);
			const float vx = t.nx + rx, vy = t.ny + ry, vz = t.nz + rz;
			const float len = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
			spawn(kKindSky, t.px + t.nx * 0.001f, t.py + t.ny * 0.001f, t.pz + t.nz * 0.001f, vx / len, vy / len, vz / len);
		} else {
			afterLighting();
		}
	};

	for (;;) {
		// ---- refill idle lanes with fresh pixels (warp vote + rank compaction)
		const unsigned idle = __ballot_sync(kFullMask, s.phase == kPhaseIdle);
		if (idle != 0u && !drained && (__popc(idle) >= refillThreshold || idle == kFullMask)) {
			int want = __popc(idle);
			int myRank = __popc(idle & lowerLanes);
			while (want > 0) {
				if (chunkNext == chunkEnd) {
					unsigned long long base = 0;
					if (lane == 0) base = atomicAdd(queue, (unsigned long long)kChunk);
					base = __shfl_sync(kFullMask, base, 0);
					if (base >= tickets) { drained = true; break; }
					chunkNext = base;
					chunkEnd = (base + kChunk < tickets) ? base + kChunk : tickets;
				}
				const int avail = (int)(chunkEnd - chunkNext);
				const int take = want < avail ? want : avail;
				if (s.phase == kPhaseIdle && myRank >= 0 && myRank < take) {
					const uint64_t ticket = chunkNext + (uint64_t)myRank;
					const uint32_t tile = (uint32_t)(ticket >> 5), within = (uint32_t)ticket & 31u;
					const uint32_t tx = tile % tilesX, ty = tile / tilesX;
					const uint32_t x = p.x0 + tx * 8u + (within & 7u), ly = ty * 4u + (within >> 3);
					const uint32_t y = bandedRow(p.y0, ly, p.band_count, p.band_index);
					if (x < p.x1 && ly < rectH) {
						t.x = x; t.y = y; t.sample = 0;
						const float* px = accum + 3ull * ((uint64_t)y * p.width + x);
						t.ar = px[0]; t.ag = px[1]; t.ab = px[2];
						startSample();
					}
					myRank = -1;
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

#pragma unroll 1
		for (int k = 0; k < kStepsPerRound; k++) {
			if (s.phase == kPhaseIdle) continue;
			Hit out;
			StepResult res;
			if (s.phase == kPhaseOctant) { stack.reset(); res = stepOctant2(s, subdags); }
			else res = stepEsvo2<false>(s, fetchNext(s, nodes), nodes, stack, p.max_footprint, t.kind == kKindSurface, out);
			if (res == kStepContinue) continue;

			const bool hit = (res == kStepHit);
			if (res == kStepAbandoned) atomicAdd(abandoned, 1ull);   // treated as a miss
			if (t.kind == kKindSurface) {
				if (!hit) {
					if (p.variant == CBQ_VARIANT_ONE_BOUNCE && t.depth == 1) finishPath(1, 0.0f, 0.0f, 0.0f); // missed bounce adds nothing (:168-180)
					else finishPath(t.depth, 0.8f, 0.8f, 1.0f);                                                // sky (:124,152)
					continue;
				}
				finishHit(out, t.ray);
				float cr, cg, cb;
				surfaceColour(colours, out.material, out.position, p.add_noise != 0, cr, cg, cb);
				colStack[t.depth][0] = cr; colStack[t.depth][1] = cg; colStack[t.depth][2] = cb;
				dirStack[t.depth][0] = 0.0f; dirStack[t.depth][1] = 0.0f; dirStack[t.depth][2] = 0.0f;
				t.px = out.position[0]; t.py = out.position[1]; t.pz = out.position[2];
				t.nx = out.normal[0]; t.ny = out.normal[1]; t.nz = out.normal[2];
				// gatherLighting (:81-118)
				if (p.include_sun) spawn(kKindSun, t.px + t.nx * 0.001f, t.py + t.ny * 0.001f, t.pz + t.nz * 0.001f, sunX, sunY, sunZ);
				else skyStage();
			} else if (t.kind == kKindSun) {
				if (!hit) {
					const float d = dot3(sunX, sunY, sunZ, t.nx, t.ny, t.nz);
					const float kk = maxStd(d, 0.0f);
					dirStack[t.depth][0] += 0.1f * kk; dirStack[t.depth][1] += 0.1f * kk; dirStack[t.depth][2] += 0.1f * kk;
				}
				skyStage();
			} else {
				if (!hit) { dirStack[t.depth][0] += 1.5f; dirStack[t.depth][1] += 1.5f; dirStack[t.depth][2] += 1.5f; }
				afterLighting();
			}
		}
	}
}

} // namespace

cudaError_t launchRender(const RenderArgs& a, const LaunchConfig& cfg, cudaStream_t stream)
{
	const int block = 256;
	const size_t smem = (size_t)cfg.stackLevels * block * sizeof(uint32_t);
	cudaError_t e = cudaFuncSetAttribute(renderPersistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	const uint32_t rectW = a.params.x1 - a.params.x0, rectH = bandedRowCount(a.params.y1 - a.params.y0, a.params.band_count, a.params.band_index);
	if (rectW == 0 || rectH == 0) return cudaSuccess;
	const uint64_t tiles = (uint64_t)((rectW + 7u) / 8u) * ((rectH + 3u) / 4u);
	int grid = cfg.smCount * CBQ_RENDER_MIN_BLOCKS;
	const uint64_t needed = (tiles * 32u + block - 1) / block;
	if ((uint64_t)grid > needed) grid = (int)(needed ? needed : 1);
	renderPersistent<<<grid, block, smem, stream>>>(a.nodes, a.subdags, a.colours, a.camera, a.params, rectH, a.accum, cfg.refillThreshold, a.queue, a.abandoned);
	return cudaGetLastError();
}

} // namespace cbq

// SVDAG ray traversal core for sm_100a -- the per-ray state machine.
//
// Computes exactly what Cubiquity::intersectVolume computes (reference
// src/library/raytracing.cpp:397-478) including intersectRayNodeESVO (:213-371),
// findFirstChild (:178-196) and findNearestMaterial (:134-162), bit for bit, but restructured
// for SIMT execution: the reference's two nested loops (octants, then the ESVO do/while) are
// flattened into ONE loop of uniform "steps" so that the lanes of a warp that sit in different
// octants or at different tree depths still execute the same instruction stream, and so that a
// persistent kernel can retire a finished ray and pull a fresh one into the same lane between
// any two steps (see trace_kernels.cu).
//
// Arithmetic contract (SURVEY 8a "arithmetic rules"), which is why this file must be compiled
// with -fmad=false and default (IEEE) -prec-div/-prec-sqrt/-ftz=false:
//   * (float(int) - o) * inv is a subtract then a multiply, never an FMA;
//   * min/max have std::min/std::max comparison semantics (NaN and +-0 matter), so they are
//     written as compare + select, never fminf/fmaxf;
//   * int -> float is round-to-nearest-even, int arithmetic wraps, >> on negatives is arithmetic.
//
// The functions are __host__ __device__ so tests/host_core_check.cpp can run this very code on
// the CPU against the oracle before it ever reaches a GPU. The product never does that: the only
// caller in the library is the CUDA kernels.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define CBQ_HD __host__ __device__ __forceinline__
#else
#define CBQ_HD inline
#endif

namespace cbq {

constexpr uint32_t kMaterialCount = 256;   // Internals::MaterialCount, storage.h:55
constexpr float kFltMax = 3.402823466e+38f;
constexpr uint32_t kNoMaterial = 0xffffffffu;
constexpr uint32_t kNeedFetch = 0xffffffffu;   // never a real child: indices are < node count <= 2^32 - 1

struct SubDag {          // struct SubDAG, raytracing.h:57-65 (32 bytes)
	int32_t lower[3];
	int32_t height;
	uint32_t pad0;
	uint32_t node;
	uint32_t pad1, pad2;
};

struct Ray { float o[3]; float d[3]; };

struct Hit {             // cbq_hit
	uint32_t hit;
	float distance;
	uint32_t material;
	float position[3];
	float normal[3];
	uint32_t status;
};

// std::min / std::max semantics (raytracing.cpp:168-176).
CBQ_HD float minStd(float a, float b) { return (b < a) ? b : a; }
CBQ_HD float maxStd(float a, float b) { return (a < b) ? b : a; }
CBQ_HD float min3(float x, float y, float z) { return minStd(minStd(x, y), z); }
CBQ_HD float max3(float x, float y, float z) { return maxStd(maxStd(x, y), z); }

CBQ_HD int findMsb(uint32_t v)   // raytracing.cpp:16-24; findMSB(0) == -1
{
#if defined(__CUDA_ARCH__)
	return 31 - __clz((int)v);
#else
	return v ? 31 - __builtin_clz(v) : -1;
#endif
}

// Give up after this many trips round the ESVO loop for one sub-DAG of the given height.
// A legitimate traversal needs fewer than 8 * 2^h; the reference itself never terminates on
// the rays this catches (zero direction component with the origin exactly on a cell boundary:
// 0 * inf = NaN, no sibling flip is ever taken). Must match cbqo_iteration_cap in oracle/.
CBQ_HD uint32_t iterationCap(int subdagHeight)
{
	return (subdagHeight >= 20) ? (1u << 26) : ((64u << subdagHeight) + 4096u);
}

enum Phase : int {
	kPhaseIdle = 0,     // lane holds no ray
	kPhaseOctant = 1,   // about to try the sub-DAG of the current root octant (raytracing.cpp:441-475)
	kPhaseEsvo = 2      // inside the do/while of raytracing.cpp:253-367
};

// Everything a ray in flight carries between steps. ~26 registers.
struct RayState {
	// Reflected ray (raytracing.cpp:422-425) and its reciprocal direction (:224).
	float ox, oy, oz;
	float ix, iy, iz;
	// Octant walk (raytracing.cpp:430-434, 470-473).
	float dx, dy, dz;
	int octant;
	int octantTrips;
	uint32_t signBits;
	// ESVO state (raytracing.cpp:234-251).
	uint32_t node;
	int px, py, pz;        // childPos
	uint32_t idBits;       // childId packed x | y << 1 | z << 2
	int childSize;
	int height;            // nodeHeight
	int startHeight;
	float lastExit;
	uint32_t trips;
	int phase;
	// V2 only: t-values of the current child's lower (L) and upper (U) planes, i.e. exactly the
	// childT0 / childT1 of raytracing.cpp:259,270 for the current childPos, carried between steps.
	float Lx, Ly, Lz;
	float Ux, Uy, Uz;
	// V3 only: the node word fetched for the current child (kNeedFetch = not fetched yet).
	uint32_t child;
};

template <typename Nodes>
CBQ_HD uint32_t nearestMaterial(const Nodes& nodes, uint32_t node, uint32_t signBits)
{
	// raytracing.cpp:134-162
	const uint32_t order = (0x76534210u | 0x88888888u) ^ (signBits * 0x11111111u);
	int levels = 0;
	while (node >= kMaterialCount) {
		bool found = false;
		for (uint32_t ids = order; ids != 0; ids >>= 4) {
			const uint32_t child = nodes.child(node, ids & 7u);
			if (child > 0) { node = child; found = true; break; }
		}
		// An internal node with eight empty children (legal in an edited, un-baked volume: fillBrush
		// never collapses nodes) makes the reference spin here for ever; we report "no material" and
		// the caller abandons the ray. Must match nearest_material in oracle/cbq_oracle.c.
		if (!found || ++levels > 32) return kNoMaterial;
	}
	return node;
}

// raytracing.cpp:178-196 with nodeCentre = corner + half.
CBQ_HD uint32_t firstChild(float tEntry, const RayState& s, int cx, int cy, int cz)
{
	const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
	const float tx = (fx - s.ox) * s.ix, ty = (fy - s.oy) * s.iy, tz = (fz - s.oz) * s.iz;
	uint32_t id = (tx < tEntry ? 1u : 0u) | (ty < tEntry ? 2u : 0u) | (tz < tEntry ? 4u : 0u);
	if (tEntry <= 0.0f) {
		id |= (s.ox >= fx ? 1u : 0u) | (s.oy >= fy ? 2u : 0u) | (s.oz >= fz ? 4u : 0u);
	}
	return id;
}

// Set up a ray: raytracing.cpp:407-434.
CBQ_HD void beginRay(RayState& s, const Ray& r)
{
	const uint32_t nx = r.d[0] < 0.0f ? 1u : 0u, ny = r.d[1] < 0.0f ? 1u : 0u, nz = r.d[2] < 0.0f ? 1u : 0u;
	s.signBits = nx | (ny << 1) | (nz << 2);
	const float sx = nx ? -1.0f : 1.0f, sy = ny ? -1.0f : 1.0f, sz = nz ? -1.0f : 1.0f;
	s.ox = (r.o[0] + 0.5f) * sx; s.oy = (r.o[1] + 0.5f) * sy; s.oz = (r.o[2] + 0.5f) * sz;
	const float ax = fabsf(r.d[0]), ay = fabsf(r.d[1]), az = fabsf(r.d[2]);
	s.ix = 1.0f / ax; s.iy = 1.0f / ay; s.iz = 1.0f / az;
	s.dx = (-s.ox) / ax; s.dy = (-s.oy) / ay; s.dz = (-s.oz) / az;
	s.octant = 0;
	if (s.dx < 0.0f) { s.octant += 1; s.dx += kFltMax; }
	if (s.dy < 0.0f) { s.octant += 2; s.dy += kFltMax; }
	if (s.dz < 0.0f) { s.octant += 4; s.dz += kFltMax; }
	s.octantTrips = 0;
	s.phase = kPhaseOctant;
}

// Result of one step.
enum StepResult : int { kStepContinue = 0, kStepHit = 1, kStepMiss = 2, kStepAbandoned = 3 };

// Try to enter the sub-DAG of the current octant; on failure advance to the next octant.
// raytracing.cpp:441-475 (one trip of the loop, minus the ESVO call itself) + :218-251.
CBQ_HD StepResult stepOctant(RayState& s, const SubDag* subdags)
{
	if (++s.octantTrips > 8) return kStepAbandoned;   // NaN in dx/dy/dz: the octant id would never advance
	const SubDag& sd = subdags[(uint32_t)s.octant ^ s.signBits];
	if (sd.node > 0) {
		const int h = sd.height;
		const uint32_t sizeU = 1u << h;
		const int size = (int)sizeU;
		// lowerBound * ivec3(rayDirSign) - signBit * nodeSize, wrapping (raytracing.cpp:454-456)
		const uint32_t nx = s.signBits & 1u, ny = (s.signBits >> 1) & 1u, nz = (s.signBits >> 2) & 1u;
		const int lx = (int)((nx ? (0u - (uint32_t)sd.lower[0]) : (uint32_t)sd.lower[0]) - nx * (uint32_t)size);
		const int ly = (int)((ny ? (0u - (uint32_t)sd.lower[1]) : (uint32_t)sd.lower[1]) - ny * (uint32_t)size);
		const int lz = (int)((nz ? (0u - (uint32_t)sd.lower[2]) : (uint32_t)sd.lower[2]) - nz * (uint32_t)size);
		// Slab test of the sub-DAG root (raytracing.cpp:224-232).
		const float fsz = (float)sizeU;
		const float flx = (float)lx, fly = (float)ly, flz = (float)lz;
		const float t0x = (flx - s.ox) * s.ix, t0y = (fly - s.oy) * s.iy, t0z = (flz - s.oz) * s.iz;
		const float t1x = ((flx + fsz) - s.ox) * s.ix, t1y = ((fly + fsz) - s.oy) * s.iy, t1z = ((flz + fsz) - s.oz) * s.iz;
		const float entry = max3(t0x, t0y, t0z);
		const float exit = min3(t1x, t1y, t1z);
		if (entry < exit) {
			s.startHeight = h;
			s.height = h;
			s.node = sd.node;
			s.childSize = (int)(sizeU / 2);
			const uint32_t half = (uint32_t)s.childSize;
			const uint32_t id = firstChild(entry, s, (int)((uint32_t)lx + half), (int)((uint32_t)ly + half), (int)((uint32_t)lz + half));
			s.idBits = id;
			s.px = (int)((uint32_t)lx + ((id & 1u) ? half : 0u));
			s.py = (int)((uint32_t)ly + ((id & 2u) ? half : 0u));
			s.pz = (int)((uint32_t)lz + ((id & 4u) ? half : 0u));
			s.lastExit = exit;
			s.trips = 0;
			s.phase = kPhaseEsvo;
			return kStepContinue;
		}
	}
	// Next octant (raytracing.cpp:470-475).
	const float nearest = min3(s.dx, s.dy, s.dz);
	if (s.dx <= nearest) { s.octant += 1; s.dx += kFltMax; }
	if (s.dy <= nearest) { s.octant += 2; s.dy += kFltMax; }
	if (s.dz <= nearest) { s.octant += 4; s.dz += kFltMax; }
	return (s.octant <= 7) ? kStepContinue : kStepMiss;
}

// Leaving a sub-DAG without a hit: raytracing.cpp:463-475 when intersection.hit is false.
CBQ_HD StepResult leaveSubDag(RayState& s)
{
	const float nearest = min3(s.dx, s.dy, s.dz);
	if (s.dx <= nearest) { s.octant += 1; s.dx += kFltMax; }
	if (s.dy <= nearest) { s.octant += 2; s.dy += kFltMax; }
	if (s.dz <= nearest) { s.octant += 4; s.dz += kFltMax; }
	s.phase = kPhaseOctant;
	return (s.octant <= 7) ? kStepContinue : kStepMiss;
}

// One trip round the ESVO loop, raytracing.cpp:253-367.
//   Stack: load(height) / store(height, node); heights 0..32.
//   On kStepHit: tEntry and the child entry times are left in `out` (distance, normal, material).
template <typename Nodes, typename Stack>
CBQ_HD StepResult stepEsvo(RayState& s, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	if (++s.trips > iterationCap(s.startHeight)) return kStepAbandoned;

	const uint32_t cs = (uint32_t)s.childSize;
	const float c1x = ((float)(int)((uint32_t)s.px + cs) - s.ox) * s.ix;
	const float c1y = ((float)(int)((uint32_t)s.py + cs) - s.oy) * s.iy;
	const float c1z = ((float)(int)((uint32_t)s.pz + cs) - s.oz) * s.iz;
	const float tExit = min3(c1x, c1y, c1z);

	const uint32_t child = nodes.child(s.node, s.idBits ^ s.signBits);

	if (child > 0) {
		const float c0x = ((float)s.px - s.ox) * s.ix;
		const float c0y = ((float)s.py - s.oy) * s.iy;
		const float c0z = ((float)s.pz - s.oz) * s.iz;
		const float tEntry = max3(c0x, c0y, c0z);
		const bool internal = child >= kMaterialCount;
		const bool bigEnough = ((float)s.childSize / tExit) > maxFootprint;
		if (internal && bigEnough) {
			// PUSH (raytracing.cpp:279-301)
			if (tExit < s.lastExit) stack.store(s.height, s.node);
			s.lastExit = tExit;
			s.height--;
			s.node = child;
			s.childSize /= 2;
			const uint32_t half = (uint32_t)s.childSize;
			const uint32_t id = firstChild(tEntry, s, (int)((uint32_t)s.px + half), (int)((uint32_t)s.py + half), (int)((uint32_t)s.pz + half));
			s.idBits = id;
			s.px = (int)((uint32_t)s.px + ((id & 1u) ? half : 0u));
			s.py = (int)((uint32_t)s.py + ((id & 2u) ? half : 0u));
			s.pz = (int)((uint32_t)s.pz + ((id & 4u) ? half : 0u));
			return kStepContinue;
		}
		// HIT (raytracing.cpp:302-320)
		out.hit = 1;
		out.distance = tEntry;
		if (kSurface) {
			out.material = nearestMaterial(nodes, child, s.signBits);
			if (out.material == kNoMaterial) return kStepAbandoned;
			const float sx = (s.signBits & 1u) ? -1.0f : 1.0f, sy = (s.signBits & 2u) ? -1.0f : 1.0f, sz = (s.signBits & 4u) ? -1.0f : 1.0f;
			out.normal[0] = ((tEntry == c0x) ? 1.0f : 0.0f) * (-sx);
			out.normal[1] = ((tEntry == c0y) ? 1.0f : 0.0f) * (-sy);
			out.normal[2] = ((tEntry == c0z) ? 1.0f : 0.0f) * (-sz);
		}
		return kStepHit;
	}

	// ADVANCE (raytracing.cpp:325-332)
	const uint32_t flips = (c1x <= tExit ? 1u : 0u) | (c1y <= tExit ? 2u : 0u) | (c1z <= tExit ? 4u : 0u);
	const uint32_t newId = s.idBits ^ flips;
	const int oldx = s.px, oldy = s.py, oldz = s.pz;
	s.px = (int)((uint32_t)s.px + ((flips & 1u) ? cs : 0u));
	s.py = (int)((uint32_t)s.py + ((flips & 2u) ? cs : 0u));
	s.pz = (int)((uint32_t)s.pz + ((flips & 4u) ? cs : 0u));
	s.idBits = newId;
	if ((newId & flips) != flips) {
		// POP (raytracing.cpp:339-364)
		const uint32_t diff = (uint32_t)(oldx ^ s.px) | (uint32_t)(oldy ^ s.py) | (uint32_t)(oldz ^ s.pz);
		const int msb = findMsb(diff);
		s.height = msb + 1;
		if (s.height > s.startHeight) {
			// Climbed out of the sub-DAG: the reference's loop condition fails (raytracing.cpp:367)
			// and nothing it computed after findMSB is used.
			return leaveSubDag(s);
		}
		s.node = stack.load(s.height);
		s.childSize = (int)(1u << msb);
		const uint32_t bx = ((uint32_t)(s.px >> msb)) & 1u, by = ((uint32_t)(s.py >> msb)) & 1u, bz = ((uint32_t)(s.pz >> msb)) & 1u;
		s.idBits = bx | (by << 1) | (bz << 2);
		const uint32_t big = (uint32_t)s.childSize;
		s.px = (int)((((uint32_t)(s.px >> s.height)) << s.height) + (bx ? big : 0u));
		s.py = (int)((((uint32_t)(s.py >> s.height)) << s.height) + (by ? big : 0u));
		s.pz = (int)((((uint32_t)(s.pz >> s.height)) << s.height) + (bz ? big : 0u));
		s.lastExit = 0.0f;
	}
	return kStepContinue;
}


// ------------------------------------------------------------------------------------------------
// V2 of the step functions: the same traversal, bit for bit, with far fewer instructions per step.
//
// The reference recomputes childT1 = (float(childPos + size) - o) * inv on every trip and childT0,
// plus the mid-plane times of findFirstChild, whenever a child is occupied (raytracing.cpp:259,270,181).
// Every one of those is a pure function of ONE integer plane coordinate, so a value computed once
// for a plane is bit-identical to any later recomputation for the same plane. V2 therefore carries
// the current child's lower/upper plane times (L, U) in registers and derives the next child's by
// SELECTION:
//   descend into sub-child bit b:  (L', U') = b ? (M, U) : (L, M)     (M = centre-plane times)
//   advance across a flipped axis: L' = U, U' = time of the next plane  (one new plane per flip)
//   pop:                           recompute both from the integers
// An advance costs no multiplies for un-flipped axes and a descend three plane evaluations instead
// of nine. The LOD test `float(size) / tExit > maxFootprint` keeps its IEEE division except when
// maxFootprint is exactly -1 (MAX_FOOTPRINT_DISABLED, raytracing.h:71), where the quotient's
// comparison with -1 is decided exactly by comparing |tExit| with size (see lodOffTest).
// tests/test_host_core.py checks V2 against the oracle on the host; the GPU tests do so on device.

CBQ_HD uint32_t floatBits(float f)
{
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, sizeof(u)); return u;
#endif
}

CBQ_HD float planeT(int plane, float o, float inv) { return ((float)plane - o) * inv; }

// Exact value of `((float)size / t) > -1.0f` for size = 2^k > 0 without dividing:
//   t > 0 or t == +0  -> quotient >= +0 or +inf          -> true
//   t == -0           -> -inf                            -> false
//   t < 0             -> RN(size / t) > -1  <=>  |t| > size   (size is a power of two, so the only
//                        float quotients that round to <= -1 are those with |t| <= size)
//   NaN               -> false
CBQ_HD bool lodOffTest(float t, int size)
{
	return (t > 0.0f) || (floatBits(t) == 0u) || ((-t) > (float)size);
}

// stepOctant for V2: identical control flow, additionally leaves L/U of the first child in `s`.
CBQ_HD StepResult stepOctant2(RayState& s, const SubDag* subdags)
{
	if (++s.octantTrips > 8) return kStepAbandoned;
	const SubDag& sd = subdags[(uint32_t)s.octant ^ s.signBits];
	if (sd.node > 0) {
		const int h = sd.height;
		const uint32_t sizeU = 1u << h;
		const uint32_t nx = s.signBits & 1u, ny = (s.signBits >> 1) & 1u, nz = (s.signBits >> 2) & 1u;
		const int lx = (int)((nx ? (0u - (uint32_t)sd.lower[0]) : (uint32_t)sd.lower[0]) - nx * sizeU);
		const int ly = (int)((ny ? (0u - (uint32_t)sd.lower[1]) : (uint32_t)sd.lower[1]) - ny * sizeU);
		const int lz = (int)((nz ? (0u - (uint32_t)sd.lower[2]) : (uint32_t)sd.lower[2]) - nz * sizeU);
		const float fsz = (float)sizeU;
		const float flx = (float)lx, fly = (float)ly, flz = (float)lz;
		const float t0x = (flx - s.ox) * s.ix, t0y = (fly - s.oy) * s.iy, t0z = (flz - s.oz) * s.iz;
		const float t1x = ((flx + fsz) - s.ox) * s.ix, t1y = ((fly + fsz) - s.oy) * s.iy, t1z = ((flz + fsz) - s.oz) * s.iz;
		const float entry = max3(t0x, t0y, t0z);
		const float exit = min3(t1x, t1y, t1z);
		if (entry < exit) {
			s.startHeight = h;
			s.height = h;
			s.node = sd.node;
			const uint32_t half = sizeU / 2;
			s.childSize = (int)half;
			const int cx = (int)((uint32_t)lx + half), cy = (int)((uint32_t)ly + half), cz = (int)((uint32_t)lz + half);
			const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
			const float mx = (fx - s.ox) * s.ix, my = (fy - s.oy) * s.iy, mz = (fz - s.oz) * s.iz;
			bool bx = mx < entry, by = my < entry, bz = mz < entry;
			if (entry <= 0.0f) { bx |= (s.ox >= fx); by |= (s.oy >= fy); bz |= (s.oz >= fz); }
			// The far planes of the upper children are float(int(l + half + half)), which is what the loop's
			// childT1 would compute -- NOT the slab test's float(l) + float(size) above.
			const float ux = planeT((int)((uint32_t)cx + half), s.ox, s.ix);
			const float uy = planeT((int)((uint32_t)cy + half), s.oy, s.iy);
			const float uz = planeT((int)((uint32_t)cz + half), s.oz, s.iz);
			s.px = bx ? cx : lx; s.py = by ? cy : ly; s.pz = bz ? cz : lz;
			s.Lx = bx ? mx : t0x; s.Ly = by ? my : t0y; s.Lz = bz ? mz : t0z;
			s.Ux = bx ? ux : mx; s.Uy = by ? uy : my; s.Uz = bz ? uz : mz;
			s.idBits = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
			s.lastExit = exit;
			s.trips = iterationCap(h);      // counts DOWN in V2
			s.child = kNeedFetch;
			s.phase = kPhaseEsvo;
			return kStepContinue;
		}
	}
	const float nearest = min3(s.dx, s.dy, s.dz);
	if (s.dx <= nearest) { s.octant += 1; s.dx += kFltMax; }
	if (s.dy <= nearest) { s.octant += 2; s.dy += kFltMax; }
	if (s.dz <= nearest) { s.octant += 4; s.dz += kFltMax; }
	return (s.octant <= 7) ? kStepContinue : kStepMiss;
}

// `child` is the node word of the current position, nodes.child(s.node, s.idBits ^ s.signBits). The
// caller supplies it so that a kernel can issue that load at the END of the previous step (right after
// the position changed) and hide its latency behind the loop's bookkeeping: fetchNext() below.
template <typename Nodes>
CBQ_HD uint32_t fetchNext(const RayState& s, const Nodes& nodes) { return nodes.child(s.node, s.idBits ^ s.signBits); }

template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD StepResult stepEsvo2(RayState& s, const uint32_t child, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	if (s.trips == 0u) return kStepAbandoned;
	s.trips--;

	const float tExit = min3(s.Ux, s.Uy, s.Uz);

	if (child > 0) {
		const float tEntry = max3(s.Lx, s.Ly, s.Lz);
		const bool internal = child >= kMaterialCount;
		const bool bigEnough = kLodOff ? lodOffTest(tExit, s.childSize) : (((float)s.childSize / tExit) > maxFootprint);
		if (internal && bigEnough) {
			// PUSH (raytracing.cpp:279-301)
			nodes.prefetch(child);
			if (tExit < s.lastExit) stack.store(s.height, s.node);
			s.lastExit = tExit;
			s.height--;
			s.node = child;
			const uint32_t half = (uint32_t)s.childSize >> 1;   // childSize > 0, so >> 1 == / 2
			s.childSize = (int)half;
			const int cx = (int)((uint32_t)s.px + half), cy = (int)((uint32_t)s.py + half), cz = (int)((uint32_t)s.pz + half);
			const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
			const float mx = (fx - s.ox) * s.ix, my = (fy - s.oy) * s.iy, mz = (fz - s.oz) * s.iz;
			bool bx = mx < tEntry, by = my < tEntry, bz = mz < tEntry;
			if (tEntry <= 0.0f) { bx |= (s.ox >= fx); by |= (s.oy >= fy); bz |= (s.oz >= fz); }
			s.px = bx ? cx : s.px; s.py = by ? cy : s.py; s.pz = bz ? cz : s.pz;
			s.Lx = bx ? mx : s.Lx; s.Ly = by ? my : s.Ly; s.Lz = bz ? mz : s.Lz;
			s.Ux = bx ? s.Ux : mx; s.Uy = by ? s.Uy : my; s.Uz = bz ? s.Uz : mz;
			s.idBits = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
			return kStepContinue;
		}
		// HIT (raytracing.cpp:302-320)
		out.hit = 1;
		out.distance = tEntry;
		if (kSurface) {
			out.material = nearestMaterial(nodes, child, s.signBits);
			if (out.material == kNoMaterial) return kStepAbandoned;
			const float sx = (s.signBits & 1u) ? -1.0f : 1.0f, sy = (s.signBits & 2u) ? -1.0f : 1.0f, sz = (s.signBits & 4u) ? -1.0f : 1.0f;
			out.normal[0] = ((tEntry == s.Lx) ? 1.0f : 0.0f) * (-sx);
			out.normal[1] = ((tEntry == s.Ly) ? 1.0f : 0.0f) * (-sy);
			out.normal[2] = ((tEntry == s.Lz) ? 1.0f : 0.0f) * (-sz);
		}
		return kStepHit;
	}

	// ADVANCE (raytracing.cpp:325-332)
	const bool fx = s.Ux <= tExit, fy = s.Uy <= tExit, fz = s.Uz <= tExit;
	const uint32_t flips = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
	const uint32_t newId = s.idBits ^ flips;
	const uint32_t cs = (uint32_t)s.childSize;
	const int oldx = s.px, oldy = s.py, oldz = s.pz;
	s.px = (int)((uint32_t)s.px + (fx ? cs : 0u));
	s.py = (int)((uint32_t)s.py + (fy ? cs : 0u));
	s.pz = (int)((uint32_t)s.pz + (fz ? cs : 0u));
	const bool stayed = (s.idBits & flips) == 0u;   // == ((newId & flips) == flips): no flipped axis was already at bit 1
	s.idBits = newId;
	if (stayed) {
		// Stayed inside the parent: a flipped axis' old upper plane is its new lower plane.
		const float nx = planeT((int)((uint32_t)s.px + cs), s.ox, s.ix);
		const float ny = planeT((int)((uint32_t)s.py + cs), s.oy, s.iy);
		const float nz = planeT((int)((uint32_t)s.pz + cs), s.oz, s.iz);
		s.Lx = fx ? s.Ux : s.Lx; s.Ly = fy ? s.Uy : s.Ly; s.Lz = fz ? s.Uz : s.Lz;
		s.Ux = fx ? nx : s.Ux; s.Uy = fy ? ny : s.Uy; s.Uz = fz ? nz : s.Uz;
		return kStepContinue;
	}
	// POP (raytracing.cpp:339-364)
	const uint32_t diff = (uint32_t)(oldx ^ s.px) | (uint32_t)(oldy ^ s.py) | (uint32_t)(oldz ^ s.pz);
	const int msb = findMsb(diff);
	s.height = msb + 1;
	if (s.height > s.startHeight) return leaveSubDag(s);
	s.node = stack.load(s.height);
	const uint32_t big = 1u << msb;
	s.childSize = (int)big;
	// The reference re-derives childId = (pos >> msb) & 1 and childPos = ((pos >> height) << height) +
	// childId * size (raytracing.cpp:359-361). Aligning to 2^height and adding back bit `msb` is the same
	// as clearing the bits BELOW msb, so: pos &= -size.
	const uint32_t keep = 0u - big;
	s.idBits = (((uint32_t)s.px >> msb) & 1u) | ((((uint32_t)s.py >> msb) & 1u) << 1) | ((((uint32_t)s.pz >> msb) & 1u) << 2);
	s.px = (int)((uint32_t)s.px & keep);
	s.py = (int)((uint32_t)s.py & keep);
	s.pz = (int)((uint32_t)s.pz & keep);
	s.lastExit = 0.0f;
	s.Lx = planeT(s.px, s.ox, s.ix); s.Ly = planeT(s.py, s.oy, s.iy); s.Lz = planeT(s.pz, s.oz, s.iz);
	s.Ux = planeT((int)((uint32_t)s.px + big), s.ox, s.ix);
	s.Uy = planeT((int)((uint32_t)s.py + big), s.oy, s.iy);
	s.Uz = planeT((int)((uint32_t)s.pz + big), s.oz, s.iz);
	return kStepContinue;
}


// ------------------------------------------------------------------------------------------------
// V3: the V2 trip split into two SECTIONS that a warp executes back to back,
//     descendSection: lanes whose current child is OCCUPIED descend (or hit),
//     advanceSection: lanes whose current child is EMPTY advance (or pop),
// each starting with "fetch the child word unless it is already known". In V2 a lane performs one
// event per warp step while the warp pays for BOTH divergent paths every step; here a lane that
// descends and lands on an empty child advances in the same trip (and an advance followed by a
// descend takes consecutive sections too), so the same instruction stream retires ~1.4 events per
// lane. One fetch == one trip of the reference's loop (raytracing.cpp:253-367), so the trip budget
// and every arithmetic result are exactly those of V2.
// MEASURED (profiles/r01_analysis.md): a warp-scheduling simulation predicted -13 % instructions, the
// B200 says 3.58 vs 4.23 Grays/s -- each section exposes its own load->use wait, and 63 registers. The
// kernels therefore use V2; V3 stays here (host-checked, -DCBQ_TRIP_V3) as a recorded negative result.

template <typename Nodes>
CBQ_HD bool fetchChild(RayState& s, const Nodes& nodes)
{
	if (s.child != kNeedFetch) return true;
	if (s.trips == 0u) return false;          // abandoned
	s.trips--;
	s.child = nodes.child(s.node, s.idBits ^ s.signBits);
	return true;
}

template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD StepResult descendSection(RayState& s, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	if (!fetchChild(s, nodes)) return kStepAbandoned;
	const uint32_t child = s.child;
	if (child == 0u) return kStepContinue;    // the advance section's business
	const float tExit = min3(s.Ux, s.Uy, s.Uz);
	const float tEntry = max3(s.Lx, s.Ly, s.Lz);
	const bool internal = child >= kMaterialCount;
	const bool bigEnough = kLodOff ? lodOffTest(tExit, s.childSize) : (((float)s.childSize / tExit) > maxFootprint);
	if (internal && bigEnough) {
		// PUSH (raytracing.cpp:279-301)
		if (tExit < s.lastExit) stack.store(s.height, s.node);
		s.lastExit = tExit;
		s.height--;
		s.node = child;
		const uint32_t half = (uint32_t)s.childSize >> 1;
		s.childSize = (int)half;
		const int cx = (int)((uint32_t)s.px + half), cy = (int)((uint32_t)s.py + half), cz = (int)((uint32_t)s.pz + half);
		const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
		const float mx = (fx - s.ox) * s.ix, my = (fy - s.oy) * s.iy, mz = (fz - s.oz) * s.iz;
		bool bx = mx < tEntry, by = my < tEntry, bz = mz < tEntry;
		if (tEntry <= 0.0f) { bx |= (s.ox >= fx); by |= (s.oy >= fy); bz |= (s.oz >= fz); }
		s.px = bx ? cx : s.px; s.py = by ? cy : s.py; s.pz = bz ? cz : s.pz;
		s.Lx = bx ? mx : s.Lx; s.Ly = by ? my : s.Ly; s.Lz = bz ? mz : s.Lz;
		s.Ux = bx ? s.Ux : mx; s.Uy = by ? s.Uy : my; s.Uz = bz ? s.Uz : mz;
		s.idBits = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
		s.child = kNeedFetch;
		return kStepContinue;
	}
	// HIT (raytracing.cpp:302-320)
	out.hit = 1;
	out.distance = tEntry;
	if (kSurface) {
		out.material = nearestMaterial(nodes, child, s.signBits);
		if (out.material == kNoMaterial) return kStepAbandoned;
		const float sx = (s.signBits & 1u) ? -1.0f : 1.0f, sy = (s.signBits & 2u) ? -1.0f : 1.0f, sz = (s.signBits & 4u) ? -1.0f : 1.0f;
		out.normal[0] = ((tEntry == s.Lx) ? 1.0f : 0.0f) * (-sx);
		out.normal[1] = ((tEntry == s.Ly) ? 1.0f : 0.0f) * (-sy);
		out.normal[2] = ((tEntry == s.Lz) ? 1.0f : 0.0f) * (-sz);
	}
	return kStepHit;
}

template <typename Nodes, typename Stack>
CBQ_HD StepResult advanceSection(RayState& s, const Nodes& nodes, Stack& stack)
{
	if (!fetchChild(s, nodes)) return kStepAbandoned;
	if (s.child != 0u) return kStepContinue;  // occupied: the next descend section's business
	s.child = kNeedFetch;
	// ADVANCE (raytracing.cpp:325-332)
	const float tExit = min3(s.Ux, s.Uy, s.Uz);
	const bool fx = s.Ux <= tExit, fy = s.Uy <= tExit, fz = s.Uz <= tExit;
	const uint32_t flips = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
	const uint32_t cs = (uint32_t)s.childSize;
	const int oldx = s.px, oldy = s.py, oldz = s.pz;
	s.px = (int)((uint32_t)s.px + (fx ? cs : 0u));
	s.py = (int)((uint32_t)s.py + (fy ? cs : 0u));
	s.pz = (int)((uint32_t)s.pz + (fz ? cs : 0u));
	const bool stayed = (s.idBits & flips) == 0u;
	s.idBits ^= flips;
	if (stayed) {
		const float nx = planeT((int)((uint32_t)s.px + cs), s.ox, s.ix);
		const float ny = planeT((int)((uint32_t)s.py + cs), s.oy, s.iy);
		const float nz = planeT((int)((uint32_t)s.pz + cs), s.oz, s.iz);
		s.Lx = fx ? s.Ux : s.Lx; s.Ly = fy ? s.Uy : s.Ly; s.Lz = fz ? s.Uz : s.Lz;
		s.Ux = fx ? nx : s.Ux; s.Uy = fy ? ny : s.Uy; s.Uz = fz ? nz : s.Uz;
		return kStepContinue;
	}
	// POP (raytracing.cpp:339-364)
	const uint32_t diff = (uint32_t)(oldx ^ s.px) | (uint32_t)(oldy ^ s.py) | (uint32_t)(oldz ^ s.pz);
	const int msb = findMsb(diff);
	s.height = msb + 1;
	if (s.height > s.startHeight) return leaveSubDag(s);
	s.node = stack.load(s.height);
	const uint32_t big = 1u << msb;
	s.childSize = (int)big;
	const uint32_t keep = 0u - big;
	s.idBits = (((uint32_t)s.px >> msb) & 1u) | ((((uint32_t)s.py >> msb) & 1u) << 1) | ((((uint32_t)s.pz >> msb) & 1u) << 2);
	s.px = (int)((uint32_t)s.px & keep);
	s.py = (int)((uint32_t)s.py & keep);
	s.pz = (int)((uint32_t)s.pz & keep);
	s.lastExit = 0.0f;
	s.Lx = planeT(s.px, s.ox, s.ix); s.Ly = planeT(s.py, s.oy, s.iy); s.Lz = planeT(s.pz, s.oz, s.iz);
	s.Ux = planeT((int)((uint32_t)s.px + big), s.ox, s.ix);
	s.Uy = planeT((int)((uint32_t)s.py + big), s.oy, s.iy);
	s.Uz = planeT((int)((uint32_t)s.pz + big), s.oz, s.iz);
	return kStepContinue;
}

// One V3 trip for a lane that is inside a sub-DAG: descend section, then advance section.
template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD StepResult tripEsvo3(RayState& s, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	StepResult res = descendSection<kLodOff>(s, nodes, stack, maxFootprint, kSurface, out);
	if (res == kStepContinue && s.phase == kPhaseEsvo) res = advanceSection(s, nodes, stack);
	return res;
}

// Fill in the fields intersectVolume adds after a hit (raytracing.cpp:463-466).
CBQ_HD void finishHit(Hit& out, const Ray& r)
{
	out.position[0] = r.o[0] + (r.d[0] * out.distance);
	out.position[1] = r.o[1] + (r.d[1] * out.distance);
	out.position[2] = r.o[2] + (r.d[2] * out.distance);
}

CBQ_HD void clearHit(Hit& h)
{
	h.hit = 0; h.distance = 0.0f; h.material = 0;
	h.position[0] = h.position[1] = h.position[2] = 0.0f;
	h.normal[0] = h.normal[1] = h.normal[2] = 0.0f;
	h.status = 0;
}

// Whole ray, start to finish (used by the path tracer's megakernel-style helpers and by the
// host check). Returns with `out` complete.
template <bool kSurface, typename Nodes, typename Stack>
CBQ_HD void traceRay(const Ray& r, const Nodes& nodes, const SubDag* subdags, Stack& stack, float maxFootprint, Hit& out)
{
	clearHit(out);
	RayState s;
	beginRay(s, r);
	for (;;) {
		StepResult res;
		if (s.phase == kPhaseOctant) res = stepOctant(s, subdags);
		else res = stepEsvo(s, nodes, stack, maxFootprint, kSurface, out);
		if (res == kStepContinue) continue;
		if (res == kStepHit) { finishHit(out, r); return; }
		if (res == kStepAbandoned) { clearHit(out); out.status = 1; return; }
		return; // miss
	}
}

// Whole ray with the V2 steps.
template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD void traceRay2(const Ray& r, const Nodes& nodes, const SubDag* subdags, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	clearHit(out);
	RayState s;
	beginRay(s, r);
	for (;;) {
		StepResult res;
		if (s.phase == kPhaseOctant) res = stepOctant2(s, subdags);
		else res = stepEsvo2<kLodOff>(s, fetchNext(s, nodes), nodes, stack, maxFootprint, kSurface, out);
		if (res == kStepContinue) continue;
		if (res == kStepHit) { finishHit(out, r); return; }
		if (res == kStepAbandoned) { clearHit(out); out.status = 1; return; }
		return;
	}
}


// Whole ray with the V3 trips.
template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD void traceRay3(const Ray& r, const Nodes& nodes, const SubDag* subdags, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	clearHit(out);
	RayState s;
	beginRay(s, r);
	for (;;) {
		StepResult res;
		if (s.phase == kPhaseOctant) res = stepOctant2(s, subdags);
		else res = tripEsvo3<kLodOff>(s, nodes, stack, maxFootprint, kSurface, out);
		if (res == kStepContinue) continue;
		if (res == kStepHit) { finishHit(out, r); return; }
		if (res == kStepAbandoned) { clearHit(out); out.status = 1; return; }
		return;
	}
}

} // namespace cbq

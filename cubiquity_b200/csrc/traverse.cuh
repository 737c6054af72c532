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
	// t-values of the current child's lower (L) and upper (U) planes, i.e. exactly the childT0 / childT1 of
	// raytracing.cpp:259,270 for the current childPos, carried between steps.
	float Lx, Ly, Lz;
	float Ux, Uy, Uz;
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

// Try to enter the sub-DAG of the current octant; on failure advance to the next octant: raytracing.cpp:441-475
// (one trip of the loop, minus the ESVO call itself) + :218-251. Leaves L/U of the first child in `s`.
template <typename Stack>
CBQ_HD StepResult stepOctant(RayState& s, const SubDag* subdags, Stack& stack)
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
			s.trips = iterationCap(h);      // counts down
			s.phase = kPhaseEsvo;
			// The reference's stack is uninitialised (raytracing.cpp:251); the oracle zero-fills it, and a pop to a level
			// that was never pushed (only possible when NaNs or collapsed float planes defeat the `tExit < lastExit`
			// guard, :285) must read that 0 rather than what an earlier ray left in the slot: clear() makes levels
			// 0..h read as 0 until they are stored to.
			stack.clear(h);
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
CBQ_HD StepResult stepEsvo(RayState& s, const uint32_t child, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
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

// Whole ray, start to finish: what a lane of the kernel does between taking a ray and writing its hit (used by the
// host check; the kernels drive the steps themselves so that lanes can be refilled between any two).
template <bool kLodOff, typename Nodes, typename Stack>
CBQ_HD void traceRay(const Ray& r, const Nodes& nodes, const SubDag* subdags, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	clearHit(out);
	RayState s;
	beginRay(s, r);
	for (;;) {
		StepResult res;
		if (s.phase == kPhaseOctant) res = stepOctant(s, subdags, stack);
		else res = stepEsvo<kLodOff>(s, fetchNext(s, nodes), nodes, stack, maxFootprint, kSurface, out);
		if (res == kStepContinue) continue;
		if (res == kStepHit) { finishHit(out, r); return; }
		if (res == kStepAbandoned) { clearHit(out); out.status = 1; return; }
		return;
	}
}

} // namespace cbq

// SVDAG ray traversal core for sm_100a -- the per-ray state machine.
//
// Computes exactly what Cubiquity::intersectVolume computes (reference
// src/library/raytracing.cpp:397-478) including intersectRayNodeESVO (:213-371),
// findFirstChild (:178-196) and findNearestMaterial (:134-162), bit for bit, but restructured
// for SIMT execution:
//
//   * The reference's two nested loops (octants, then the ESVO do/while) are flattened into ONE
//     loop of "steps" so that lanes in different octants or at different depths share an
//     instruction stream and a persistent kernel can retire a finished ray and pull a fresh one
//     into the same lane between any two steps (trace_kernels.cu).
//
//   * A step is one NODE VISIT, not one trip of the reference's loop. The reference reads one child
//     word per trip (raytracing.cpp:264-265) only to learn, most of the time, that the child is empty.
//     Here the traversal follows PACKED REFERENCES (see "Packed node references" below): the word
//     that leads to a node also carries that node's 8-bit occupancy mask, so the siblings a ray
//     crosses inside a node are tested with bit operations and the only memory access of a step
//     is the one load of the occupied child it descends into (or hits). A pop takes the ancestor's
//     mask from the stack with its index. Trips over empty siblings -- 46 % of the reference's loop
//     trips on the 1080p bench frame -- cost no load, no loop control and no lane-divergent branch.
//
//   * The state is kept per NODE: the times t0 / tm / t1 at which the ray crosses the node's lower,
//     centre and upper planes. Each is a pure function of one integer plane coordinate
//     ((float(plane) - o) * inv, raytracing.cpp:181,259,270), so reusing a value is bit-identical to
//     the reference's recomputation. The current child is three bits; its childT0 / childT1 are
//     SELECTIONS (bit ? tm : t0, bit ? t1 : tm), an advance changes no float at all, a descend
//     evaluates the 3 new centre planes, a pop re-evaluates the ancestor's 9.
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

// ------------------------------------------------------------------------------------------------
// Packed node references.
//
// The reference's node is 8 child words: 0 = empty, 1..255 = a uniform region of that material,
// >= 256 = index of an internal node (storage.h:59-70). The device keeps that array as it is (it is
// what edits, bake and download work on) and derives from it the array the ray cast reads, same
// shape, where every word is a REFERENCE:
//      0                      empty
//      1..255                 material (unchanged)
//      (index << 8) | mask    internal node `index`, mask bit c set iff its child c is non-zero
// Ref is uint32_t while every index is below 2^24 (all BASELINE volumes: the 16384^3 soup has 16.2 M
// nodes) and uint64_t otherwise; an internal reference is >= 2^16, so `ref >= 256` still means
// "internal" (raytracing.cpp:275). packNode() below is the whole transcoding rule.
template <typename Ref> CBQ_HD uint32_t refIndex(Ref r) { return (uint32_t)(r >> 8); }

CBQ_HD uint32_t occupancyMask(const uint32_t* node)
{
	uint32_t m = 0;
	for (uint32_t c = 0; c < 8; c++) m |= (node[c] != 0u ? 1u : 0u) << c;
	return m;
}

// Reference to node `index` as a traversal root (sub-DAG roots; also valid for the material nodes
// 0..255, which are real self-referencing entries, storage.cpp:110-122).
template <typename Ref>
CBQ_HD Ref nodeRef(const uint32_t* nodes, uint32_t index) { return ((Ref)index << 8) | (Ref)occupancyMask(nodes + (size_t)index * 8); }

// The reference stored in a child slot.
template <typename Ref>
CBQ_HD Ref childRef(const uint32_t* nodes, uint32_t child) { return child < kMaterialCount ? (Ref)child : nodeRef<Ref>(nodes, child); }

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

CBQ_HD uint32_t floatBits(float f)
{
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, sizeof(u)); return u;
#endif
}

CBQ_HD float planeT(int plane, float o, float inv) { return ((float)plane - o) * inv; }

// Give up after this many node visits in one sub-DAG of the given height. A legitimate traversal
// needs fewer than 8 * 2^h trips of the reference's loop (and a visit is at least one trip); the
// reference itself never terminates on the rays this catches. The common one -- a zero direction
// component with the origin exactly on a cell boundary: 0 * inf = NaN, no sibling flip is ever taken,
// raytracing.cpp:328 -- is detected directly (stepEsvo: flips == 0); the budget is the backstop for
// NaN-driven cycles through the stack. Same value as cbqo_iteration_cap in oracle/.
CBQ_HD uint32_t iterationCap(int subdagHeight)
{
	return (subdagHeight >= 20) ? (1u << 26) : ((64u << subdagHeight) + 4096u);
}

// Exact value of `((float)size / t) > -1.0f` for size = 2^k > 0 without dividing:
//   t > 0 or t == +0  -> quotient >= +0 or +inf          -> true
//   t == -0           -> -inf                            -> false
//   t < 0             -> RN(size / t) > -1  <=>  |t| > size   (size is a power of two, so the only
//                        float quotients that round to <= -1 are those with |t| <= size)
//   NaN               -> false
CBQ_HD bool lodOffTest(float t, uint32_t size)
{
	return (t > 0.0f) || (floatBits(t) == 0u) || ((-t) > (float)size);
}

enum Phase : int {
	kPhaseIdle = 0,     // lane holds no ray
	kPhaseOctant = 1,   // about to try the sub-DAG of the current root octant (raytracing.cpp:441-475)
	kPhaseEsvo = 2      // inside the do/while of raytracing.cpp:253-367
};

// Everything a ray in flight carries between steps.
template <typename Ref>
struct RayState {
	// Reflected ray (raytracing.cpp:422-425) and its reciprocal direction (:224).
	float ox, oy, oz;
	float ix, iy, iz;
	// Octant walk (raytracing.cpp:430-434, 470-473).
	float dx, dy, dz;
	int octant;
	int octantTrips;
	uint32_t signBits;
	// ESVO state (raytracing.cpp:234-251), per node.
	Ref node;              // reference to the current node: index and occupancy mask
	int nx, ny, nz;        // its lower corner, reflected space (childPos = n + idBit * childSize)
	uint32_t idBits;       // current child: childId packed x | y << 1 | z << 2
	int height;            // nodeHeight: the node is 2^height wide, its children 2^(height - 1)
	int startHeight;
	float lastExit;
	uint32_t trips;        // node visits left (iterationCap)
	int phase;
	// Crossing times of the node's lower / centre / upper planes.
	float t0x, t0y, t0z;
	float tmx, tmy, tmz;
	float t1x, t1y, t1z;
};

// findNearestMaterial, raytracing.cpp:134-162: step into the first occupied child in a fixed
// near-to-far order until a material is reached. The occupancy mask picks the child; one load per level.
template <typename Nodes, typename Ref>
CBQ_HD uint32_t nearestMaterial(const Nodes& nodes, Ref node, uint32_t signBits)
{
	const uint32_t order = (0x76534210u | 0x88888888u) ^ (signBits * 0x11111111u);
	int levels = 0;
	while (node >= (Ref)kMaterialCount) {
		const uint32_t mask = (uint32_t)node & 0xffu;
		bool found = false;
		for (uint32_t ids = order; ids != 0; ids >>= 4) {
			if ((mask >> (ids & 7u)) & 1u) { node = nodes.child(node, ids & 7u); found = true; break; }
		}
		// An internal node with eight empty children (legal in an edited, un-baked volume: fillBrush
		// never collapses nodes) makes the reference spin here for ever; we report "no material" and
		// the caller abandons the ray. Must match nearest_material in oracle/cbq_oracle.c.
		if (!found || ++levels > 32) return kNoMaterial;
	}
	return (uint32_t)node;
}

// Set up a ray: raytracing.cpp:407-434.
template <typename Ref>
CBQ_HD void beginRay(RayState<Ref>& s, const Ray& r)
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

// Next root octant: raytracing.cpp:470-475.
template <typename Ref>
CBQ_HD StepResult nextOctant(RayState<Ref>& s)
{
	const float nearest = min3(s.dx, s.dy, s.dz);
	if (s.dx <= nearest) { s.octant += 1; s.dx += kFltMax; }
	if (s.dy <= nearest) { s.octant += 2; s.dy += kFltMax; }
	if (s.dz <= nearest) { s.octant += 4; s.dz += kFltMax; }
	s.phase = kPhaseOctant;
	return (s.octant <= 7) ? kStepContinue : kStepMiss;
}

// Try to enter the sub-DAG of the current octant; on failure advance to the next octant.
// raytracing.cpp:441-475 (one trip of the loop, minus the ESVO call itself) + :218-251.
// rootRefs[i] is the packed reference of subdags[i].node.
template <typename Ref, typename Stack>
CBQ_HD StepResult stepOctant(RayState<Ref>& s, const SubDag* subdags, const Ref* rootRefs, Stack& stack)
{
	if (++s.octantTrips > 8) return kStepAbandoned;   // NaN in dx/dy/dz: the octant id would never advance
	const uint32_t which = (uint32_t)s.octant ^ s.signBits;
	const SubDag& sd = subdags[which];
	if (sd.node > 0) {
		const int h = sd.height;
		const uint32_t sizeU = 1u << h;
		// lowerBound * ivec3(rayDirSign) - signBit * nodeSize, wrapping (raytracing.cpp:454-456)
		const uint32_t nx = s.signBits & 1u, ny = (s.signBits >> 1) & 1u, nz = (s.signBits >> 2) & 1u;
		const int lx = (int)((nx ? (0u - (uint32_t)sd.lower[0]) : (uint32_t)sd.lower[0]) - nx * sizeU);
		const int ly = (int)((ny ? (0u - (uint32_t)sd.lower[1]) : (uint32_t)sd.lower[1]) - ny * sizeU);
		const int lz = (int)((nz ? (0u - (uint32_t)sd.lower[2]) : (uint32_t)sd.lower[2]) - nz * sizeU);
		// Slab test of the sub-DAG root (raytracing.cpp:224-232): the upper planes are float(l) + float(size) HERE.
		const float fsz = (float)sizeU;
		const float flx = (float)lx, fly = (float)ly, flz = (float)lz;
		const float t0x = (flx - s.ox) * s.ix, t0y = (fly - s.oy) * s.iy, t0z = (flz - s.oz) * s.iz;
		const float e1x = ((flx + fsz) - s.ox) * s.ix, e1y = ((fly + fsz) - s.oy) * s.iy, e1z = ((flz + fsz) - s.oz) * s.iz;
		const float entry = max3(t0x, t0y, t0z);
		const float exit = min3(e1x, e1y, e1z);
		if (entry < exit) {
			s.startHeight = h;
			s.height = h;
			s.node = rootRefs[which];
			s.nx = lx; s.ny = ly; s.nz = lz;
			const uint32_t half = sizeU >> 1;
			const int cx = (int)((uint32_t)lx + half), cy = (int)((uint32_t)ly + half), cz = (int)((uint32_t)lz + half);
			const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
			s.t0x = t0x; s.t0y = t0y; s.t0z = t0z;
			s.tmx = (fx - s.ox) * s.ix; s.tmy = (fy - s.oy) * s.iy; s.tmz = (fz - s.oz) * s.iz;
			// Inside the loop the children's far planes are float(int(l + half + half)) (raytracing.cpp:259), not the
			// slab test's float(l) + float(size).
			s.t1x = planeT((int)((uint32_t)cx + half), s.ox, s.ix);
			s.t1y = planeT((int)((uint32_t)cy + half), s.oy, s.iy);
			s.t1z = planeT((int)((uint32_t)cz + half), s.oz, s.iz);
			// findFirstChild (raytracing.cpp:178-196)
			bool bx = s.tmx < entry, by = s.tmy < entry, bz = s.tmz < entry;
			if (entry <= 0.0f) { bx |= (s.ox >= fx); by |= (s.oy >= fy); bz |= (s.oz >= fz); }
			s.idBits = (bx ? 1u : 0u) | (by ? 2u : 0u) | (bz ? 4u : 0u);
			s.lastExit = exit;
			s.trips = iterationCap(h);
			s.phase = kPhaseEsvo;
			// The reference's stack is uninitialised (raytracing.cpp:251); the oracle zero-fills it, and a pop to a
			// level that was never pushed (only possible when NaNs or collapsed float planes defeat the
			// `tExit < lastExit` guard, :285) must read that 0 rather than what an earlier ray left behind.
			stack.clear(h);
			return kStepContinue;
		}
	}
	return nextOctant(s);
}

// One node visit of the ESVO loop, raytracing.cpp:253-367:
//   scan     the trips over EMPTY children (:325-332): advance from sibling to sibling by the occupancy mask until
//            an occupied child is found or the ray leaves the node;
//   occupied (:268-320) one load of the child's reference, then descend into it or report the hit;
//   pop      (:339-364) climb to the ancestor that holds the next sibling.
//   Stack: store(height, ref) / load(height) / clear(topHeight).
//   On kStepHit: distance, material and normal are left in `out`.
template <bool kLodOff, typename Ref, typename Nodes, typename Stack>
CBQ_HD StepResult stepEsvo(RayState<Ref>& s, const Nodes& nodes, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	if (s.trips == 0u) return kStepAbandoned;
	s.trips--;

	const uint32_t mask = (uint32_t)s.node;   // bits 0..7: the node's occupancy, by un-reflected child slot
	uint32_t b = s.idBits;
	float ux, uy, uz, tExit;
	uint32_t flips = 0;
	bool occupied, stuck = false;
#if defined(__CUDA_ARCH__)
	const unsigned together = __activemask();   // the lanes that entered this visit in step
#endif
	for (;;) {
		// childT1 of the current child and tChildExit (raytracing.cpp:259-260)
		ux = (b & 1u) ? s.t1x : s.tmx; uy = (b & 2u) ? s.t1y : s.tmy; uz = (b & 4u) ? s.t1z : s.tmz;
		tExit = min3(ux, uy, uz);
		occupied = ((mask >> (b ^ s.signBits)) & 1u) != 0u;
		if (occupied) break;
		// ADVANCE (raytracing.cpp:325-337)
		flips = (ux <= tExit ? 1u : 0u) | (uy <= tExit ? 2u : 0u) | (uz <= tExit ? 4u : 0u);
		stuck = flips == 0u;                        // NaN planes: the reference's loop makes no progress from here on
		if ((b & flips) != 0u || stuck) break;      // a flipped axis was already at its upper child: left the node
		b |= flips;
	}
#if defined(__CUDA_ARCH__)
	// Lanes leave the scan after 0..3 advances; without this the compiler lets each group run the rest of the
	// visit on its own (measured: the descend and pop sections executed 1.6x per warp visit).
	__syncwarp(together);
#endif
	if (stuck) return kStepAbandoned;

	const uint32_t cs = (1u << s.height) >> 1;      // childNodeSize

	if (occupied) {
		const Ref child = nodes.child(s.node, b ^ s.signBits);
		// childT0 and tChildEntry (raytracing.cpp:270-271)
		const float lx = (b & 1u) ? s.tmx : s.t0x, ly = (b & 2u) ? s.tmy : s.t0y, lz = (b & 4u) ? s.tmz : s.t0z;
		const float tEntry = max3(lx, ly, lz);
		const bool internal = child >= (Ref)kMaterialCount;
		const bool bigEnough = kLodOff ? lodOffTest(tExit, cs) : (((float)(int)cs / tExit) > maxFootprint);
		if (internal && bigEnough) {
			// PUSH (raytracing.cpp:279-301)
			if (tExit < s.lastExit) stack.store(s.height, s.node);
			s.lastExit = tExit;
			s.height--;
			s.node = child;
			s.nx = (int)((uint32_t)s.nx + ((b & 1u) ? cs : 0u));
			s.ny = (int)((uint32_t)s.ny + ((b & 2u) ? cs : 0u));
			s.nz = (int)((uint32_t)s.nz + ((b & 4u) ? cs : 0u));
			const uint32_t half = cs >> 1;
			const int cx = (int)((uint32_t)s.nx + half), cy = (int)((uint32_t)s.ny + half), cz = (int)((uint32_t)s.nz + half);
			const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
			s.t0x = lx; s.t0y = ly; s.t0z = lz;
			s.t1x = ux; s.t1y = uy; s.t1z = uz;
			s.tmx = (fx - s.ox) * s.ix; s.tmy = (fy - s.oy) * s.iy; s.tmz = (fz - s.oz) * s.iz;
			// findFirstChild (raytracing.cpp:178-196)
			bool bx = s.tmx < tEntry, by = s.tmy < tEntry, bz = s.tmz < tEntry;
			if (tEntry <= 0.0f) { bx |= (s.ox >= fx); by |= (s.oy >= fy); bz |= (s.oz >= fz); }
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
			out.normal[0] = ((tEntry == lx) ? 1.0f : 0.0f) * (-sx);
			out.normal[1] = ((tEntry == ly) ? 1.0f : 0.0f) * (-sy);
			out.normal[2] = ((tEntry == lz) ? 1.0f : 0.0f) * (-sz);
		}
		return kStepHit;
	}

	// POP (raytracing.cpp:339-364). The reference XORs the child position before and after the flip. An axis that
	// flipped from the lower to the upper child differs in bit `cs` only; an axis that wrapped went from n + cs to
	// n + 2 cs, which differs from n ^ (n + 2 cs) in that same bit only. The highest differing bit is therefore that
	// of the wrapped axes' n ^ (n + 2 cs), and after the shifts below only the wrapped axes have moved.
	const uint32_t wrapped = b & flips;
	const uint32_t size = cs << 1;
	const uint32_t qx = (uint32_t)s.nx + ((wrapped & 1u) ? size : 0u);
	const uint32_t qy = (uint32_t)s.ny + ((wrapped & 2u) ? size : 0u);
	const uint32_t qz = (uint32_t)s.nz + ((wrapped & 4u) ? size : 0u);
	const uint32_t diff = ((uint32_t)s.nx ^ qx) | ((uint32_t)s.ny ^ qy) | ((uint32_t)s.nz ^ qz);
	const int msb = findMsb(diff);
	const int h = msb + 1;
	// Climbed out of the sub-DAG: the reference's loop condition fails (raytracing.cpp:367) and nothing it
	// computed after findMSB is used.
	if (h > s.startHeight) return nextOctant(s);
	s.height = h;
	s.node = stack.load(h);
	const uint32_t big = 1u << msb;                  // the ancestor's child size; msb <= 30 here
	// childId = (pos >> msb) & 1; childPos = ((pos >> height) << height) + childId * size (raytracing.cpp:359-361)
	s.idBits = ((qx >> msb) & 1u) | (((qy >> msb) & 1u) << 1) | (((qz >> msb) & 1u) << 2);
	const uint32_t keep = 0u - (big << 1);
	s.nx = (int)(qx & keep); s.ny = (int)(qy & keep); s.nz = (int)(qz & keep);
	s.lastExit = 0.0f;
	s.t0x = planeT(s.nx, s.ox, s.ix); s.t0y = planeT(s.ny, s.oy, s.iy); s.t0z = planeT(s.nz, s.oz, s.iz);
	s.tmx = planeT((int)((uint32_t)s.nx + big), s.ox, s.ix);
	s.tmy = planeT((int)((uint32_t)s.ny + big), s.oy, s.iy);
	s.tmz = planeT((int)((uint32_t)s.nz + big), s.oz, s.iz);
	s.t1x = planeT((int)((uint32_t)s.nx + (big << 1)), s.ox, s.ix);
	s.t1y = planeT((int)((uint32_t)s.ny + (big << 1)), s.oy, s.iy);
	s.t1z = planeT((int)((uint32_t)s.nz + (big << 1)), s.oz, s.iz);
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
template <bool kLodOff, typename Ref, typename Nodes, typename Stack>
CBQ_HD void traceRay(const Ray& r, const Nodes& nodes, const SubDag* subdags, const Ref* rootRefs, Stack& stack, float maxFootprint, const bool kSurface, Hit& out)
{
	clearHit(out);
	RayState<Ref> s;
	beginRay(s, r);
	for (;;) {
		StepResult res;
		if (s.phase == kPhaseOctant) res = stepOctant(s, subdags, rootRefs, stack);
		else res = stepEsvo<kLodOff>(s, nodes, stack, maxFootprint, kSurface, out);
		if (res == kStepContinue) continue;
		if (res == kStepHit) { finishHit(out, r); return; }
		if (res == kStepAbandoned) { clearHit(out); out.status = 1; return; }
		return; // miss
	}
}

} // namespace cbq

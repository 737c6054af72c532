// Hierarchical procedural SVDAG builder (host C++).
//
// The reference builds volumes one voxel at a time (Volume::setVoxel, reference
// src/library/storage.cpp:396-438, driven by commands/generate/generate.cpp:83-181) and then
// hash-conses them in Volume::bake -> NodeStore::merge (storage.cpp:208-290). That is O(size^3)
// and cannot produce the 4096^3 .. 65536^3 inputs of BASELINE.json. This builder evaluates a
// closed-form scene top-down, stops at uniform regions, and interns nodes bottom-up, emitting
// exactly the array NodeStore::rawBytesPtr() exposes (storage.h:101): 8 x u32 per node, entries
// 0..255 the self-referencing material nodes (storage.cpp:110-122), child slot
// c = x | y << 1 | z << 2 (storage.cpp:57-67), root at height 32 spanning [-2^31, 2^31)
// (storage.cpp:45-50). Like merge_node (storage.cpp:258-260) it collapses a node whose eight
// children are one material, so the result is the canonical DAG bake() would give (the node
// ORDER differs; traversal results do not depend on it).
//
// All scene maths is integer so every machine builds bit-identical volumes from a seed.
#include "cbq_scenes.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef std::array<uint32_t, 8> Node;

inline uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= UINT64_C(0xbf58476d1ce4e5b9);
	x ^= x >> 27; x *= UINT64_C(0x94d049bb133111eb);
	x ^= x >> 31;
	return x;
}

inline uint32_t hash3(int64_t x, int64_t y, int64_t z, uint64_t seed)
{
	uint64_t h = mix64(seed ^ (uint64_t)x * UINT64_C(0x9e3779b97f4a7c15));
	h = mix64(h ^ (uint64_t)y * UINT64_C(0xc2b2ae3d27d4eb4f));
	h = mix64(h ^ (uint64_t)z * UINT64_C(0x165667b19e3779f9));
	return (uint32_t)(h >> 32);
}

// ------------------------------------------------------------------ node pool

class NodePool {
public:
	NodePool()
	{
		nodes_.resize(256);
		for (uint32_t i = 0; i < 256; i++) nodes_[i].fill(i);
		table_.assign(1u << 16, 0);
	}

	uint32_t intern(const Node& n)
	{
		if ((nodes_.size() - 256) * 10 >= table_.size() * 6) grow();
		const size_t mask = table_.size() - 1;
		size_t slot = hashNode(n) & mask;
		for (;;) {
			const uint32_t idx = table_[slot];
			if (idx == 0) break;
			if (nodes_[idx] == n) return idx;
			slot = (slot + 1) & mask;
		}
		nodes_.push_back(n);
		table_[slot] = (uint32_t)(nodes_.size() - 1);
		return table_[slot];
	}

	std::vector<Node>& nodes() { return nodes_; }

private:
	static uint64_t hashNode(const Node& n)
	{
		uint64_t h = 0x2545f4914f6cdd1dULL;
		for (int i = 0; i < 8; i += 2) h = mix64(h ^ ((uint64_t)n[i] | ((uint64_t)n[i + 1] << 32)));
		return h;
	}

	void grow()
	{
		std::vector<uint32_t> bigger(table_.size() * 4, 0);
		const size_t mask = bigger.size() - 1;
		for (size_t i = 256; i < nodes_.size(); i++) {
			size_t slot = hashNode(nodes_[i]) & mask;
			while (bigger[slot] != 0) slot = (slot + 1) & mask;
			bigger[slot] = (uint32_t)i;
		}
		table_.swap(bigger);
	}

	std::vector<Node> nodes_;
	std::vector<uint32_t> table_;
};

// ------------------------------------------------------------------ scenes

struct Scene {
	virtual ~Scene() {}
	// Material of one voxel.
	virtual uint8_t voxel(int64_t x, int64_t y, int64_t z) const = 0;
	// Material id if the aligned cube [p, p + 2^h) is certainly uniform, else -1. Conservative.
	virtual int classify(int64_t x, int64_t y, int64_t z, int h) const = 0;
	virtual void colours(float* rgb768) const = 0;
	int64_t lo[3], hi[3]; // inclusive occupied bounding box (conservative)

	bool outside(int64_t x, int64_t y, int64_t z, int h) const
	{
		const int64_t s = INT64_C(1) << h;
		return x > hi[0] || y > hi[1] || z > hi[2] || x + s <= lo[0] || y + s <= lo[1] || z + s <= lo[2];
	}
};

void defaultColours(float* c)
{
	// Purple for unset entries, like Viewer (reference viewer.cpp:50-55).
	for (int i = 0; i < 256; i++) { c[3 * i] = 1.0f; c[3 * i + 1] = 0.0f; c[3 * i + 2] = 1.0f; }
}

void setColour(float* c, int i, float r, float g, float b) { c[3 * i] = r; c[3 * i + 1] = g; c[3 * i + 2] = b; }

// C1: "sphere + noise", N^3 centred on the origin. Solid iff |p|^2 < (R + bump(p >> 2))^2 with
// R = 0.42 N and bump in [0, 0.04 N); material 1 + hash(p >> 4) % 6.
struct SphereNoise : Scene {
	int64_t N, R, B; uint64_t seed;
	SphereNoise(int sizeLog2, uint64_t s) : N(INT64_C(1) << sizeLog2), seed(s)
	{
		R = (N * 42) / 100; B = std::max<int64_t>(1, (N * 4) / 100);
		for (int a = 0; a < 3; a++) { lo[a] = -(R + B) - 1; hi[a] = R + B + 1; }
	}
	int64_t bump(int64_t x, int64_t y, int64_t z) const { return (int64_t)(hash3(x >> 2, y >> 2, z >> 2, seed) & 0xff) * B / 256; }
	uint8_t material(int64_t x, int64_t y, int64_t z) const { return (uint8_t)(1 + hash3(x >> 4, y >> 4, z >> 4, seed ^ 0x5151) % 6); }
	uint8_t voxel(int64_t x, int64_t y, int64_t z) const override
	{
		const int64_t r = R + bump(x, y, z);
		return (x * x + y * y + z * z < r * r) ? material(x, y, z) : 0;
	}
	int classify(int64_t x, int64_t y, int64_t z, int h) const override
	{
		if (outside(x, y, z, h)) return 0;
		const int64_t s = INT64_C(1) << h;
		int64_t nearest2 = 0, farthest2 = 0;
		const int64_t p[3] = { x, y, z };
		for (int a = 0; a < 3; a++) {
			const int64_t l = p[a], u = p[a] + s - 1;
			const int64_t n = (l > 0) ? l : (u < 0 ? -u : 0);
			const int64_t f = std::max(l < 0 ? -l : l, u < 0 ? -u : u);
			nearest2 += n * n; farthest2 += f * f;
		}
		if (nearest2 >= (R + B) * (R + B)) return 0;
		if (farthest2 < R * R && h <= 4) return material(x, y, z);
		return -1;
	}
	void colours(float* c) const override
	{
		defaultColours(c);
		setColour(c, 1, 0.80f, 0.25f, 0.20f); setColour(c, 2, 0.25f, 0.70f, 0.30f); setColour(c, 3, 0.25f, 0.35f, 0.85f);
		setColour(c, 4, 0.85f, 0.80f, 0.30f); setColour(c, 5, 0.70f, 0.70f, 0.70f); setColour(c, 6, 0.55f, 0.30f, 0.65f);
	}
};

// C2/C3: multi-material terrain, N^3 centred on the origin. z is up (the reference viewer's
// convention, camera.cpp:45-48). Height field = integer-lattice value-noise fBm; materials by
// depth below the surface, altitude, a coarse biome cell and warped rock strata.
struct Terrain : Scene {
	int sizeLog2; int64_t N, half; uint64_t seed;
	std::vector<std::vector<int32_t>> mipMin, mipMax; // [level][ (y>>level) * (N>>level) + (x>>level) ]
	enum { SoilDepth = 4, BandLog2 = 5, WarpCellLog2 = 7, BiomeCellLog2 = 8 };

	Terrain(int sl, uint64_t s) : sizeLog2(sl), N(INT64_C(1) << sl), half(N / 2), seed(s)
	{
		buildHeights();
		lo[0] = lo[1] = -half; hi[0] = hi[1] = half - 1;
		lo[2] = -half; hi[2] = mipMax.back()[0];
	}

	// Value noise on an integer lattice of spacing 2^cellLog2, smoothstep-interpolated in 16.16.
	int64_t octave(int64_t x, int64_t y, int cellLog2, int o) const
	{
		const int64_t cx = x >> cellLog2, cy = y >> cellLog2;
		const int64_t mask = (INT64_C(1) << cellLog2) - 1;
		int64_t fx = ((x & mask) << 16) >> cellLog2, fy = ((y & mask) << 16) >> cellLog2; // 0..65535
		fx = (fx * fx >> 16) * (3 * 65536 - 2 * fx) >> 16;
		fy = (fy * fy >> 16) * (3 * 65536 - 2 * fy) >> 16;
		const int64_t v00 = hash3(cx, cy, o, seed) & 0xffff, v10 = hash3(cx + 1, cy, o, seed) & 0xffff;
		const int64_t v01 = hash3(cx, cy + 1, o, seed) & 0xffff, v11 = hash3(cx + 1, cy + 1, o, seed) & 0xffff;
		const int64_t a = v00 + ((v10 - v00) * fx >> 16), b = v01 + ((v11 - v01) * fx >> 16);
		return a + ((b - a) * fy >> 16); // 0..65535
	}

	int32_t heightAt(int64_t x, int64_t y) const
	{
		// Octave o has lattice spacing (N/4) >> o and amplitude (N/5) * (7/16)^o; stop at spacing 8.
		int64_t h = 0;
		int64_t amp = N / 5;
		for (int o = 0, cell = sizeLog2 - 2; cell >= 3 && amp > 0; o++, cell--, amp = amp * 7 / 16) {
			h += (octave(x, y, cell, o) - 32768) * amp >> 15;
		}
		return (int32_t)(h - N / 16);
	}

	void buildHeights()
	{
		mipMin.resize(sizeLog2 + 1); mipMax.resize(sizeLog2 + 1);
		mipMin[0].resize((size_t)N * N);
		const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
		std::vector<std::thread> pool;
		for (unsigned t = 0; t < hw; t++) {
			pool.emplace_back([this, t, hw]() {
				for (int64_t row = t; row < N; row += hw)
					for (int64_t col = 0; col < N; col++)
						mipMin[0][(size_t)row * N + col] = heightAt(col - half, row - half);
			});
		}
		for (auto& th : pool) th.join();
		mipMax[0] = mipMin[0];
		for (int l = 1; l <= sizeLog2; l++) {
			const int64_t w = N >> l, pw = N >> (l - 1);
			mipMin[l].resize((size_t)w * w); mipMax[l].resize((size_t)w * w);
			for (int64_t r = 0; r < w; r++) for (int64_t c = 0; c < w; c++) {
				const size_t i00 = (size_t)(2 * r) * pw + 2 * c, i10 = i00 + 1, i01 = i00 + pw, i11 = i01 + 1;
				mipMin[l][(size_t)r * w + c] = std::min(std::min(mipMin[l-1][i00], mipMin[l-1][i10]), std::min(mipMin[l-1][i01], mipMin[l-1][i11]));
				mipMax[l][(size_t)r * w + c] = std::max(std::max(mipMax[l-1][i00], mipMax[l-1][i10]), std::max(mipMax[l-1][i01], mipMax[l-1][i11]));
			}
		}
	}

	int32_t surface(int64_t x, int64_t y) const { return mipMin[0][(size_t)(y + half) * N + (x + half)]; }
	int warp(int64_t x, int64_t y) const { return (int)(hash3(x >> WarpCellLog2, y >> WarpCellLog2, 77, seed) & 31); }
	int biome(int64_t x, int64_t y) const { return (int)(hash3(x >> BiomeCellLog2, y >> BiomeCellLog2, 99, seed) & 3); }
	uint8_t rock(int64_t x, int64_t y, int64_t z) const { return (uint8_t)(8 + (((z + warp(x, y)) >> BandLog2) & 7)); }

	uint8_t voxel(int64_t x, int64_t y, int64_t z) const override
	{
		if (x < -half || y < -half || z < -half || x >= half || y >= half || z >= half) return 0;
		const int64_t s = surface(x, y);
		if (z >= s) return 0;
		const int64_t depth = s - z;          // 1 = the top voxel of the column
		if (depth > SoilDepth) return rock(x, y, z);
		const int64_t snowLine = N / 10, sandLine = -N / 7;
		if (s > snowLine) return (uint8_t)(depth <= 2 ? 16 : 17);             // snow over scree
		if (s < sandLine) return (uint8_t)(depth <= 3 ? 18 : 19);             // sand over clay
		if (depth == 1) return (uint8_t)(1 + biome(x, y));                    // four kinds of turf
		return (uint8_t)(5 + (biome(x, y) & 1));                              // two kinds of soil
	}

	int classify(int64_t x, int64_t y, int64_t z, int h) const override
	{
		if (outside(x, y, z, h)) return 0;
		if (h >= sizeLog2) return -1;   // aligned cubes this big straddle the scene box
		const int64_t s = INT64_C(1) << h;
		const int64_t w = N >> h;
		const size_t idx = (size_t)((y + half) >> h) * w + (size_t)((x + half) >> h);
		const int64_t hmin = mipMin[h][idx], hmax = mipMax[h][idx];
		if (z >= hmax) return 0;
		if (z + s - 1 < hmin - SoilDepth) {
			// Solid rock throughout: uniform iff one warp cell and one stratum.
			if (h <= WarpCellLog2) {
				const int wv = warp(x, y);
				if (((z + wv) >> BandLog2) == ((z + s - 1 + wv) >> BandLog2)) return rock(x, y, z);
			}
		}
		return -1;
	}

	void colours(float* c) const override
	{
		defaultColours(c);
		setColour(c, 1, 0.33f, 0.55f, 0.20f); setColour(c, 2, 0.40f, 0.60f, 0.22f); setColour(c, 3, 0.28f, 0.48f, 0.18f); setColour(c, 4, 0.45f, 0.58f, 0.25f);
		setColour(c, 5, 0.45f, 0.32f, 0.20f); setColour(c, 6, 0.40f, 0.28f, 0.18f);
		for (int i = 0; i < 8; i++) { const float g = 0.35f + 0.05f * (float)i; setColour(c, 8 + i, g, g * 0.95f, g * 0.9f); }
		setColour(c, 16, 0.95f, 0.95f, 0.98f); setColour(c, 17, 0.55f, 0.55f, 0.58f);
		setColour(c, 18, 0.85f, 0.78f, 0.55f); setColour(c, 19, 0.60f, 0.45f, 0.35f);
	}
};

// C4 stand-in: closed solids (balls and boxes, the shapes a watertight triangle soup of
// icospheres and cuboids voxelises to under the winding-number rule of reference
// src/library/voxelization.cpp:692-744) scattered in an N^3 cube around the origin.
struct Soup : Scene {
	struct Prim { int kind; int64_t c[3]; int64_t r[3]; uint8_t mat; };
	int64_t N; uint64_t seed; std::vector<Prim> prims;
	// Coarse grid of primitive lists so voxel() / classify() stay cheap.
	int gridLog2; int64_t cell; std::vector<std::vector<uint32_t>> grid;

	Soup(int sizeLog2, uint64_t s, int count) : N(INT64_C(1) << sizeLog2), seed(s)
	{
		const int64_t half = N / 2;
		for (int a = 0; a < 3; a++) { lo[a] = -half; hi[a] = half - 1; }
		// Ground slab so that bounce and shadow rays have something to land on.
		Prim ground = { 1, { 0, 0, -half + N / 32 }, { half, half, N / 32 }, 1 };
		prims.push_back(ground);
		for (int i = 0; i < count; i++) {
			Prim p;
			p.kind = (int)(hash3(i, 0, 0, seed) & 1);
			for (int a = 0; a < 3; a++) {
				p.c[a] = (int64_t)(hash3(i, 1, a, seed) % (uint32_t)(N * 3 / 4)) - N * 3 / 8;
				p.r[a] = N / 64 + (int64_t)(hash3(i, 2, a, seed) % (uint32_t)(N / 12));
			}
			p.c[2] = p.c[2] / 2 - N / 8;
			if (p.kind == 0) p.r[1] = p.r[2] = p.r[0];
			p.mat = (uint8_t)(2 + hash3(i, 3, 0, seed) % 20);
			prims.push_back(p);
		}
		gridLog2 = std::max(0, sizeLog2 - 4); cell = INT64_C(1) << gridLog2;
		const int64_t g = N >> gridLog2;
		grid.resize((size_t)(g * g * g));
		for (uint32_t i = 0; i < prims.size(); i++) {
			const Prim& p = prims[i];
			int64_t a0[3], a1[3];
			for (int a = 0; a < 3; a++) {
				a0[a] = std::max<int64_t>(0, (p.c[a] - p.r[a] + half) >> gridLog2);
				a1[a] = std::min<int64_t>(g - 1, (p.c[a] + p.r[a] + half) >> gridLog2);
			}
			for (int64_t z = a0[2]; z <= a1[2]; z++) for (int64_t y = a0[1]; y <= a1[1]; y++) for (int64_t x = a0[0]; x <= a1[0]; x++)
				grid[(size_t)((z * g + y) * g + x)].push_back(i);
		}
	}

	static bool inside(const Prim& p, int64_t x, int64_t y, int64_t z)
	{
		const int64_t dx = x - p.c[0], dy = y - p.c[1], dz = z - p.c[2];
		if (p.kind == 1) return std::llabs(dx) <= p.r[0] && std::llabs(dy) <= p.r[1] && std::llabs(dz) <= p.r[2];
		return dx * dx + dy * dy + dz * dz <= p.r[0] * p.r[0];
	}

	// 0 = cube certainly outside p, 1 = certainly inside, 2 = straddles (conservative).
	static int cubeVs(const Prim& p, int64_t x, int64_t y, int64_t z, int64_t s)
	{
		const int64_t q[3] = { x, y, z };
		if (p.kind == 1) {
			bool in = true;
			for (int a = 0; a < 3; a++) {
				const int64_t l = q[a], u = q[a] + s - 1;
				if (u < p.c[a] - p.r[a] || l > p.c[a] + p.r[a]) return 0;
				if (l < p.c[a] - p.r[a] || u > p.c[a] + p.r[a]) in = false;
			}
			return in ? 1 : 2;
		}
		int64_t n2 = 0, f2 = 0;
		for (int a = 0; a < 3; a++) {
			const int64_t l = q[a] - p.c[a], u = q[a] + s - 1 - p.c[a];
			const int64_t n = (l > 0) ? l : (u < 0 ? -u : 0);
			const int64_t f = std::max(std::llabs(l), std::llabs(u));
			n2 += n * n; f2 += f * f;
		}
		if (n2 > p.r[0] * p.r[0]) return 0;
		return (f2 <= p.r[0] * p.r[0]) ? 1 : 2;
	}

	const std::vector<uint32_t>& cellList(int64_t x, int64_t y, int64_t z) const
	{
		const int64_t half = N / 2, g = N >> gridLog2;
		return grid[(size_t)((((z + half) >> gridLog2) * g + ((y + half) >> gridLog2)) * g + ((x + half) >> gridLog2))];
	}

	// First primitive in list order wins (a fixed, order-defined union).
	uint8_t voxel(int64_t x, int64_t y, int64_t z) const override
	{
		const int64_t half = N / 2;
		if (x < -half || y < -half || z < -half || x >= half || y >= half || z >= half) return 0;
		for (uint32_t i : cellList(x, y, z)) if (inside(prims[i], x, y, z)) return prims[i].mat;
		return 0;
	}

	int classify(int64_t x, int64_t y, int64_t z, int h) const override
	{
		if (outside(x, y, z, h)) return 0;
		if (h > gridLog2) return -1;
		const int64_t s = INT64_C(1) << h;
		for (uint32_t i : cellList(x, y, z)) {
			const int r = cubeVs(prims[i], x, y, z, s);
			if (r == 1) return prims[i].mat;   // earlier primitives were all certainly outside
			if (r == 2) return -1;
		}
		return 0;
	}

	void colours(float* c) const override
	{
		defaultColours(c);
		setColour(c, 1, 0.55f, 0.55f, 0.50f);
		for (int i = 2; i < 22; i++) {
			const uint32_t hsh = hash3(i, 9, 9, 1234);
			setColour(c, i, 0.25f + 0.7f * (float)(hsh & 0xff) / 255.0f, 0.25f + 0.7f * (float)((hsh >> 8) & 0xff) / 255.0f,
				0.25f + 0.7f * (float)((hsh >> 16) & 0xff) / 255.0f);
		}
	}
};

// C5: "city" -- a grid of lots, each holding one of a few dozen procedural building types, so
// the DAG shares whole buildings by construction. N^3 centred on the origin, ground at z = 0.
struct City : Scene {
	int sizeLog2; int64_t N, half; uint64_t seed;
	enum { LotLog2 = 8, Types = 48 };  // 256-voxel lots
	City(int sl, uint64_t s) : sizeLog2(sl), N(INT64_C(1) << sl), half(N / 2), seed(s)
	{
		lo[0] = lo[1] = -half; hi[0] = hi[1] = half - 1;
		lo[2] = -64; hi[2] = maxHeight();
	}
	static int64_t maxHeight() { return 64 + 8 * 120; }
	int lotType(int64_t x, int64_t y) const { return (int)(hash3(x >> LotLog2, y >> LotLog2, 5, seed) % Types); }

	// Building of type t in lot-local coordinates u, v in [0, 256), w = z.
	uint8_t building(int t, int64_t u, int64_t v, int64_t w) const
	{
		const uint32_t g = hash3(t, 0, 0, seed ^ 0xc17);
		const int64_t margin = 24 + (int64_t)(g & 31);                    // street + pavement
		const int64_t floors = 8 + (int64_t)((g >> 5) % 113);
		const int64_t top = floors * 8;
		if (w < 0) return 0;
		if (u < 16 || v < 16 || u >= 240 || v >= 240) return (w < 2) ? 20 : 0;      // road
		if (u < margin || v < margin || u >= 256 - margin || v >= 256 - margin) return (w < 4) ? 21 : 0; // pavement
		if (w >= top) {
			// roof furniture: a small plant room
			const int64_t c = 128, r = 10 + (int64_t)((g >> 12) & 15);
			return (w < top + 12 && std::llabs(u - c) < r && std::llabs(v - c) < r) ? 24 : 0;
		}
		// setbacks: every 32 floors the tower steps in by 8
		const int64_t inset = margin + 8 * (w / 256);
		if (u < inset || v < inset || u >= 256 - inset || v >= 256 - inset) return 0;
		const bool shell = (u < inset + 2 || v < inset + 2 || u >= 254 - inset || v >= 254 - inset);
		if (!shell) return 23;                                              // solid core
		const bool window = ((w & 7) >= 2 && (w & 7) <= 5) && ((((u + v) >> 2) & 3) != 0);
		return window ? (uint8_t)(25 + ((g >> 20) & 3)) : (uint8_t)(29 + ((g >> 24) % 12));
	}

	uint8_t voxel(int64_t x, int64_t y, int64_t z) const override
	{
		if (x < -half || y < -half || x >= half || y >= half) return 0;
		if (z < 0) return (z >= -64) ? 22 : 0;                                 // bedrock slab
		if (z > maxHeight()) return 0;
		return building(lotType(x, y), x & 255, y & 255, z);
	}

	int classify(int64_t x, int64_t y, int64_t z, int h) const override
	{
		if (outside(x, y, z, h)) return 0;
		const int64_t s = INT64_C(1) << h;
		if (x >= -half && y >= -half && x + s <= half && y + s <= half && z >= -64 && z + s <= 0) return 22;
		if (h <= LotLog2 && z >= 0 && x >= -half && y >= -half && x < half && y < half) {
			// The cube lies inside one lot and one 256-voxel layer (aligned, s <= 256).
			const int64_t u0 = x & 255, v0 = y & 255, u1 = u0 + s - 1, v1 = v0 + s - 1, w0 = z, w1 = z + s - 1;
			const uint32_t g = hash3(lotType(x, y), 0, 0, seed ^ 0xc17);
			const int64_t margin = 24 + (int64_t)(g & 31);
			const int64_t top = (8 + (int64_t)((g >> 5) % 113)) * 8;
			if (w0 >= top + 12) return 0;
			const int64_t inset = margin + 8 * (w0 / 256);
			if (w0 >= 4 && (u1 < inset || v1 < inset || u0 >= 256 - inset || v0 >= 256 - inset)) return 0;
			if (w1 < top && u0 >= inset + 2 && v0 >= inset + 2 && u1 < 254 - inset && v1 < 254 - inset) return 23;
		}
		return -1;
	}

	void colours(float* c) const override
	{
		defaultColours(c);
		setColour(c, 20, 0.18f, 0.18f, 0.20f); setColour(c, 21, 0.55f, 0.55f, 0.55f); setColour(c, 22, 0.30f, 0.27f, 0.25f);
		setColour(c, 23, 0.70f, 0.68f, 0.62f); setColour(c, 24, 0.45f, 0.45f, 0.48f);
		for (int i = 0; i < 4; i++) setColour(c, 25 + i, 0.35f + 0.1f * (float)i, 0.55f, 0.75f);
		for (int i = 0; i < 12; i++) { const float g = 0.4f + 0.04f * (float)i; setColour(c, 29 + i, g, g * 0.92f, g * 0.85f); }
	}
};

// ------------------------------------------------------------------ builder

// Three passes so that big scenes build on all host cores yet come out bit-identical whatever the
// thread count: (1) walk the top of the tree single-threaded and list the mixed sub-cubes at
// `splitHeight` as tasks; (2) build each task into a private pool in parallel; (3) walk the top
// again, folding the private pools into the global one in task order.
struct Builder {
	const Scene& scene;
	NodePool pool;
	int splitHeight;
	struct Task { int64_t x, y, z; int h; uint32_t result; std::unique_ptr<NodePool> local; };
	std::vector<Task> tasks;
	size_t cursor = 0;
	int pass = 0;

	Builder(const Scene& s, int split) : scene(s), splitHeight(split) {}
	virtual ~Builder() {}

	static uint32_t buildInto(const Scene& sc, NodePool& np, int64_t x, int64_t y, int64_t z, int h)
	{
		if (h == 0) return sc.voxel(x, y, z);
		const int c = sc.classify(x, y, z, h);
		if (c >= 0) return (uint32_t)c;
		Node n;
		const int64_t half = INT64_C(1) << (h - 1);
		bool uniform = true;
		for (int i = 0; i < 8; i++) {
			n[i] = buildInto(sc, np, x + ((i & 1) ? half : 0), y + ((i & 2) ? half : 0), z + ((i & 4) ? half : 0), h - 1);
			if (n[i] != n[0]) uniform = false;
		}
		if (uniform && n[0] < 256) return n[0];
		return np.intern(n);
	}

	// Identity of a sub-cube for sharing whole sub-trees between tasks (city lots); 0 = none.
	virtual uint64_t shareKey(int64_t, int64_t, int64_t, int) const { return 0; }

	uint32_t leafTask(int64_t x, int64_t y, int64_t z, int h)
	{
		if (pass == 1) {
			Task t; t.x = x; t.y = y; t.z = z; t.h = h; t.result = 0;
			tasks.push_back(std::move(t));
			return 256; // placeholder: "some internal node"
		}
		Task& t = tasks[cursor++];
		if (!t.local) return t.result; // shared with an earlier task, already folded in
		// Fold the private pool into the global one; children precede parents in creation order.
		std::vector<Node>& ln = t.local->nodes();
		std::vector<uint32_t> remap(ln.size());
		for (uint32_t i = 0; i < 256; i++) remap[i] = i;
		for (size_t i = 256; i < ln.size(); i++) {
			Node n = ln[i];
			for (int c = 0; c < 8; c++) n[c] = remap[n[c]];
			remap[i] = pool.intern(n);
		}
		t.result = remap[t.result];
		t.local.reset();
		return t.result;
	}

	uint32_t top(int64_t x, int64_t y, int64_t z, int h)
	{
		if (h == 0) return scene.voxel(x, y, z);
		const int c = scene.classify(x, y, z, h);
		if (c >= 0) return (uint32_t)c;
		if (h <= splitHeight) return leafTask(x, y, z, h);
		Node n;
		const int64_t half = INT64_C(1) << (h - 1);
		bool uniform = true;
		for (int i = 0; i < 8; i++) {
			n[i] = top(x + ((i & 1) ? half : 0), y + ((i & 2) ? half : 0), z + ((i & 4) ? half : 0), h - 1);
			if (n[i] != n[0]) uniform = false;
		}
		if (uniform && n[0] < 256) return n[0];
		return (pass == 1) ? 256u : pool.intern(n);
	}

	uint32_t run()
	{
		const int64_t lowest = -(INT64_C(1) << 31);
		pass = 1;
		top(lowest, lowest, lowest, 32);

		// Tasks with the same share key are built once.
		std::vector<size_t> owner(tasks.size());
		{
			std::vector<std::pair<uint64_t, size_t>> seen;
			for (size_t i = 0; i < tasks.size(); i++) {
				owner[i] = i;
				const uint64_t k = shareKey(tasks[i].x, tasks[i].y, tasks[i].z, tasks[i].h);
				if (k == 0) continue;
				bool found = false;
				for (auto& e : seen) if (e.first == k) { owner[i] = e.second; found = true; break; }
				if (!found) seen.push_back(std::make_pair(k, i));
			}
		}
		std::atomic<size_t> next(0);
		const unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
		std::vector<std::thread> workers;
		for (unsigned w = 0; w < hw; w++) {
			workers.emplace_back([&]() {
				for (;;) {
					const size_t i = next.fetch_add(1);
					if (i >= tasks.size()) break;
					if (owner[i] != i) continue;
					Task& t = tasks[i];
					t.local.reset(new NodePool());
					t.result = buildInto(scene, *t.local, t.x, t.y, t.z, t.h);
				}
			});
		}
		for (auto& th : workers) th.join();

		pass = 2;
		cursor = 0;
		// An owner is always earlier in task order than its sharers, so it is folded in first.
		return topResolve(owner);
	}

	uint32_t topResolve(const std::vector<size_t>& owner)
	{
		owner_ = &owner;
		const int64_t lowest = -(INT64_C(1) << 31);
		return topPass2(lowest, lowest, lowest, 32);
	}

	const std::vector<size_t>* owner_ = nullptr;

	uint32_t topPass2(int64_t x, int64_t y, int64_t z, int h)
	{
		if (h == 0) return scene.voxel(x, y, z);
		const int c = scene.classify(x, y, z, h);
		if (c >= 0) return (uint32_t)c;
		if (h <= splitHeight) {
			const size_t i = cursor;
			const size_t o = (*owner_)[i];
			if (o != i) { cursor++; return tasks[o].result; } // owner already folded (earlier in order)
			return leafTask(x, y, z, h);
		}
		Node n;
		const int64_t half = INT64_C(1) << (h - 1);
		bool uniform = true;
		for (int i = 0; i < 8; i++) {
			n[i] = topPass2(x + ((i & 1) ? half : 0), y + ((i & 2) ? half : 0), z + ((i & 4) ? half : 0), h - 1);
			if (n[i] != n[0]) uniform = false;
		}
		if (uniform && n[0] < 256) return n[0];
		return pool.intern(n);
	}
};

struct CityBuilder : Builder {
	const City& city;
	explicit CityBuilder(const City& c) : Builder(c, City::LotLog2), city(c) {}
	// A lot-aligned 256^3 cube above ground depends only on (building type, layer).
	uint64_t shareKey(int64_t x, int64_t y, int64_t z, int h) const override
	{
		if (h != City::LotLog2 || z < 0) return 0;
		return UINT64_C(1) + (uint64_t)city.lotType(x, y) * 4096 + (uint64_t)(z >> City::LotLog2);
	}
};

} // namespace

struct cbq_scene {
	std::unique_ptr<Scene> scene;
	std::vector<Node> nodes;
	uint32_t root = 0;
};

extern "C" {

int cbq_scene_build(const char* kind, uint32_t size_log2, uint64_t seed, cbq_scene** out)
{
	if (!kind || !out || size_log2 < 3 || size_log2 > 31) return CBQ_ERROR_INVALID_ARGUMENT;
	std::unique_ptr<cbq_scene> s(new cbq_scene());
	const std::string k(kind);
	try {
		if (k == "sphere_noise") s->scene.reset(new SphereNoise((int)size_log2, seed));
		else if (k == "terrain") { if (size_log2 > 14) return CBQ_ERROR_INVALID_ARGUMENT; s->scene.reset(new Terrain((int)size_log2, seed)); }
		else if (k == "soup") s->scene.reset(new Soup((int)size_log2, seed, 160));
		else if (k == "city") { if (size_log2 < 9) return CBQ_ERROR_INVALID_ARGUMENT; s->scene.reset(new City((int)size_log2, seed)); }
		else return CBQ_ERROR_INVALID_ARGUMENT;

		if (k == "city") {
			CityBuilder b(static_cast<const City&>(*s->scene));
			s->root = b.run();
			s->nodes.swap(b.pool.nodes());
		} else {
			// Split four levels below the scene cube: up to 4096 tasks, enough to balance any host.
			Builder b(*s->scene, std::max(3, (int)size_log2 - 4));
			s->root = b.run();
			s->nodes.swap(b.pool.nodes());
		}
	} catch (const std::bad_alloc&) {
		return CBQ_ERROR_OUT_OF_MEMORY;
	}
	*out = s.release();
	return CBQ_OK;
}

const uint32_t* cbq_scene_nodes(const cbq_scene* s, uint64_t* node_count)
{
	if (node_count) *node_count = s->nodes.size();
	return reinterpret_cast<const uint32_t*>(s->nodes.data());
}

uint32_t cbq_scene_root(const cbq_scene* s) { return s->root; }

void cbq_scene_bounds(const cbq_scene* s, int32_t lower[3], int32_t upper[3])
{
	for (int a = 0; a < 3; a++) { lower[a] = (int32_t)s->scene->lo[a]; upper[a] = (int32_t)s->scene->hi[a]; }
}

void cbq_scene_colours(const cbq_scene* s, float* rgb768) { s->scene->colours(rgb768); }

void cbq_scene_voxels(const cbq_scene* s, const int32_t* xyz, uint64_t n, uint8_t* out)
{
	for (uint64_t i = 0; i < n; i++) out[i] = s->scene->voxel(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

void cbq_scene_free(cbq_scene* s) { delete s; }

} // extern "C"

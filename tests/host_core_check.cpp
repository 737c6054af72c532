// TEST-ONLY: compiles the device traversal core (cubiquity_b200/csrc/traverse.cuh) for the HOST so
// that its per-ray state machine can be checked against the oracle on a machine without a GPU.
// This object is built into tests/_build/ by tests/conftest.py; it is never part of the product
// library and the product has no CPU path.
#include "../cubiquity_b200/csrc/traverse.cuh"

#include <cstring>

namespace {
struct HostNodes {
	const uint32_t* base;
	uint32_t child(uint32_t node, uint32_t slot) const { return base[(size_t)node * 8 + slot]; }
	void prefetch(uint32_t) const {}
};
struct HostStack {
	uint32_t v[33];
	void store(int h, uint32_t n) { v[h] = n; }
	uint32_t load(int h) const { return v[h]; }
	void clear(int top) { for (int h = 0; h <= top; h++) v[h] = 0; }
};
}

extern "C" void host_core_trace(const uint32_t* nodes, const void* subdags, const void* rays, uint64_t n,
	int surface, float maxFootprint, void* hits)
{
	static_assert(sizeof(cbq::Hit) == 40 && sizeof(cbq::Ray) == 24 && sizeof(cbq::SubDag) == 32, "layouts");
	const cbq::Ray* r = static_cast<const cbq::Ray*>(rays);
	cbq::Hit* h = static_cast<cbq::Hit*>(hits);
	const cbq::SubDag* sd = static_cast<const cbq::SubDag*>(subdags);
	HostNodes hn{ nodes };
	const bool lodOff = (maxFootprint == -1.0f);
	for (uint64_t i = 0; i < n; i++) {
		HostStack st;
		std::memset(&st, 0xee, sizeof(st));    // stale garbage, like a shared-memory column another ray has used
		if (lodOff) cbq::traceRay<true>(r[i], hn, sd, st, maxFootprint, surface != 0, h[i]);
		else cbq::traceRay<false>(r[i], hn, sd, st, maxFootprint, surface != 0, h[i]);
	}
}

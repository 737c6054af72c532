// TEST-ONLY: compiles the device traversal core (cubiquity_b200/csrc/traverse.cuh) for the HOST so
// that its per-ray state machine can be checked against the oracle on a machine without a GPU.
// This object is built into tests/_build/ by tests/conftest.py; it is never part of the product
// library and the product has no CPU path.
#include "../cubiquity_b200/csrc/traverse.cuh"

#include <cstring>
#include <vector>

namespace {
template <typename Ref>
struct HostNodes {
	const Ref* base;
	Ref child(Ref node, uint32_t slot) const { return base[(size_t)cbq::refIndex(node) * 8 + slot]; }
};
template <typename Ref>
struct HostStack {
	Ref v[33];
	void store(int h, Ref n) { v[h] = n; }
	Ref load(int h) const { return v[h]; }
	void clear(int top) { for (int h = 0; h <= top; h++) v[h] = 0; }
};

template <typename Ref>
void run(const uint32_t* nodes, uint64_t nodeCount, const cbq::SubDag* sd, const cbq::Ray* r, uint64_t n, int surface, float maxFootprint, cbq::Hit* h)
{
	// The transcoding the device does in packNodes (csrc/pack_kernels.cu), with the same functions.
	std::vector<Ref> packed(nodeCount * 8);
	for (uint64_t i = 0; i < nodeCount * 8; i++) packed[i] = cbq::childRef<Ref>(nodes, nodes[i]);
	Ref roots[8];
	for (int i = 0; i < 8; i++) roots[i] = cbq::nodeRef<Ref>(nodes, sd[i].node);
	HostNodes<Ref> hn{ packed.data() };
	const bool lodOff = (maxFootprint == -1.0f);
	for (uint64_t i = 0; i < n; i++) {
		HostStack<Ref> st;
		std::memset(&st, 0xee, sizeof(st));    // stale garbage, like a shared-memory column another ray has used
		if (lodOff) cbq::traceRay<true>(r[i], hn, sd, roots, st, maxFootprint, surface != 0, h[i]);
		else cbq::traceRay<false>(r[i], hn, sd, roots, st, maxFootprint, surface != 0, h[i]);
	}
}
}

// ref_bits: 32 or 64, the two widths of packed reference the kernels are instantiated for.
extern "C" void host_core_trace(const uint32_t* nodes, uint64_t nodeCount, const void* subdags, const void* rays, uint64_t n,
	int surface, float maxFootprint, int refBits, void* hits)
{
	static_assert(sizeof(cbq::Hit) == 40 && sizeof(cbq::Ray) == 24 && sizeof(cbq::SubDag) == 32, "layouts");
	const cbq::Ray* r = static_cast<const cbq::Ray*>(rays);
	cbq::Hit* h = static_cast<cbq::Hit*>(hits);
	const cbq::SubDag* sd = static_cast<const cbq::SubDag*>(subdags);
	if (refBits == 32) run<uint32_t>(nodes, nodeCount, sd, r, n, surface, maxFootprint, h);
	else run<uint64_t>(nodes, nodeCount, sd, r, n, surface, maxFootprint, h);
}

// DAG -> GPU serialisation, second half: the array of PACKED REFERENCES the ray cast follows (traverse.cuh),
// derived on the device from the reference-layout node array that cbq_upload / cbq_update / the device-side
// edit and bake leave in HBM.
//
// Reference counterpart: none -- gpu_pathtracing_viewer.cpp:43-67 uploads NodeStore::rawBytesPtr() as it is and the
// GLSL ray cast reads one child word per loop trip, like the CPU code (raytracing.cpp:264-265).
//
// Per node: 32 B read (its 8 child words) + for each internal child the child's own 32 B (one sector, to form its
// occupancy mask) + 8 references written: ~ 32 + 8 * 32 + 32..64 B of sector traffic, bound by HBM/L2 sector rate.
// One thread per (node, slot); a warp covers 4 nodes, so the 8 threads of a node read one sector together.
#include "cbq_internal.h"

namespace cbq {

namespace {

template <typename Ref>
__global__ void __launch_bounds__(256)
packNodes(const uint32_t* __restrict__ nodes, uint64_t begin, uint64_t end, Ref* __restrict__ refs)
{
	const uint64_t words = (end - begin) * 8u;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t w = begin * 8u + i;
		const uint32_t child = nodes[w];
		Ref r = (Ref)child;
		if (child >= kMaterialCount) {
			const uint4* c = reinterpret_cast<const uint4*>(nodes) + (size_t)child * 2;
			const uint4 a = c[0], b = c[1];
			const uint32_t mask = (a.x ? 1u : 0u) | (a.y ? 2u : 0u) | (a.z ? 4u : 0u) | (a.w ? 8u : 0u) |
			                      (b.x ? 16u : 0u) | (b.y ? 32u : 0u) | (b.z ? 64u : 0u) | (b.w ? 128u : 0u);
			r = ((Ref)child << 8) | (Ref)mask;
		}
		refs[w] = r;
	}
}

// References of the 8 sub-DAG roots, always 64-bit (the kernels narrow them).
__global__ void packRoots(const uint32_t* __restrict__ nodes, const SubDag* __restrict__ subdags, unsigned long long* __restrict__ rootRefs)
{
	const uint32_t i = threadIdx.x;
	if (i < 8u) rootRefs[i] = nodeRef<unsigned long long>(nodes, subdags[i].node);
}

} // namespace

cudaError_t launchPackNodes(const uint32_t* nodes, uint64_t begin, uint64_t end, void* refs, int refBits, int smCount, cudaStream_t stream)
{
	if (end <= begin) return cudaSuccess;
	uint64_t blocks = ((end - begin) * 8u + 255u) / 256u;
	const uint64_t cap = (uint64_t)smCount * 16u;
	if (blocks > cap) blocks = cap;
	if (refBits == 32) packNodes<uint32_t><<<(int)blocks, 256, 0, stream>>>(nodes, begin, end, static_cast<uint32_t*>(refs));
	else packNodes<uint64_t><<<(int)blocks, 256, 0, stream>>>(nodes, begin, end, static_cast<uint64_t*>(refs));
	return cudaGetLastError();
}

cudaError_t launchPackRoots(const uint32_t* nodes, const SubDag* subdags, unsigned long long* rootRefs, cudaStream_t stream)
{
	packRoots<<<1, 32, 0, stream>>>(nodes, subdags, rootRefs);
	return cudaGetLastError();
}

} // namespace cbq

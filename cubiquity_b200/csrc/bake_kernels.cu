// Device-side `bake`: hash-cons merge of a node array that is already resident in HBM.
//
// What the reference does on one CPU thread (Volume::bake -> NodeStore::merge / merge_node, reference
// src/library/storage.cpp:208-290, 388-395): walk the DAG depth first from the root; replace every child by its
// merged index; a node whose eight (merged) children are the same material BECOMES that material
// (isMaterialNode(const Node&), storage.cpp:69-75); otherwise look the node up by CONTENT in an open-addressing
// table and append it if absent; finally compact the table and rewrite the child indices. The result is the
// canonical minimal DAG of everything reachable from the root; which slot a node lands in is an artefact of the
// reference's hash function and visiting order and is NOT part of the format (children are absolute indices).
//
// Here the same canonical DAG is produced level-synchronously, bottom-up by rank (a node can be finished once
// all its children are), one thread per node:
//   1. reachKernel, <= 33 passes: mark what the root reaches (the DAG is at most 32 levels deep).
//   2. resolveKernel + insertKernel, <= 33 passes: a node whose children all have final ids writes its canonical
//      content (children replaced by ids) and is hash-consed into a table of OWNER indices with one atomicCAS;
//      keys are compared exactly against the owner's canonical content (no reliance on hash uniqueness).
//      The id of a distinct node is the owner's ORIGINAL index; duplicates record the smallest original index of
//      their class with atomicMin, which makes the final order independent of who won the CAS races.
//   3. flag the representative (smallest original index) of every class, exclusive scan, emit: distinct nodes
//      keep the relative order their first occurrence had in the input, so the output is deterministic and
//      baking a baked array is the identity.
// All of it is integer gather/scatter work bound by HBM sectors (DESIGN.md section 4.7 has the bytes per node).
#include "cbq_internal.h"

#include <vector>

namespace cbq {

namespace {

constexpr uint32_t kUnresolved = 0xffffffffu;
constexpr uint32_t kPending = 0xfffffffeu;
constexpr uint32_t kEmptySlot = 0xffffffffu;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                        // per thread
constexpr uint32_t kScanTile = kScanThreads * kScanItems;

struct NodeWords {
	uint32_t w[8];
};

__device__ __forceinline__ NodeWords loadNode(const uint32_t* nodes, uint32_t i)
{
	const uint4* p = reinterpret_cast<const uint4*>(nodes) + (size_t)i * 2;
	const uint4 a = p[0], b = p[1];
	NodeWords n;
	n.w[0] = a.x; n.w[1] = a.y; n.w[2] = a.z; n.w[3] = a.w;
	n.w[4] = b.x; n.w[5] = b.y; n.w[6] = b.z; n.w[7] = b.w;
	return n;
}

__device__ __forceinline__ void storeNode(uint32_t* nodes, uint32_t i, const NodeWords& n)
{
	uint4* p = reinterpret_cast<uint4*>(nodes) + (size_t)i * 2;
	p[0] = make_uint4(n.w[0], n.w[1], n.w[2], n.w[3]);
	p[1] = make_uint4(n.w[4], n.w[5], n.w[6], n.w[7]);
}

__device__ __forceinline__ uint64_t hashNode(const NodeWords& n)
{
	uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
	for (int k = 0; k < 8; k += 2) {
		h ^= (uint64_t)n.w[k] | ((uint64_t)n.w[k + 1] << 32);
		h *= 0xff51afd7ed558ccdull;
		h ^= h >> 29;
	}
	h *= 0xc4ceb9fe1a85ec53ull;
	h ^= h >> 32;
	return h;
}

// The passes below walk the array in INDEX order, 256 consecutive nodes (one chunk) per block iteration, because
// that is where the locality is: children sit near their siblings, so the 4-byte gathers of a chunk share
// sectors (a version driven by compacted work lists in breadth-first order was 1.7x SLOWER: 9.3 GB instead of
// 4.8 GB of DRAM reads in the resolve passes, profiles/r01_analysis.md). What makes the <= 33 passes cheap is a
// per-chunk activity word: a block fetches the words of its next 256 chunks with one coalesced load and only
// touches chunks that still have work.
constexpr uint32_t kChunk = 256;

// Calls body(chunk, word) for every chunk of this block whose activity word is non-zero. Block-uniform control flow.
template <typename Word, typename F>
__device__ __forceinline__ void forActiveChunks(const Word* activity, uint32_t chunks, F body)
{
	__shared__ uint32_t activeChunk[kChunk], activeWord[kChunk];
	__shared__ unsigned int activeCount;
	for (uint32_t first = blockIdx.x; first < chunks; first += gridDim.x * kChunk) {
		const uint32_t mine = first + threadIdx.x * gridDim.x;
		__syncthreads();                                  // the previous batch is done with the lists
		if (threadIdx.x == 0) activeCount = 0;
		__syncthreads();
		const uint32_t word = mine < chunks ? (uint32_t)activity[mine] : 0u;
		if (word != 0) {                                  // order within the batch does not matter
			const unsigned int at = atomicAdd(&activeCount, 1u);
			activeChunk[at] = mine;
			activeWord[at] = word;
		}
		__syncthreads();
		const unsigned int count = activeCount;
		for (unsigned int k = 0; k < count; k++) body(activeChunk[k], activeWord[k]);
	}
}

__global__ void bakeInit(uint32_t n, uint32_t root, uint8_t* mark, uint32_t* newIndex, uint32_t* rep, uint32_t* flags,
	uint8_t* frontierA, uint8_t* frontierB, uint32_t* chunkRemaining, uint32_t* chunkPending, uint32_t chunks)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		mark[i] = (i == root) ? 1 : 0;
		newIndex[i] = i < kMaterialCount ? i : kUnresolved;
		rep[i] = 0xffffffffu;
		flags[i] = 0;
		if (i < chunks) {
			frontierA[i] = 0;
			frontierB[i] = (i == root / kChunk) ? 1 : 0;     // depth 1 (odd) reads B: the root is its frontier
			chunkRemaining[i] = 0;
			chunkPending[i] = 0;
		}
	}
}

// Pass `depth` (1-based): expand the nodes first reached at that depth; `cur` flags the chunks that hold some.
__global__ void bakeReach(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t depth, uint8_t* mark, uint8_t* cur, uint8_t* next, uint32_t chunks)
{
	forActiveChunks(cur, chunks, [&](uint32_t chunk, uint32_t) {
		const uint32_t i = chunk * kChunk + threadIdx.x;
		if (i < n && mark[i] == depth) {
			const NodeWords node = loadNode(nodes, i);
#pragma unroll
			for (int k = 0; k < 8; k++) {
				const uint32_t c = node.w[k];
				if (c < kMaterialCount || c >= n) continue;
				// Parents racing for the same child all store the same bytes.
				if (mark[c] == 0) { mark[c] = (uint8_t)(depth + 1u); next[c / kChunk] = 1; }
			}
		}
		if (threadIdx.x == 0) cur[chunk] = 0;             // clean for its next turn as `next`
	});
}

// chunkRemaining[chunk] = non-material nodes of the chunk the root reaches; counters[0] = results[3] = their total.
__global__ void bakeCountReached(uint32_t n, const uint8_t* __restrict__ mark, uint32_t* chunkRemaining, uint32_t chunks,
	unsigned long long* counters, unsigned long long* results)
{
	for (uint32_t chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
		const uint32_t i = chunk * kChunk + threadIdx.x;
		const int c = __syncthreads_count(i >= kMaterialCount && i < n && mark[i] != 0);
		if (threadIdx.x == 0 && c) {
			chunkRemaining[chunk] = (uint32_t)c;
			atomicAdd(&counters[0], (unsigned long long)c);
			atomicAdd(&results[3], (unsigned long long)c);
		}
	}
}

// A reachable node whose children all have final ids becomes either a material (all eight the same material) or
// a pending hash-cons candidate with its canonical content written out.
__global__ void bakeResolve(const uint32_t* __restrict__ nodes, uint32_t n, const uint8_t* __restrict__ mark,
	uint32_t* newIndex, uint32_t* canon, uint32_t* chunkRemaining, uint32_t* chunkPending, uint32_t chunks, unsigned long long* counters)
{
	forActiveChunks(chunkRemaining, chunks, [&](uint32_t chunk, uint32_t remaining) {
		const uint32_t i = chunk * kChunk + threadIdx.x;
		bool collapsed = false, pending = false;
		if (i < n && mark[i] != 0 && newIndex[i] == kUnresolved) {
			NodeWords node = loadNode(nodes, i);
			bool ready = true;
#pragma unroll
			for (int k = 0; k < 8; k++) {
				uint32_t c = node.w[k];
				if (c >= kMaterialCount) c = (c < n) ? __ldcg(&newIndex[c]) : kUnresolved;
				ready = ready && (c < kPending);
				node.w[k] = c;
			}
			if (ready) {
				bool uniform = node.w[0] < kMaterialCount;
#pragma unroll
				for (int k = 1; k < 8; k++) uniform = uniform && (node.w[k] == node.w[0]);
				if (uniform) {
					newIndex[i] = node.w[0];         // the whole cube is one material (storage.cpp:261-263)
					collapsed = true;
				} else {
					storeNode(canon, i, node);
					newIndex[i] = kPending;
					pending = true;
				}
			}
		}
		const int nCollapsed = __syncthreads_count(collapsed);
		const int nPending = __syncthreads_count(pending);
		if (threadIdx.x == 0) {
			if (nCollapsed) { chunkRemaining[chunk] = remaining - (uint32_t)nCollapsed; atomicAdd(&counters[0], ~(unsigned long long)nCollapsed + 1ull); }
			if (nPending) chunkPending[chunk] = (uint32_t)nPending;
		}
	});
}

// Hash-cons the pending nodes. table[slot] = original index of the node that owns the slot.
__global__ void bakeInsert(uint32_t n, uint32_t* newIndex, const uint32_t* __restrict__ canon, uint32_t* table, uint32_t tableMask,
	uint32_t* rep, uint32_t* chunkRemaining, uint32_t* chunkPending, uint32_t chunks, unsigned long long* counters)
{
	forActiveChunks(chunkPending, chunks, [&](uint32_t chunk, uint32_t pending) {
		const uint32_t i = chunk * kChunk + threadIdx.x;
		bool won = false;
		if (i < n && newIndex[i] == kPending) {
			const NodeWords key = loadNode(canon, i);
			uint32_t slot = (uint32_t)hashNode(key) & tableMask;
			uint32_t owner;
			for (;;) {
				owner = table[slot];
				if (owner == kEmptySlot) {
					owner = atomicCAS(&table[slot], kEmptySlot, i);
					if (owner == kEmptySlot) { owner = i; won = true; break; }
				}
				// canon[owner] was written by bakeResolve, i.e. before this kernel started: plain loads see it.
				const NodeWords other = loadNode(canon, owner);
				bool same = true;
#pragma unroll
				for (int k = 0; k < 8; k++) same = same && (other.w[k] == key.w[k]);
				if (same) break;
				slot = (slot + 1u) & tableMask;
			}
			newIndex[i] = owner;
			atomicMin(&rep[owner], i);
		}
		const int distinct = __syncthreads_count(won);
		if (threadIdx.x == 0) {
			chunkPending[chunk] = 0;
			chunkRemaining[chunk] -= pending;
			atomicAdd(&counters[0], ~(unsigned long long)pending + 1ull);
			if (distinct) atomicAdd(&counters[1], (unsigned long long)distinct);
		}
	});
}

__global__ void bakeFlag(uint32_t n, const uint8_t* __restrict__ mark, const uint32_t* __restrict__ newIndex, const uint32_t* __restrict__ rep, uint32_t* flags)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (i >= kMaterialCount && mark[i] != 0 && newIndex[i] == i) flags[rep[i]] = 1;   // i owns a class; rep[i] <= i is its first occurrence
	}
}

// ---- exclusive scan of 32-bit flags/counts: tile reduce -> scan of tile sums (recursive) -> tile scan ----
__global__ void scanTileSums(const uint32_t* __restrict__ in, uint64_t n, uint32_t* sums)
{
	__shared__ uint32_t warpSums[kScanThreads / 32];
	const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
	uint32_t s = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		const uint64_t i = base + (uint64_t)k * kScanThreads + threadIdx.x;
		if (i < n) s += in[i];
	}
	for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < kScanThreads / 32; w++) t += warpSums[w];
		sums[blockIdx.x] = t;
	}
}

// out[i] = offsets[tile] + exclusive prefix of in[] within the tile (in == out allowed). offsets == nullptr: 0.
__global__ void scanTiles(const uint32_t* in, uint64_t n, const uint32_t* __restrict__ offsets, uint32_t* out)
{
	__shared__ uint32_t warpSums[kScanThreads / 32];
	const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;   // blocked arrangement
	uint32_t v[kScanItems];
	uint32_t s = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		v[k] = (base + k < n) ? in[base + k] : 0u;
		s += v[k];
	}
	uint32_t incl = s;
	const int lane = threadIdx.x & 31;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) warpSums[threadIdx.x >> 5] = incl;
	__syncthreads();
	uint32_t before = offsets ? offsets[blockIdx.x] : 0u;
	for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += warpSums[w];
	uint32_t run = before + incl - s;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		if (base + k < n) out[base + k] = run;
		run += v[k];
	}
}

cudaError_t exclusiveScan(uint32_t* data, uint64_t n, uint32_t* scratch, cudaStream_t stream, uint64_t* launches)
{
	// scratch holds the tile sums of every level back to back: n/2048 + n/2048^2 + ... + 1 words.
	const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
	if (tiles <= 1) {
		scanTiles<<<1, kScanThreads, 0, stream>>>(data, n, nullptr, data);
		*launches += 1;
		return cudaGetLastError();
	}
	scanTileSums<<<(unsigned)tiles, kScanThreads, 0, stream>>>(data, n, scratch);
	cudaError_t e = exclusiveScan(scratch, tiles, scratch + tiles, stream, launches);
	if (e != cudaSuccess) return e;
	scanTiles<<<(unsigned)tiles, kScanThreads, 0, stream>>>(data, n, scratch, data);
	*launches += 2;
	return cudaGetLastError();
}

// Emit the distinct nodes in first-occurrence order with their children renumbered; the 256 material nodes first.
__global__ void bakeEmit(uint32_t n, const uint8_t* __restrict__ mark, const uint32_t* __restrict__ newIndex, const uint32_t* __restrict__ rep,
	const uint32_t* __restrict__ position, const uint32_t* __restrict__ canon, uint32_t* out)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (i < kMaterialCount) {
			NodeWords self;
#pragma unroll
			for (int k = 0; k < 8; k++) self.w[k] = i;
			storeNode(out, i, self);
			continue;
		}
		if (mark[i] == 0 || newIndex[i] != i) continue;
		NodeWords node = loadNode(canon, i);
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const uint32_t c = node.w[k];
			node.w[k] = c < kMaterialCount ? c : kMaterialCount + position[rep[c]];
		}
		storeNode(out, kMaterialCount + position[rep[i]], node);
	}
}

// result[0] = new root index
__global__ void bakeRoot(uint32_t root, const uint32_t* newIndex, const uint32_t* rep, const uint32_t* position, uint32_t* result)
{
	const uint32_t r = newIndex[root];
	result[0] = r < kMaterialCount ? r : (r >= kPending ? 0u : kMaterialCount + position[rep[r]]);   // unresolved: the caller reports the cycle
}

// findSubDAG (reference src/library/raytracing.cpp:43-87) on the device copy: one thread per root octant.
__global__ void subdagKernel(const uint32_t* __restrict__ nodes, uint32_t nodeCount, uint32_t root, const unsigned long long* rootPtr, SubDag* out, uint32_t* status)
{
	const uint32_t octant = threadIdx.x;
	if (octant >= 8) return;
	if (rootPtr) root = (uint32_t)*rootPtr;
	int height = 32;
	uint32_t lower[3] = { 0x80000000u, 0x80000000u, 0x80000000u };
	uint32_t only = octant;
	SubDag sd;
	sd.lower[0] = sd.lower[1] = sd.lower[2] = 0; sd.height = 0; sd.pad0 = 0; sd.node = 0; sd.pad1 = 0; sd.pad2 = 0;
	bool ok = root < nodeCount;
	uint32_t next = ok ? nodes[(size_t)root * 8 + only] : 0;
	uint32_t node = 0, occupied = 1;
	while (ok && occupied == 1) {
		height--;
		if (height < 0) { ok = false; break; }
		node = next;
		if (node >= nodeCount) { ok = false; break; }
		for (int a = 0; a < 3; a++) lower[a] ^= ((only >> a) & 1u) << height;
		occupied = 0;
		for (uint32_t c = 0; c < 8; c++) {
			const uint32_t child = nodes[(size_t)node * 8 + c];
			if (child > 0) { next = child; occupied++; only = c; }
		}
	}
	if (ok) {
		for (int a = 0; a < 3; a++) sd.lower[a] = (int32_t)lower[a];
		sd.height = height;
		sd.node = node;
	} else {
		atomicOr(status, 1u);
	}
	out[octant] = sd;
}

__global__ void bakeResults(const unsigned long long* counters, const uint32_t* rootOut, unsigned long long* results)
{
	results[0] = counters[0];
	results[1] = counters[1];
	results[2] = rootOut[0];
}

int gridFor(uint64_t n, int smCount)
{
	const uint64_t want = (n + 255) / 256;
	const uint64_t cap = (uint64_t)smCount * 8;     // 8 x 256 threads = every resident lane of an SM
	return (int)(want < cap ? (want ? want : 1) : cap);
}

} // namespace

// The scratch buffer, carved into 256-byte aligned sections.
struct BakeScratch {
	uint32_t* canon; uint32_t* newIndex; uint32_t* rep; uint32_t* flags; uint8_t* mark; uint32_t* table; uint32_t* scan;
	uint8_t* frontierA; uint8_t* frontierB; uint32_t* chunkRemaining; uint32_t* chunkPending;
	unsigned long long* counters; uint32_t* rootOut;
	size_t bytes;
};

static BakeScratch carve(uint8_t* base, uint64_t n, uint64_t tableSlots)
{
	size_t at = 0;
	auto take = [&](size_t bytes) { uint8_t* p = base + at; at += (bytes + 255) & ~(size_t)255; return p; };
	uint64_t scanWords = 0;
	for (uint64_t t = (n + kScanTile - 1) / kScanTile; ; t = (t + kScanTile - 1) / kScanTile) { scanWords += t; if (t <= 1) break; }
	BakeScratch s;
	s.canon = reinterpret_cast<uint32_t*>(take(n * 32));
	s.newIndex = reinterpret_cast<uint32_t*>(take(n * 4));
	s.rep = reinterpret_cast<uint32_t*>(take(n * 4));
	s.flags = reinterpret_cast<uint32_t*>(take(n * 4));
	s.mark = take(n);
	const uint64_t chunks = (n + kChunk - 1) / kChunk;
	s.frontierA = take(chunks);
	s.frontierB = take(chunks);
	s.chunkRemaining = reinterpret_cast<uint32_t*>(take(chunks * 4));
	s.chunkPending = reinterpret_cast<uint32_t*>(take(chunks * 4));
	s.table = reinterpret_cast<uint32_t*>(take(tableSlots * 4));
	s.scan = reinterpret_cast<uint32_t*>(take((scanWords + 8) * 4));
	s.counters = reinterpret_cast<unsigned long long*>(take(64 * 8));   // [0..3] counters, [8 + d] frontier of depth d
	s.rootOut = reinterpret_cast<uint32_t*>(take(256));
	s.bytes = at;
	return s;
}

size_t bakeScratchBytes(uint64_t n, uint64_t* tableSlots)
{
	uint64_t slots = 1024;
	while (slots < 2 * n) slots <<= 1;
	if (slots > 0x80000000ull) slots = 0x80000000ull;
	*tableSlots = slots;
	return carve(nullptr, n, slots).bytes;
}

// ---- dense grid -> octree ------------------------------------------------------------------------------------
// Height-1 nodes: the eight voxels of a 2x2x2 cell, child slot x | y << 1 | z << 2 (storage.cpp:57-67).
__global__ void denseLeafKernel(const uint8_t* __restrict__ voxels, uint32_t k, uint32_t* nodes)
{
	const uint32_t half = k - 1;                          // log2 of cells per side
	const uint64_t cells = 1ull << (3 * half);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t X = i & ((1ull << half) - 1), Y = (i >> half) & ((1ull << half) - 1), Z = i >> (2 * half);
		NodeWords n;
#pragma unroll
		for (int z = 0; z < 2; z++)
#pragma unroll
			for (int y = 0; y < 2; y++) {
				const uint64_t at = ((2 * Z + z) << (2 * k)) + ((2 * Y + y) << k) + 2 * X;
				const uint16_t pair = *reinterpret_cast<const uint16_t*>(voxels + at);
				n.w[z * 4 + y * 2 + 0] = pair & 0xffu;
				n.w[z * 4 + y * 2 + 1] = pair >> 8;
			}
		storeNode(nodes, (uint32_t)(kMaterialCount + i), n);
	}
}

// Height-j nodes (j >= 2) point at the eight height-(j-1) nodes of their cell: pure index arithmetic.
__global__ void denseInnerKernel(uint32_t cellsLog2, uint32_t base, uint32_t childBase, uint32_t* nodes)
{
	const uint64_t cells = 1ull << (3 * cellsLog2);
	const uint32_t childSide = cellsLog2 + 1;             // log2 of child cells per side
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t X = i & ((1ull << cellsLog2) - 1), Y = (i >> cellsLog2) & ((1ull << cellsLog2) - 1), Z = i >> (2 * cellsLog2);
		NodeWords n;
#pragma unroll
		for (int c = 0; c < 8; c++) {
			const uint64_t x = 2 * X + (c & 1), y = 2 * Y + ((c >> 1) & 1), z = 2 * Z + (c >> 2);
			n.w[c] = childBase + (uint32_t)((z << (2 * childSide)) + (y << childSide) + x);
		}
		storeNode(nodes, (uint32_t)(base + i), n);
	}
}

uint64_t denseNodeCount(uint32_t k)
{
	uint64_t n = kMaterialCount;
	for (uint32_t j = 1; j + 1 <= k; j++) n += 1ull << (3 * (k - j));
	return n + 8 * (33 - k);                                  // room for the host-built top of the tree
}

// The ancestors of a 2^k grid's eight height-(k-1) cubes up to the height-32 root: at most 8 chains of 33 - k nodes,
// shared where the cubes share a parent (one chain when the origin is a multiple of the side, up to eight when the
// grid straddles 0). A tiny trie built on the host, slot = next bit of the unsigned position (make_position_unsigned /
// extract_next_child_id, storage.cpp:45-67). `cubes[c]` is what stands for cube c = (z * 2 + y) * 2 + x (a node index
// or a material); the trie's node i will live at index base + i; returns the root's i.
uint32_t topTrie(uint32_t k, const int32_t origin[3], const uint32_t cubes[8], uint32_t base, std::vector<uint32_t>& words)
{
	struct Top { int h; uint64_t x, y, z; uint32_t child[8]; };
	std::vector<Top> top;
	auto nodeFor = [&](int h, uint64_t x, uint64_t y, uint64_t z) -> uint32_t {
		for (size_t i = 0; i < top.size(); i++) if (top[i].h == h && top[i].x == x && top[i].y == y && top[i].z == z) return (uint32_t)i;
		Top t; t.h = h; t.x = x; t.y = y; t.z = z;
		for (int c = 0; c < 8; c++) t.child[c] = 0;
		top.push_back(t);
		return (uint32_t)(top.size() - 1);
	};
	const uint64_t u[3] = { (uint64_t)((uint32_t)origin[0] + 0x80000000u), (uint64_t)((uint32_t)origin[1] + 0x80000000u), (uint64_t)((uint32_t)origin[2] + 0x80000000u) };
	const uint32_t rootAt = nodeFor(32, 0, 0, 0);
	for (uint32_t c = 0; c < 8; c++) {
		const uint64_t U[3] = { (u[0] >> (k - 1)) + (c & 1u), (u[1] >> (k - 1)) + ((c >> 1) & 1u), (u[2] >> (k - 1)) + (c >> 2) };
		uint32_t cur = rootAt;
		for (int h = 32; h >= (int)k; h--) {
			const int bit = h - (int)k;
			const uint32_t slot = (uint32_t)((U[0] >> bit) & 1u) | (uint32_t)(((U[1] >> bit) & 1u) << 1) | (uint32_t)(((U[2] >> bit) & 1u) << 2);
			if (h == (int)k) { top[cur].child[slot] = cubes[c]; break; }
			const uint32_t next = nodeFor(h - 1, U[0] >> bit, U[1] >> bit, U[2] >> bit);
			top[cur].child[slot] = base + next;
			cur = next;
		}
	}
	words.resize(top.size() * 8);
	for (size_t i = 0; i < top.size(); i++) for (int c = 0; c < 8; c++) words[i * 8 + c] = top[i].child[c];
	return rootAt;
}

// Levels 1 .. k-1 are written on the device; what is left are the eight height-(k-1) cubes the grid consists of.
// Their ancestors up to the height-32 root are topTrie's.
cudaError_t launchBuildDense(const uint8_t* voxels, uint32_t k, const int32_t origin[3], uint32_t* nodes, uint32_t* root, uint64_t* nodeCount,
	int smCount, cudaStream_t stream, uint64_t* launches, bool chain)
{
	// Material nodes are written by launchBake into ITS output; the input copy of them is never read (children < 256
	// are ids, not indices), so [0, 256) of `nodes` stays untouched.
	uint32_t base = kMaterialCount;
	const uint64_t leaves = 1ull << (3 * (k - 1));
	denseLeafKernel<<<gridFor(leaves, smCount), 256, 0, stream>>>(voxels, k, nodes);
	uint64_t count = 1;
	uint32_t childBase = base;
	base += (uint32_t)leaves;
	// chain == false (one brick of a larger grid, whose own position is dealt with by the caller): the levels go all the
	// way up to the single node that is the whole grid, and that node is the root.
	for (uint32_t j = 2; j + (chain ? 1u : 0u) <= k; j++) {
		const uint32_t cellsLog2 = k - j;
		const uint64_t cells = 1ull << (3 * cellsLog2);
		denseInnerKernel<<<gridFor(cells, smCount), 256, 0, stream>>>(cellsLog2, base, childBase, nodes);
		count++;
		childBase = base;
		base += (uint32_t)cells;
	}
	if (!chain) {
		*root = childBase;                // the level of one cell written last (k >= 2)
		*nodeCount = (uint64_t)base;
		if (launches) *launches += count;
		return cudaGetLastError();
	}
	// childBase .. childBase + 8: the height-(k-1) cubes, index (z * 2 + y) * 2 + x within the grid.
	uint32_t cubes[8];
	for (uint32_t c = 0; c < 8; c++) cubes[c] = childBase + c;
	std::vector<uint32_t> words;
	const uint32_t rootAt = topTrie(k, origin, cubes, base, words);
	cudaError_t e = cudaMemcpyAsync(nodes + (size_t)base * 8, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
	if (e != cudaSuccess) return e;
	if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;     // `words` is on our stack frame
	*root = base + rootAt;
	*nodeCount = (uint64_t)base + words.size() / 8;
	if (launches) *launches += count;
	return cudaGetLastError();
}

// Brick-wise dense build (cbq_build_dense past 1024^3).
// Append a brick's merged nodes [256, 256 + count) to the collection: child indices move by `delta`, materials stay.
__global__ void __launch_bounds__(256) appendNodesKernel(const uint32_t* __restrict__ src, uint64_t words, uint32_t delta, uint32_t* __restrict__ dst)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t c = src[i];
		dst[i] = (c >= kMaterialCount) ? c + delta : c;
	}
}

cudaError_t launchAppendNodes(const uint32_t* src, uint64_t count, uint32_t delta, uint32_t* dst, int smCount, cudaStream_t stream)
{
	if (count == 0) return cudaSuccess;
	appendNodesKernel<<<gridFor(count * 8, smCount), 256, 0, stream>>>(src, count * 8, delta, dst);
	return cudaGetLastError();
}

// Largest child word of a node array (cbq_upload_device / cbq_update_device: every child index must be < the node count,
// the kernels follow them unchecked). One 16-byte load per thread, a warp max, one atomicMax per warp.
__global__ void __launch_bounds__(256) maxChildKernel(const uint4* __restrict__ words, uint64_t quads, uint32_t* worst)
{
	uint32_t m = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint4 w = words[i];
		m = max(max(m, max(w.x, w.y)), max(w.z, w.w));
	}
	m = __reduce_max_sync(0xffffffffu, m);
	if ((threadIdx.x & 31u) == 0u && m != 0u) atomicMax(worst, m);
}

cudaError_t launchMaxChild(const uint32_t* nodes, uint64_t count, uint32_t* worst, int smCount, cudaStream_t stream)
{
	if (count == 0) return cudaSuccess;
	uint64_t blocks = (count * 2 + 255) / 256;
	if (blocks > (uint64_t)smCount * 8u) blocks = (uint64_t)smCount * 8u;
	maxChildKernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(nodes), count * 2, worst);
	return cudaGetLastError();
}

cudaError_t launchSubdags(const uint32_t* nodes, uint32_t nodeCount, uint32_t root, const unsigned long long* rootPtr, SubDag* out, uint32_t* status, cudaStream_t stream)
{
	subdagKernel<<<1, 32, 0, stream>>>(nodes, nodeCount, root, rootPtr, out, status);
	return cudaGetLastError();
}

// Enqueue the whole bake on `stream`. `scratch` is bakeScratchBytes(n) bytes; `out` has room for n nodes.
// Afterwards (stream order) results[0] = reachable nodes left without an id (0 unless the array has a cycle or is
// deeper than 32 levels), results[1] = distinct non-material nodes, results[2] = new root index,
// results[3] = nodes reachable from the root before merging.
cudaError_t launchBake(const uint32_t* nodes, uint64_t n64, uint32_t root, uint8_t* scratch, uint64_t tableSlots, uint32_t* out,
	unsigned long long* results, int smCount, cudaStream_t stream, uint64_t* launches, uint32_t levels)
{
	// `levels`: an upper bound on the number of node levels below (and including) the root; 33 fits any volume. The
	// brick-wise dense build knows its trees are 2^brickLog2 voxels tall and saves two thirds of the launches.
	if (levels == 0 || levels > 33) levels = 33;
	const uint32_t n = (uint32_t)n64;
	const BakeScratch sc = carve(scratch, n64, tableSlots);
	uint32_t* canon = sc.canon; uint32_t* newIndex = sc.newIndex; uint32_t* rep = sc.rep; uint32_t* flags = sc.flags;
	uint8_t* mark = sc.mark; uint32_t* table = sc.table; uint32_t* scanScratch = sc.scan;
	unsigned long long* counters = sc.counters; uint32_t* rootOut = sc.rootOut;

	const int grid = gridFor(n64, smCount);
	uint64_t count = 0;
	cudaError_t e;
	if ((e = cudaMemsetAsync(counters, 0, 64 * 8, stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(table, 0xff, tableSlots * 4, stream)) != cudaSuccess) return e;
	const uint32_t chunks = (uint32_t)((n64 + kChunk - 1) / kChunk);
	bakeInit<<<grid, 256, 0, stream>>>(n, root, mark, newIndex, rep, flags, sc.frontierA, sc.frontierB, sc.chunkRemaining, sc.chunkPending, chunks); count++;
	for (uint32_t depth = 1; depth <= levels; depth++) {
		uint8_t* cur = (depth & 1u) ? sc.frontierB : sc.frontierA;
		uint8_t* next = (depth & 1u) ? sc.frontierA : sc.frontierB;
		bakeReach<<<grid, 256, 0, stream>>>(nodes, n, depth, mark, cur, next, chunks); count++;
	}
	if ((e = cudaMemsetAsync(results, 0, 4 * sizeof(unsigned long long), stream)) != cudaSuccess) return e;
	bakeCountReached<<<grid, 256, 0, stream>>>(n, mark, sc.chunkRemaining, chunks, counters, results); count++;
	for (uint32_t pass = 0; pass < levels; pass++) {
		bakeResolve<<<grid, 256, 0, stream>>>(nodes, n, mark, newIndex, canon, sc.chunkRemaining, sc.chunkPending, chunks, counters);
		bakeInsert<<<grid, 256, 0, stream>>>(n, newIndex, canon, table, (uint32_t)(tableSlots - 1), rep, sc.chunkRemaining, sc.chunkPending, chunks, counters);
		count += 2;
	}
	bakeFlag<<<grid, 256, 0, stream>>>(n, mark, newIndex, rep, flags); count++;
	if ((e = exclusiveScan(flags, n64, scanScratch, stream, &count)) != cudaSuccess) return e;
	bakeEmit<<<grid, 256, 0, stream>>>(n, mark, newIndex, rep, flags, canon, out); count++;
	bakeRoot<<<1, 1, 0, stream>>>(root, newIndex, rep, flags, rootOut); count++;
	bakeResults<<<1, 1, 0, stream>>>(counters, rootOut, results); count++;
	if (launches) *launches += count;
	return cudaGetLastError();
}

} // namespace cbq

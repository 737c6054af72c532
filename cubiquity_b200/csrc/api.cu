// extern "C" layer of cubiquity_b200 (include/cubiquity_b200.h): context, DAG -> GPU serialisation
// with delta re-upload, batched ray cast, frame ray cast, path-traced render.
//
// Reference interfaces replaced (all under /root/reference):
//   upload      src/application/commands/view/gpu_pathtracing_viewer.cpp:43-67 (glBufferData of
//               NodeStore::rawBytesPtr(), the 8 SubDAGs and the colour table)
//   update      Viewer::onMouseButtonDown -> onVolumeModified (viewer.cpp:152-172,
//               pathtracing_demo.cpp:335-341); COW rules storage.cpp:152-167,298-303
//   trace       Cubiquity::intersectVolume, src/library/raytracing.cpp:397-478
//   render      PathtracingDemo::raytrace, src/application/commands/view/pathtracing_demo.cpp:214-229
#include "cbq_internal.h"

extern "C" const char* cbq_dag_error(void);     // host_shim.cpp: why the last cbq_dag_load / cbq_dag_save on this thread failed
#include "meshmath.cuh"
namespace cbq {                                  // voxelize_host.cpp
void analyseMesh(const Tri* tris, uint64_t n, cbq_mesh_info* info);
}
#include <vector>
namespace cbq {
void splitTriangles(const Tri* tris, const uint8_t* materials, uint64_t n, std::vector<Tri>& out, std::vector<uint8_t>& mats);
}

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_lastError;
void (*g_logHandler)(const char*) = nullptr;

int fail(int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
	if (g_logHandler) g_logHandler(buf);
	return code;
}

#define CBQ_CUDA(expr)                                                                            \
	do {                                                                                          \
		cudaError_t e_ = (expr);                                                                  \
		if (e_ != cudaSuccess) {                                                                  \
			return fail(e_ == cudaErrorMemoryAllocation ? CBQ_ERROR_OUT_OF_MEMORY : CBQ_ERROR_CUDA, \
				"%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);      \
		}                                                                                         \
	} while (0)

constexpr int kQueueSlots = 64;        // ticket counters, one per stream that has launched a trace (the kernel re-arms its own)
constexpr uint64_t kPipelineChunkMin = 1u << 18;  // rays per host<->device pipeline stage: at least 2^18 (6 MB in; 2^16 was 45 % slower end to end, API overhead per stage) ...
constexpr uint64_t kPipelineChunkMax = 1u << 21;  // ... and for big batches an eighth of the batch up to 2^21: a stage's kernel must fill the GPU, or the kernels, not the link, set the pace
                                                  // (132 M rays in 2^18-ray stages: 79.6 ms, kernel-bound at ~150 us a stage; the link needs 58 ms)
constexpr int kStages = 3;                     // staging buffers in flight

// findSubDAG (reference src/library/raytracing.cpp:43-87), host side, bounds-checked because the
// node array comes from outside.
bool findOneSubDag(const uint32_t* nodes, uint64_t nodeCount, uint32_t root, uint32_t octant, cbq::SubDag& out)
{
	int height = 32;
	uint32_t lower[3] = { 0x80000000u, 0x80000000u, 0x80000000u };
	uint32_t only = octant;
	if (root >= nodeCount) return false;
	uint32_t next = nodes[(uint64_t)root * 8 + only];
	uint32_t node = 0;
	uint32_t occupied = 1;
	while (occupied == 1) {
		height--;
		if (height < 0) return false;   // a chain of single children deeper than the tree: corrupt
		node = next;
		if (node >= nodeCount) return false;
		for (int a = 0; a < 3; a++) lower[a] ^= ((only >> a) & 1u) << height;
		occupied = 0;
		for (uint32_t c = 0; c < 8; c++) {
			const uint32_t child = nodes[(uint64_t)node * 8 + c];
			if (child > 0) { next = child; occupied++; only = c; }
		}
	}
	std::memset(&out, 0, sizeof(out));
	for (int a = 0; a < 3; a++) out.lower[a] = (int32_t)lower[a];
	out.height = height;
	out.node = node;
	return true;
}

// Every child index must address a node that exists: the kernels follow them unchecked.
bool childrenInRange(const uint32_t* nodes, uint64_t begin, uint64_t end, uint64_t nodeCount)
{
	const uint32_t limit = (uint32_t)std::min<uint64_t>(nodeCount, 0xffffffffull);
	uint32_t worst = 0;
	const uint32_t* p = nodes + begin * 8;
	const uint64_t words = (end - begin) * 8;
	for (uint64_t i = 0; i < words; i++) worst = std::max(worst, p[i]);
	return worst < limit || nodeCount > 0xffffffffull;
}

uint32_t fmix32Host(uint32_t h)   // glsl/pathtracing.frag:287-296 (bitMix)
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}

bool findSubDags(const uint32_t* nodes, uint64_t nodeCount, uint32_t root, cbq::SubDag out[8])
{
	for (uint32_t c = 0; c < 8; c++) if (!findOneSubDag(nodes, nodeCount, root, c, out[c])) return false;
	return true;
}

} // namespace

struct cbq_context {
	int device = 0;
	cudaDeviceProp prop{};
	cudaStream_t stream = nullptr;       // compute + default
	cudaStream_t copyIn = nullptr, copyOut = nullptr;
	cudaEvent_t evIn[kStages]{}, evKernel[kStages]{}, evOut[kStages]{};

	// Volume buffers and bake work space come from a private stream-ordered pool that keeps what is freed
	// (release threshold = max): a re-upload or a bake of a multi-GB DAG does not pay cudaMalloc/cudaFree each time.
	cudaMemPool_t pool = nullptr;
	int denseBrickLog2 = 0;              // cbq_build_dense: 0 = one piece up to 1024^3, bricks of 512^3 beyond; else always bricks of this size

	// The volume: one linear device buffer.
	uint8_t* volume = nullptr;
	size_t volumeBytes = 0;
	uint64_t nodeCount = 0, nodeCapacity = 0;
	uint32_t root = 0;
	uint64_t generation = 0;
	cbq::SubDag subdags[8]{};
	int maxSubDagHeight = 0;

	// Launch bookkeeping.
	unsigned long long* queues = nullptr; // kQueueSlots x {ticket counter, finished CTAs} + 1 abandoned counter at the end
	cudaStream_t queueStream[kQueueSlots]{};
	int queueUsed = 0;
	cbq::LaunchConfig cfg{};
	int l2Persist = 1;
	cudaStream_t windowStream[kQueueSlots]{};   // streams the access-policy window has been applied to, and for which generation
	uint64_t windowGeneration[kQueueSlots]{};
	int windowUsed = 0;
	size_t l2Carve = ~(size_t)0;

	// cbq_trace staging (device side), double buffered.
	cbq::Ray* stageRays[kStages]{};
	cbq::Hit* stageHits[kStages]{};
	uint64_t stageCapacity = 0;          // rays each staging buffer holds

	// cbq_raycast_frame_device: primary rays of the frame in 8x4-tile order
	cbq::Ray* frameRays = nullptr;
	size_t frameRayCapacity = 0;

	// tile-group rendering (cbq_pt_params::tile_group_count) and the progressive pass
	uint32_t* tileList = nullptr;
	size_t tileListCapacity = 0;
	float* progressiveScratch = nullptr;      // width x height x 3, all zeros between passes
	size_t progressiveScratchPixels = 0;

	// cbq_render staging
	float* stageAccum = nullptr;
	size_t stageAccumBytes = 0;
	cbq::WavefrontBuffers wavefront;

	// Cost feedback for coherent batches (refill threshold 32): the kernel records how long each 32-ray ticket
	// took; the next launch over the same batch (same ray buffer, size and stream) deals them longest first.
	std::vector<std::pair<uintptr_t, uintptr_t>> remote;   // [begin, end) of the buffers opened with cbq_shared_open
	int parkResults = 0;              // option "park_results": use the coalescing sink even for a local buffer (testing)
	int adaptiveOrder = 1;
	uint32_t* ticketCost = nullptr;
	uint32_t* ticketHist = nullptr;
	uint32_t* ticketOrder[2] = { nullptr, nullptr };
	uint64_t ticketCapacity = 0;
	int orderWhich = 0;
	int orderAge = 0;                 // launches since the order was last rebuilt from the recorded costs
	int orderRefresh = 4;             // rebuild it every this many launches (option "order_refresh")
	cbq_camera lastFrameCamera{};     // cbq_raycast_frame_device: a frame whose camera moved rebuilds the order at once
	cudaEvent_t orderEvent = nullptr; // end of the last launch that used the cost / order buffers (they are shared by all streams)
	bool orderEventSet = false;
	uint64_t orderTickets = 0;
	const void* orderRays = nullptr;
	cudaStream_t orderStream = nullptr;

	// Counters
	uint64_t launches = 0, raysTraced = 0, bytesH2D = 0, bytesD2H = 0;
	uint64_t bakeReachable = 0;   // nodes the root reached in the last cbq_bake, before merging
	uint64_t voxelizeLeaves = 0, voxelizePieces = 0;   // octree leaves classified / triangle pieces scan-converted by the last cbq_voxelize
	bool deviceDiverged = false;  // a device-side edit / bake / build made the device copy differ from any host array

	const uint32_t* nodesPtr() const { return reinterpret_cast<const uint32_t*>(volume + cbq::kNodeOffset); }
	const cbq::SubDag* subdagsPtr() const { return reinterpret_cast<const cbq::SubDag*>(volume + cbq::kSubDagOffset); }
	const float4* coloursPtr() const { return reinterpret_cast<const float4*>(volume + cbq::kColourOffset); }
	unsigned long long* abandonedPtr() const { return queues + 2 * kQueueSlots; }
	cbq::VolumeView view() const
	{
		return cbq::VolumeView{ nodesPtr(), subdagsPtr() };
	}
};

namespace {

int bind(cbq_context* ctx)
{
	if (!ctx) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null context");
	CBQ_CUDA(cudaSetDevice(ctx->device));
	return CBQ_OK;
}

cudaError_t poolAlloc(cbq_context* ctx, uint8_t** out, size_t bytes)
{
	return cudaMallocFromPoolAsync(reinterpret_cast<void**>(out), bytes, ctx->pool, ctx->stream);
}

void poolFree(cbq_context* ctx, void* p)
{
	if (p) cudaFreeAsync(p, ctx->stream);
}

// The ticket counter of `stream`. Launches on one stream are ordered and every trace kernel leaves its counter
// zeroed again (its last CTA re-arms it), so a stream needs exactly one; different streams never share one.
int nextQueue(cbq_context* ctx, cudaStream_t stream, unsigned long long** out)
{
	for (int i = 0; i < ctx->queueUsed; i++) if (ctx->queueStream[i] == stream) { *out = ctx->queues + 2 * i; return CBQ_OK; }
	if (ctx->queueUsed == kQueueSlots) {
		// More streams than slots (they may have been destroyed since): wait for everything, then start over.
		CBQ_CUDA(cudaDeviceSynchronize());
		ctx->queueUsed = 0;
	}
	ctx->queueStream[ctx->queueUsed] = stream;
	*out = ctx->queues + 2 * ctx->queueUsed++;
	return CBQ_OK;
}

int writeHeaderAndSubdags(cbq_context* ctx)
{
	cbq::VolumeHeader h;
	std::memset(&h, 0, sizeof(h));
	h.magic = 0x31514243u; h.version = CBQ_VERSION;
	h.nodeCount = ctx->nodeCount; h.nodeCapacity = ctx->nodeCapacity;
	h.rootIndex = ctx->root; h.maxSubDagHeight = (uint32_t)ctx->maxSubDagHeight; h.generation = ctx->generation;
	CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kHeaderOffset, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
	CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kSubDagOffset, ctx->subdags, sizeof(ctx->subdags), cudaMemcpyHostToDevice, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream)); // h is on our stack
	ctx->bytesH2D += sizeof(h) + sizeof(ctx->subdags);
	return CBQ_OK;
}

// Sub-DAGs of a host array, not yet adopted: callers commit them with adoptSubdags() once nothing can fail any more.
int computeSubdags(const uint32_t* nodes, uint64_t nodeCount, uint32_t root, cbq::SubDag sd[8])
{
	if (!findSubDags(nodes, nodeCount, root, sd))
		return fail(CBQ_ERROR_CORRUPT_VOLUME, "node array is not a valid DAG below root %u (child index out of range or runaway chain)", root);
	return CBQ_OK;
}

void adoptSubdags(cbq_context* ctx, const cbq::SubDag sd[8])
{
	int maxH = 0;
	for (int i = 0; i < 8; i++) if (sd[i].node > 0) maxH = std::max(maxH, sd[i].height);
	std::memcpy(ctx->subdags, sd, sizeof(ctx->subdags));
	ctx->maxSubDagHeight = maxH;
	ctx->cfg.stackLevels = maxH + 1;
}

// Pin the node array in L2 (persisting access-policy window) for kernels on `stream`. Applied once per stream and
// volume generation: the calls below are not free (cudaDeviceSetLimit synchronises the device), and a caller that
// alternates between streams must not pay them per launch.
void applyL2Window(cbq_context* ctx, cudaStream_t stream)
{
	for (int i = 0; i < ctx->windowUsed; i++) {
		if (ctx->windowStream[i] == stream) {
			if (ctx->windowGeneration[i] == ctx->generation) return;
			ctx->windowGeneration[i] = ctx->generation;
			goto apply;
		}
	}
	if (ctx->windowUsed == kQueueSlots) ctx->windowUsed = 0;
	ctx->windowStream[ctx->windowUsed] = stream;
	ctx->windowGeneration[ctx->windowUsed++] = ctx->generation;
apply:
	cudaStreamAttrValue attr;
	std::memset(&attr, 0, sizeof(attr));
	if (ctx->l2Persist && ctx->volume && ctx->prop.persistingL2CacheMaxSize > 0) {
		const size_t nodeBytes = (size_t)ctx->nodeCount * 32;
		const size_t window = std::min(nodeBytes + cbq::kNodeOffset, (size_t)ctx->prop.accessPolicyMaxWindowSize);
		const size_t carve = std::min((size_t)ctx->prop.persistingL2CacheMaxSize, window);
		if (carve != ctx->l2Carve) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve); ctx->l2Carve = carve; }
		attr.accessPolicyWindow.base_ptr = ctx->volume;
		attr.accessPolicyWindow.num_bytes = window;
		attr.accessPolicyWindow.hitRatio = window ? (float)std::min(1.0, (double)carve / (double)window) : 0.0f;
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	} else {
		attr.accessPolicyWindow.num_bytes = 0;
	}
	if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

int ensureStaging(cbq_context* ctx, uint64_t chunk)
{
	if (chunk <= ctx->stageCapacity) return CBQ_OK;
	CBQ_CUDA(cudaDeviceSynchronize());
	for (int i = 0; i < kStages; i++) {
		cudaFree(ctx->stageRays[i]); cudaFree(ctx->stageHits[i]);
		ctx->stageRays[i] = nullptr; ctx->stageHits[i] = nullptr;
	}
	ctx->stageCapacity = 0;
	for (int i = 0; i < kStages; i++) {
		CBQ_CUDA(cudaMalloc(&ctx->stageRays[i], chunk * sizeof(cbq::Ray)));
		CBQ_CUDA(cudaMalloc(&ctx->stageHits[i], chunk * sizeof(cbq::Hit)));
	}
	ctx->stageCapacity = chunk;
	return CBQ_OK;
}

// Cost-feedback scheduling: fills a.ticketOrder / a.ticketCost when this launch qualifies. Returns true if the
// launch must be followed by orderAfterTrace().
bool orderBeforeTrace(cbq_context* ctx, cbq::TraceArgs& a, const cbq::LaunchConfig& cfg, cudaStream_t stream)
{
	const uint64_t tickets = (a.count + 31) / 32;
	const uint64_t warps = (uint64_t)cfg.smCount * cfg.blocksPerSm * (cfg.blockThreads / 32);
	// Worth it only when a warp serves a handful of tickets (a tail exists) and one block can sort them.
	if (!ctx->adaptiveOrder || !a.rays || a.countPtr || cfg.refillThreshold < 32 || tickets < 2 * warps || tickets > (1u << 18)) return false;
	if (tickets > ctx->ticketCapacity) {
		if (cudaDeviceSynchronize() != cudaSuccess) return false;
		cudaFree(ctx->ticketCost); cudaFree(ctx->ticketHist); cudaFree(ctx->ticketOrder[0]); cudaFree(ctx->ticketOrder[1]);
		ctx->ticketCost = ctx->ticketHist = ctx->ticketOrder[0] = ctx->ticketOrder[1] = nullptr;
		ctx->ticketCapacity = 0; ctx->orderTickets = 0;
		if (cudaMalloc(&ctx->ticketCost, tickets * 4) != cudaSuccess || cudaMalloc(&ctx->ticketOrder[0], tickets * 4) != cudaSuccess ||
			cudaMalloc(&ctx->ticketOrder[1], tickets * 4) != cudaSuccess || cudaMalloc(&ctx->ticketHist, ((tickets + 1023) / 1024) * 256 * 4) != cudaSuccess) { cudaGetLastError(); return false; }
		ctx->ticketCapacity = tickets;
	}
	// One set of buffers serves every stream: a launch on another stream waits for the previous user (trace + sort).
	if (ctx->orderEventSet && ctx->orderStream != stream && cudaStreamWaitEvent(stream, ctx->orderEvent, 0) != cudaSuccess) { cudaGetLastError(); return false; }
	if (ctx->orderTickets == tickets && ctx->orderRays == a.rays && ctx->orderStream == stream) a.ticketOrder = ctx->ticketOrder[ctx->orderWhich];
	a.ticketCost = ctx->ticketCost;
	return true;
}

int orderAfterTrace(cbq_context* ctx, const cbq::TraceArgs& a, cudaStream_t stream, bool refreshNow = false)
{
	const uint64_t tickets = (a.count + 31) / 32;
	// Costs of an unchanged batch do not drift: rebuild the order on the first repeat and then every orderRefresh-th
	// launch (4). The frame call rebuilds after every frame whose camera moved (refreshNow): measured on an orbit at
	// 0.25 deg per frame, 4.94 Grays/s with a rebuild per frame, 4.76 every 4th, 4.51 without feedback.
	const bool sameBatch = ctx->orderTickets == tickets && ctx->orderRays == a.rays && ctx->orderStream == stream;
	const bool keep = sameBatch && a.ticketOrder && ++ctx->orderAge < ctx->orderRefresh && !refreshNow;
	if (!keep) {
		ctx->orderAge = 0;
		const int next = ctx->orderWhich ^ 1;
		CBQ_CUDA(cbq::launchOrderTickets(ctx->ticketCost, (uint32_t)tickets, ctx->ticketHist, ctx->ticketOrder[next], stream));
		ctx->orderWhich = next; ctx->orderTickets = tickets; ctx->orderRays = a.rays;
		ctx->launches += 2;
	}
	ctx->orderStream = stream;
	CBQ_CUDA(cudaEventRecord(ctx->orderEvent, stream));
	ctx->orderEventSet = true;
	return CBQ_OK;
}

int traceDevice(cbq_context* ctx, const cbq::Ray* dRays, uint64_t n, uint32_t flags, float maxFootprint,
	cbq::Hit* dHits, cudaStream_t stream, const cbq_camera* cam, uint32_t width, uint32_t height, cbq_hit_compact* dCompact = nullptr)
{
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	if (n == 0) return CBQ_OK;
	cbq::TraceArgs a;
	std::memset(&a, 0, sizeof(a));
	a.volume = ctx->view();
	a.rays = dRays; a.hits = dHits; a.compact = dCompact; a.count = n; a.maxFootprint = maxFootprint;
	if (dCompact) {
		// A result buffer opened from another process / GPU (cbq_shared_open) is written over NVLink: say so to the kernel.
		const uintptr_t p = reinterpret_cast<uintptr_t>(dCompact);
		for (const auto& r : ctx->remote) if (p >= r.first && p < r.second) a.remoteResults = true;
		if (ctx->parkResults) a.remoteResults = true;
	}
	a.abandoned = ctx->abandonedPtr();
	if (cam) { a.camera = *cam; a.width = width; a.height = height; }
	int rc = nextQueue(ctx, stream, &a.queue);
	if (rc != CBQ_OK) return rc;
	applyL2Window(ctx, stream);
	cbq::LaunchConfig cfg = ctx->cfg;
	if (cam) cfg.refillThreshold = 32;   // tile-ordered primary rays are coherent by construction: no mid-flight refill
	const bool feedback = orderBeforeTrace(ctx, a, cfg, stream);
	CBQ_CUDA(cbq::launchTrace(a, (flags & CBQ_TRACE_SURFACE) != 0, cfg, stream));
	ctx->launches++;
	ctx->raysTraced += n;
	if (feedback) return orderAfterTrace(ctx, a, stream);
	return CBQ_OK;
}

// Hash-cons merge of the device array `dNodes` (cbq::launchBake) and installation of the result as the context's
// volume: new buffer from the pool, colours carried over, sub-DAGs recomputed on the device.
int bakeAndInstall(cbq_context* ctx, const uint32_t* dNodes, uint64_t n, uint32_t root, const char* what)
{
	uint64_t slots = 0;
	const size_t scratchBytes = cbq::bakeScratchBytes(n, &slots);
	const size_t tailBytes = 4 * sizeof(unsigned long long) + 8 * sizeof(cbq::SubDag) + 64;   // results, sub-DAGs, status
	size_t newBytes = cbq::kNodeOffset + (size_t)n * 32;     // the merged array is never longer than the input
	uint8_t* scratch = nullptr;
	uint8_t* baked = nullptr;
	CBQ_CUDA(poolAlloc(ctx, &scratch, scratchBytes + tailBytes));
	if (poolAlloc(ctx, &baked, newBytes) != cudaSuccess) { cudaGetLastError(); poolFree(ctx, scratch); return fail(CBQ_ERROR_OUT_OF_MEMORY, "%s: no room for the merged copy of the volume (%zu bytes)", what, newBytes); }
	unsigned long long* dResults = reinterpret_cast<unsigned long long*>(scratch + scratchBytes);
	cbq::SubDag* dSubdags = reinterpret_cast<cbq::SubDag*>(dResults + 4);
	uint32_t* dStatus = reinterpret_cast<uint32_t*>(dSubdags + 8);
	uint32_t* outNodes = reinterpret_cast<uint32_t*>(baked + cbq::kNodeOffset);

	struct { unsigned long long results[4]; cbq::SubDag subdags[8]; uint32_t status; } host;
	cudaError_t e = cudaMemsetAsync(dStatus, 0, 64, ctx->stream);
	if (e == cudaSuccess) e = cbq::launchBake(dNodes, n, root, scratch, slots, outNodes, dResults, ctx->cfg.smCount, ctx->stream, &ctx->launches);
	if (e == cudaSuccess) e = cbq::launchSubdags(outNodes, (uint32_t)n, 0, dResults + 2, dSubdags, dStatus, ctx->stream);
	if (e == cudaSuccess && ctx->volume) e = cudaMemcpyAsync(baked + cbq::kColourOffset, ctx->volume + cbq::kColourOffset, cbq::kNodeOffset - cbq::kColourOffset, cudaMemcpyDeviceToDevice, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&host, dResults, sizeof(host.results) + sizeof(host.subdags) + sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	poolFree(ctx, scratch);
	if (e != cudaSuccess) { poolFree(ctx, baked); return fail(CBQ_ERROR_CUDA, "%s failed: %s", what, cudaGetErrorString(e)); }
	ctx->launches += 1;
	ctx->bytesD2H += sizeof(host);
	if (host.results[0] != 0) {
		poolFree(ctx, baked);
		return fail(CBQ_ERROR_CORRUPT_VOLUME, "%s: %llu reachable nodes never resolved (a cycle, or a DAG deeper than 32 levels); the volume is unchanged",
			what, host.results[0]);
	}
	if (host.status != 0) { poolFree(ctx, baked); return fail(CBQ_ERROR_CORRUPT_VOLUME, "%s: the merged array has no valid sub-DAGs; the volume is unchanged", what); }

	const uint64_t count = cbq::kMaterialCount + host.results[1];
	uint64_t capacity = n;
	const uint64_t wanted = count + std::max<uint64_t>(count / 4, 1u << 16);     // the head-room cbq_upload gives
	if (capacity > 2 * wanted) {
		// Mostly merged away (a dense grid, a long edit session): move into a buffer of the usual size.
		uint8_t* snug = nullptr;
		const size_t snugBytes = cbq::kNodeOffset + (size_t)wanted * 32;
		if (poolAlloc(ctx, &snug, snugBytes) == cudaSuccess) {
			CBQ_CUDA(cudaMemcpyAsync(snug, baked, cbq::kNodeOffset + (size_t)count * 32, cudaMemcpyDeviceToDevice, ctx->stream));
			poolFree(ctx, baked);
			baked = snug; newBytes = snugBytes; capacity = wanted;
		} else cudaGetLastError();
	}

	CBQ_CUDA(cudaDeviceSynchronize());           // nobody may still be reading the old copy
	poolFree(ctx, ctx->volume);
	ctx->volume = baked;
	ctx->volumeBytes = newBytes;
	ctx->nodeCapacity = capacity;
	ctx->nodeCount = count;
	ctx->root = (uint32_t)host.results[2];
	ctx->generation++;
	ctx->bakeReachable = host.results[3];
	ctx->deviceDiverged = true;
	std::memcpy(ctx->subdags, host.subdags, sizeof(host.subdags));
	int maxH = 0;
	for (int i = 0; i < 8; i++) if (ctx->subdags[i].node > 0) maxH = std::max(maxH, ctx->subdags[i].height);
	ctx->maxSubDagHeight = maxH;
	ctx->cfg.stackLevels = maxH + 1;
	return writeHeaderAndSubdags(ctx);
}

// findSubDAGs on the device copy for `root` (read from *dRoot when dRoot != nullptr), result adopted by the context.
int refreshSubdagsOnDevice(cbq_context* ctx, uint32_t root, const unsigned long long* dRoot, uint8_t* dScratch /* >= 512 bytes */, const char* what)
{
	cbq::SubDag* dSubdags = reinterpret_cast<cbq::SubDag*>(dScratch);
	uint32_t* dStatus = reinterpret_cast<uint32_t*>(dScratch + 256);
	struct { cbq::SubDag subdags[8]; uint32_t status; } host;
	CBQ_CUDA(cudaMemsetAsync(dStatus, 0, 64, ctx->stream));
	CBQ_CUDA(cbq::launchSubdags(ctx->nodesPtr(), (uint32_t)ctx->nodeCapacity, root, dRoot, dSubdags, dStatus, ctx->stream));
	CBQ_CUDA(cudaMemcpyAsync(&host, dSubdags, sizeof(host.subdags) + sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->launches += 1;
	ctx->bytesD2H += sizeof(host);
	if (host.status != 0) return fail(CBQ_ERROR_CORRUPT_VOLUME, "%s: no valid sub-DAGs below the root", what);
	std::memcpy(ctx->subdags, host.subdags, sizeof(host.subdags));
	int maxH = 0;
	for (int i = 0; i < 8; i++) if (ctx->subdags[i].node > 0) maxH = std::max(maxH, ctx->subdags[i].height);
	ctx->maxSubDagHeight = maxH;
	ctx->cfg.stackLevels = maxH + 1;
	return CBQ_OK;
}

// More room for nodes, keeping what is there.
int growVolume(cbq_context* ctx, uint64_t capacity)
{
	const size_t bytes = cbq::kNodeOffset + (size_t)capacity * 32;
	uint8_t* bigger = nullptr;
	CBQ_CUDA(cudaDeviceSynchronize());
	CBQ_CUDA(poolAlloc(ctx, &bigger, bytes));
	CBQ_CUDA(cudaMemcpyAsync(bigger, ctx->volume, cbq::kNodeOffset + (size_t)ctx->nodeCount * 32, cudaMemcpyDeviceToDevice, ctx->stream));
	poolFree(ctx, ctx->volume);
	ctx->volume = bigger; ctx->volumeBytes = bytes; ctx->nodeCapacity = capacity;
	return CBQ_OK;
}

} // namespace

extern "C" {

const char* cbq_last_error(void) { return g_lastError.c_str(); }

int cbq_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int cbq_create(int device, cbq_context** out)
{
	if (!out) return fail(CBQ_ERROR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	const int n = cbq_device_count();
	if (n <= 0) return fail(CBQ_ERROR_NO_DEVICE, "no CUDA device is visible; cubiquity_b200 has no CPU fallback");
	if (device < 0 || device >= n) return fail(CBQ_ERROR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, n - 1);
	CBQ_CUDA(cudaSetDevice(device));
	cbq_context* ctx = new cbq_context();
	ctx->device = device;
	CBQ_CUDA(cudaGetDeviceProperties(&ctx->prop, device));
	if (ctx->prop.major < 10)
		{ const int maj = ctx->prop.major, mnr = ctx->prop.minor; delete ctx; return fail(CBQ_ERROR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, maj, mnr); }
	CBQ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	CBQ_CUDA(cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking));
	CBQ_CUDA(cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking));
	for (int i = 0; i < kStages; i++) {
		CBQ_CUDA(cudaEventCreateWithFlags(&ctx->evIn[i], cudaEventDisableTiming));
		CBQ_CUDA(cudaEventCreateWithFlags(&ctx->evKernel[i], cudaEventDisableTiming));
		CBQ_CUDA(cudaEventCreateWithFlags(&ctx->evOut[i], cudaEventDisableTiming));
	}
	{
		cudaMemPoolProps props;
		std::memset(&props, 0, sizeof(props));
		props.allocType = cudaMemAllocationTypePinned;
		props.handleTypes = cudaMemHandleTypeNone;
		props.location.type = cudaMemLocationTypeDevice;
		props.location.id = device;
		CBQ_CUDA(cudaMemPoolCreate(&ctx->pool, &props));
		uint64_t keep = ~0ull;
		CBQ_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
	}
	CBQ_CUDA(cudaEventCreateWithFlags(&ctx->orderEvent, cudaEventDisableTiming));
	CBQ_CUDA(cudaMalloc(&ctx->queues, sizeof(unsigned long long) * (2 * kQueueSlots + 1)));
	CBQ_CUDA(cudaMemset(ctx->queues, 0, sizeof(unsigned long long) * (2 * kQueueSlots + 1)));
	ctx->cfg.blockThreads = 256;
	ctx->cfg.blocksPerSm = 4;
	ctx->cfg.smCount = ctx->prop.multiProcessorCount;
	ctx->cfg.refillThreshold = 8;    // robust default: +68 % on incoherent rays, -7 % on coherent ones (profiles/r01_sweeps.md)
	ctx->cfg.refillQuantum = 1;
	ctx->cfg.stackLevels = 33;
	ctx->cfg.secondaryRefill = 4; // 1080p, 16 spp, 4 bounces: 994-999 M spp/s for 1..6, 990 at 8, 967 at 12 (profiles/r02_analysis.md)
	ctx->cfg.sampleGroup = 0;    // auto; 1080p, 4 bounces: 915 / 965 M spp/s for groups of 8 / 16 (profiles/r02_analysis.md)
	*out = ctx;
	return CBQ_OK;
}

void cbq_destroy(cbq_context* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (int i = 0; i < kStages; i++) {
		cudaFree(ctx->stageRays[i]); cudaFree(ctx->stageHits[i]);
		if (ctx->evIn[i]) cudaEventDestroy(ctx->evIn[i]);
		if (ctx->evKernel[i]) cudaEventDestroy(ctx->evKernel[i]);
		if (ctx->evOut[i]) cudaEventDestroy(ctx->evOut[i]);
	}
	cudaFree(ctx->stageAccum);
	cudaFree(ctx->tileList);
	cudaFree(ctx->progressiveScratch);
	cudaFree(ctx->frameRays);
	cudaFree(ctx->ticketCost); cudaFree(ctx->ticketHist); cudaFree(ctx->ticketOrder[0]); cudaFree(ctx->ticketOrder[1]);
	if (ctx->orderEvent) cudaEventDestroy(ctx->orderEvent);
	cbq::wavefrontRelease(ctx->wavefront);
	cudaFree(ctx->queues);
	poolFree(ctx, ctx->volume);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
	if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
	delete ctx;
}

int cbq_synchronize(cbq_context* ctx)
{
	int rc = bind(ctx); if (rc) return rc;
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	return CBQ_OK;
}

int cbq_find_subdags(const uint32_t* nodes, uint64_t node_count, uint32_t root_index, cbq_subdag out[8])
{
	if (!nodes || !out || node_count < cbq::kMaterialCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad node array");
	static_assert(sizeof(cbq_subdag) == sizeof(cbq::SubDag), "subdag layout");
	if (!findSubDags(nodes, node_count, root_index, reinterpret_cast<cbq::SubDag*>(out)))
		return fail(CBQ_ERROR_CORRUPT_VOLUME, "node array is not a valid DAG below root %u", root_index);
	return CBQ_OK;
}

int cbq_upload(cbq_context* ctx, const uint32_t* nodes, uint64_t node_count, uint32_t root_index, const float* colours_rgb)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!nodes || node_count < cbq::kMaterialCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node array must include the 256 material nodes");
	if (node_count > 0xffffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node indices are 32-bit");
	if (!childrenInRange(nodes, 0, node_count, node_count))
		return fail(CBQ_ERROR_CORRUPT_VOLUME, "a child index is >= the node count %llu", (unsigned long long)node_count);
	cbq::SubDag sd[8];
	rc = computeSubdags(nodes, node_count, root_index, sd); if (rc) return rc;

	// Head-room for copy-on-write growth so that edits rarely force a reallocation.
	const uint64_t capacity = node_count + std::max<uint64_t>(node_count / 4, 1u << 16);
	const size_t bytes = cbq::kNodeOffset + (size_t)capacity * 32;
	CBQ_CUDA(cudaDeviceSynchronize());   // a kernel on a caller's stream may still be reading the volume
	if (bytes > ctx->volumeBytes) {
		if (ctx->volume) { poolFree(ctx, ctx->volume); ctx->volume = nullptr; ctx->volumeBytes = 0; ctx->nodeCount = 0; }
		CBQ_CUDA(poolAlloc(ctx, &ctx->volume, bytes));
		ctx->volumeBytes = bytes;
	}
	CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kNodeOffset, nodes, (size_t)node_count * 32, cudaMemcpyHostToDevice, ctx->stream));
	ctx->bytesH2D += node_count * 32;
	ctx->nodeCapacity = (ctx->volumeBytes - cbq::kNodeOffset) / 32;
	ctx->nodeCount = node_count;
	ctx->root = root_index;
	ctx->generation++;
	ctx->deviceDiverged = false;
	adoptSubdags(ctx, sd);
	rc = cbq_set_colours(ctx, colours_rgb); if (rc) return rc;
	return writeHeaderAndSubdags(ctx);
}

int cbq_update(cbq_context* ctx, const uint32_t* nodes, uint64_t dirty_begin, uint64_t node_count, uint32_t root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "cbq_update before cbq_upload");
	if (ctx->deviceDiverged) return fail(CBQ_ERROR_INVALID_ARGUMENT, "the device copy was changed on the device (cbq_fill_sphere / cbq_bake / cbq_build_dense): "
		"a host array can no longer be applied as a delta; cbq_upload it, or cbq_download_nodes first");
	if (!nodes || node_count < cbq::kMaterialCount || dirty_begin > node_count) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad dirty range");
	if (dirty_begin > ctx->nodeCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "dirty_begin %llu is past the %llu nodes on the device", (unsigned long long)dirty_begin, (unsigned long long)ctx->nodeCount);
	if (node_count > 0xffffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node indices are 32-bit");
	if (!childrenInRange(nodes, dirty_begin, node_count, node_count))
		return fail(CBQ_ERROR_CORRUPT_VOLUME, "a child index in the dirty tail is >= the node count %llu", (unsigned long long)node_count);
	cbq::SubDag sd[8];
	rc = computeSubdags(nodes, node_count, root_index, sd); if (rc) return rc;
	CBQ_CUDA(cudaDeviceSynchronize());   // a kernel on a caller's stream may still be reading the tail we overwrite
	if (node_count > ctx->nodeCapacity) {
		// Out of head-room: grow, keeping the clean prefix that is already on the device.
		const uint64_t capacity = node_count + std::max<uint64_t>(node_count / 4, 1u << 16);
		const size_t bytes = cbq::kNodeOffset + (size_t)capacity * 32;
		uint8_t* bigger = nullptr;
		CBQ_CUDA(poolAlloc(ctx, &bigger, bytes));
		CBQ_CUDA(cudaMemcpyAsync(bigger, ctx->volume, cbq::kNodeOffset + (size_t)dirty_begin * 32, cudaMemcpyDeviceToDevice, ctx->stream));
		poolFree(ctx, ctx->volume);
		ctx->volume = bigger; ctx->volumeBytes = bytes; ctx->nodeCapacity = capacity;
	}
	const uint64_t tail = node_count - dirty_begin;
	if (tail) {
		CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kNodeOffset + (size_t)dirty_begin * 32, nodes + dirty_begin * 8, (size_t)tail * 32,
			cudaMemcpyHostToDevice, ctx->stream));
		ctx->bytesH2D += tail * 32;
	}
	ctx->nodeCount = node_count;
	ctx->root = root_index;
	ctx->generation++;
	adoptSubdags(ctx, sd);
	return writeHeaderAndSubdags(ctx);
}

namespace {

// Child-index check and findSubDAGs for a node array that is already in the volume buffer: nodes [begin, count) are
// scanned for a child >= count, the sub-DAGs of `root` are computed by subdagKernel. On success `sd` holds them.
int checkOnDevice(cbq_context* ctx, uint64_t begin, uint64_t count, uint32_t root, cbq::SubDag sd[8], const char* what)
{
	uint8_t* work = nullptr;
	CBQ_CUDA(poolAlloc(ctx, &work, 1024));
	cbq::SubDag* dSubdags = reinterpret_cast<cbq::SubDag*>(work);
	uint32_t* dStatus = reinterpret_cast<uint32_t*>(work + 256);       // [0] sub-DAG status, [1] largest child word
	struct { cbq::SubDag subdags[8]; uint32_t status, worst; } host;
	cudaError_t e = cudaMemsetAsync(dStatus, 0, 64, ctx->stream);
	if (e == cudaSuccess) e = cbq::launchMaxChild(ctx->nodesPtr() + begin * 8, count - begin, dStatus + 1, ctx->cfg.smCount, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&host.worst, dStatus + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e == cudaSuccess && host.worst >= count) { poolFree(ctx, work); return fail(CBQ_ERROR_CORRUPT_VOLUME, "%s: a child index (%u) is >= the node count %llu", what, host.worst, (unsigned long long)count); }
	// only now is it safe to walk the chains
	if (e == cudaSuccess) e = cbq::launchSubdags(ctx->nodesPtr(), (uint32_t)count, root, nullptr, dSubdags, dStatus, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&host, dSubdags, sizeof(host.subdags) + sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	poolFree(ctx, work);
	if (e != cudaSuccess) return fail(CBQ_ERROR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
	ctx->launches += 2;
	ctx->bytesD2H += sizeof(host);
	if (host.status != 0) return fail(CBQ_ERROR_CORRUPT_VOLUME, "%s: node array is not a valid DAG below root %u (runaway chain)", what, root);
	std::memcpy(sd, host.subdags, sizeof(host.subdags));
	return CBQ_OK;
}

} // namespace

int cbq_upload_device(cbq_context* ctx, const uint32_t* d_nodes, uint64_t node_count, uint32_t root_index, const float* colours_rgb, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_nodes || node_count < cbq::kMaterialCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node array must include the 256 material nodes");
	if (node_count > 0xffffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node indices are 32-bit");
	if (root_index >= node_count) return fail(CBQ_ERROR_CORRUPT_VOLUME, "root %u is >= the node count %llu", root_index, (unsigned long long)node_count);
	const uint64_t capacity = node_count + std::max<uint64_t>(node_count / 4, 1u << 16);
	const size_t bytes = cbq::kNodeOffset + (size_t)capacity * 32;
	CBQ_CUDA(cudaDeviceSynchronize());   // the producer of d_nodes has finished; nobody is still reading the old volume
	(void)stream;
	if (bytes > ctx->volumeBytes) {
		if (ctx->volume) { poolFree(ctx, ctx->volume); ctx->volume = nullptr; ctx->volumeBytes = 0; }
		CBQ_CUDA(poolAlloc(ctx, &ctx->volume, bytes));
		ctx->volumeBytes = bytes;
	}
	// From here on the old volume is gone: a failure below leaves the context without one.
	ctx->nodeCount = 0;
	CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kNodeOffset, d_nodes, (size_t)node_count * 32, cudaMemcpyDeviceToDevice, ctx->stream));
	ctx->nodeCapacity = (ctx->volumeBytes - cbq::kNodeOffset) / 32;
	cbq::SubDag sd[8];
	rc = checkOnDevice(ctx, 0, node_count, root_index, sd, "cbq_upload_device");
	if (rc) { poolFree(ctx, ctx->volume); ctx->volume = nullptr; ctx->volumeBytes = 0; return rc; }
	ctx->nodeCount = node_count;
	ctx->root = root_index;
	ctx->generation++;
	ctx->deviceDiverged = false;
	adoptSubdags(ctx, sd);
	rc = cbq_set_colours(ctx, colours_rgb); if (rc) return rc;
	return writeHeaderAndSubdags(ctx);
}

int cbq_update_device(cbq_context* ctx, const uint32_t* d_tail, uint64_t dirty_begin, uint64_t node_count, uint32_t root_index, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "cbq_update_device before cbq_upload");
	if (ctx->deviceDiverged) return fail(CBQ_ERROR_INVALID_ARGUMENT, "the device copy was changed on the device (cbq_fill_sphere / cbq_bake / cbq_build_dense): "
		"no array is a delta of it any more; upload it again");
	if (node_count < cbq::kMaterialCount || dirty_begin > node_count || (dirty_begin < node_count && !d_tail)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad dirty range");
	if (dirty_begin > ctx->nodeCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "dirty_begin %llu is past the %llu nodes on the device", (unsigned long long)dirty_begin, (unsigned long long)ctx->nodeCount);
	if (node_count > 0xffffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "node indices are 32-bit");
	if (root_index >= node_count) return fail(CBQ_ERROR_CORRUPT_VOLUME, "root %u is >= the node count %llu", root_index, (unsigned long long)node_count);
	CBQ_CUDA(cudaDeviceSynchronize());
	(void)stream;
	if (node_count > ctx->nodeCapacity) {
		const uint64_t capacity = node_count + std::max<uint64_t>(node_count / 4, 1u << 16);
		const size_t bytes = cbq::kNodeOffset + (size_t)capacity * 32;
		uint8_t* bigger = nullptr;
		CBQ_CUDA(poolAlloc(ctx, &bigger, bytes));
		CBQ_CUDA(cudaMemcpyAsync(bigger, ctx->volume, cbq::kNodeOffset + (size_t)dirty_begin * 32, cudaMemcpyDeviceToDevice, ctx->stream));
		poolFree(ctx, ctx->volume);
		ctx->volume = bigger; ctx->volumeBytes = bytes; ctx->nodeCapacity = capacity;
	}
	const uint64_t tail = node_count - dirty_begin;
	if (tail) CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kNodeOffset + (size_t)dirty_begin * 32, d_tail, (size_t)tail * 32, cudaMemcpyDeviceToDevice, ctx->stream));
	cbq::SubDag sd[8];
	rc = checkOnDevice(ctx, dirty_begin, node_count, root_index, sd, "cbq_update_device");
	if (rc) return rc;      // the tail is bad: the header still describes the previous, intact state
	ctx->nodeCount = node_count;
	ctx->root = root_index;
	ctx->generation++;
	adoptSubdags(ctx, sd);
	return writeHeaderAndSubdags(ctx);
}

int cbq_bake(cbq_context* ctx, uint64_t* node_count, uint32_t* root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "cbq_bake before cbq_upload");
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	rc = bakeAndInstall(ctx, ctx->nodesPtr(), ctx->nodeCount, ctx->root, "cbq_bake"); if (rc) return rc;
	if (node_count) *node_count = ctx->nodeCount;
	if (root_index) *root_index = ctx->root;
	return CBQ_OK;
}

namespace {

// cbq_build_dense past 1024^3 (the complete octree of a 2048^3 grid has more nodes than a 32-bit index reaches, that
// of a 4096^3 grid would not fit the memory either): the grid is cut into bricks of 2^brickLog2 voxels a side. Every
// brick is built and merged on its own with the kernels of the one-piece build, its merged nodes are appended to one
// collection (child indices shifted), and what is left above the bricks -- a complete octree over (side / brick)^3
// references plus the chains to the height-32 root -- is a few hundred nodes made on the host. One last merge over the
// collection removes what the bricks have in common. Same result as the one-piece build: the merged DAG of a voxel
// set does not depend on the order it was assembled in (tests/test_gpu_update.py builds one grid both ways).
int buildDenseBricked(cbq_context* ctx, const uint8_t* dVoxels, uint32_t k, const int32_t origin[3], uint32_t brickLog2)
{
	const uint32_t perAxisLog2 = k - brickLog2, perAxis = 1u << perAxisLog2;
	const size_t side = (size_t)1 << k, brickSide = (size_t)1 << brickLog2;
	const uint64_t treeNodes = cbq::denseNodeCount(brickLog2);
	uint64_t slots = 0;
	const size_t scratchBytes = cbq::bakeScratchBytes(treeNodes, &slots);
	cudaStream_t s = ctx->stream;
	uint8_t *dBrick = nullptr, *dTree = nullptr, *dScratch = nullptr, *dMerged = nullptr, *dAll = nullptr;
	auto cleanup = [&]() { poolFree(ctx, dBrick); poolFree(ctx, dTree); poolFree(ctx, dScratch); poolFree(ctx, dMerged); poolFree(ctx, dAll); };
#define CBQ_BRICK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return fail(e_ == cudaErrorMemoryAllocation ? CBQ_ERROR_OUT_OF_MEMORY : CBQ_ERROR_CUDA, "cbq_build_dense: %s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
	CBQ_BRICK(poolAlloc(ctx, &dBrick, brickSide * brickSide * brickSide));
	CBQ_BRICK(poolAlloc(ctx, &dTree, (size_t)treeNodes * 32));
	CBQ_BRICK(poolAlloc(ctx, &dScratch, scratchBytes + 256));
	CBQ_BRICK(poolAlloc(ctx, &dMerged, (size_t)treeNodes * 32));
	unsigned long long* dResults = reinterpret_cast<unsigned long long*>(dScratch + scratchBytes);   // launchBake's four
	uint64_t capacity = std::max<uint64_t>(treeNodes, 1u << 22), count = cbq::kMaterialCount;      // [0, 256) of a merge input is never read
	CBQ_BRICK(poolAlloc(ctx, &dAll, (size_t)capacity * 32));

	std::vector<uint32_t> level((size_t)perAxis * perAxis * perAxis);      // what stands for each brick: a node of the collection, or a material
	for (uint32_t bz = 0; bz < perAxis; bz++) for (uint32_t by = 0; by < perAxis; by++) for (uint32_t bx = 0; bx < perAxis; bx++) {
		cudaMemcpy3DParms cp;
		std::memset(&cp, 0, sizeof(cp));
		cp.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(dVoxels), side, side, side);
		cp.srcPos = make_cudaPos(bx * brickSide, by * brickSide, bz * brickSide);
		cp.dstPtr = make_cudaPitchedPtr(dBrick, brickSide, brickSide, brickSide);
		cp.extent = make_cudaExtent(brickSide, brickSide, brickSide);
		cp.kind = cudaMemcpyDeviceToDevice;
		CBQ_BRICK(cudaMemcpy3DAsync(&cp, s));
		const int32_t brickOrigin[3] = { (int32_t)(origin[0] + (int64_t)(bx * brickSide)), (int32_t)(origin[1] + (int64_t)(by * brickSide)), (int32_t)(origin[2] + (int64_t)(bz * brickSide)) };
		uint32_t root = 0;
		uint64_t n = treeNodes;
		// No chain to the height-32 root per brick (its place in the volume is the top's business, below): the tree is
		// brickLog2 levels tall, its root IS the brick, and the merge needs that many passes instead of 33.
		CBQ_BRICK(cbq::launchBuildDense(dBrick, brickLog2, brickOrigin, reinterpret_cast<uint32_t*>(dTree), &root, &n, ctx->cfg.smCount, s, &ctx->launches, false));
		CBQ_BRICK(cbq::launchBake(reinterpret_cast<const uint32_t*>(dTree), n, root, dScratch, slots, reinterpret_cast<uint32_t*>(dMerged), dResults, ctx->cfg.smCount, s, &ctx->launches,
			brickLog2 + 2));
		unsigned long long host[4];
		CBQ_BRICK(cudaMemcpyAsync(host, dResults, sizeof(host), cudaMemcpyDeviceToHost, s));
		CBQ_BRICK(cudaStreamSynchronize(s));
		ctx->launches += 1; ctx->bytesD2H += sizeof(host);
		if (host[0] != 0) { cleanup(); return fail(CBQ_ERROR_CORRUPT_VOLUME, "cbq_build_dense: brick (%u, %u, %u) did not merge", bx, by, bz); }
		const uint64_t merged = host[1];
		const uint32_t top = (uint32_t)host[2];      // the merged brick's root: a node, or a material if the brick is uniform
		if (count + merged + 4096 > 0xffffffffull) { cleanup(); return fail(CBQ_ERROR_OUT_OF_MEMORY, "cbq_build_dense: more than 2^32 nodes before the last merge"); }
		if (count + merged > capacity) {
			const uint64_t bigger = std::max<uint64_t>(capacity * 2, count + merged);
			uint8_t* grown = nullptr;
			CBQ_BRICK(poolAlloc(ctx, &grown, (size_t)bigger * 32));
			CBQ_BRICK(cudaMemcpyAsync(grown, dAll, (size_t)count * 32, cudaMemcpyDeviceToDevice, s));
			poolFree(ctx, dAll);
			dAll = grown; capacity = bigger;
		}
		const uint32_t delta = (uint32_t)(count - cbq::kMaterialCount);
		CBQ_BRICK(cbq::launchAppendNodes(reinterpret_cast<const uint32_t*>(dMerged) + (size_t)cbq::kMaterialCount * 8, merged, delta,
			reinterpret_cast<uint32_t*>(dAll) + (size_t)count * 8, ctx->cfg.smCount, s));
		ctx->launches += 1;
		level[((size_t)bz * perAxis + by) * perAxis + bx] = (top >= cbq::kMaterialCount) ? top + delta : top;
		count += merged;
	}
	poolFree(ctx, dBrick); dBrick = nullptr;
	poolFree(ctx, dTree); dTree = nullptr;
	poolFree(ctx, dScratch); dScratch = nullptr;
	poolFree(ctx, dMerged); dMerged = nullptr;

	// The octree over the bricks, level by level until eight cubes are left, then their chains to the root.
	std::vector<uint32_t> words;
	uint32_t next = (uint32_t)count;                     // index the next host-made node gets
	for (uint32_t cells = perAxis; cells > 2; cells /= 2) {
		const uint32_t half = cells / 2;
		std::vector<uint32_t> above((size_t)half * half * half);
		for (uint32_t z = 0; z < half; z++) for (uint32_t y = 0; y < half; y++) for (uint32_t x = 0; x < half; x++) {
			for (uint32_t c = 0; c < 8; c++)
				words.push_back(level[((size_t)(2 * z + (c >> 2)) * cells + (2 * y + ((c >> 1) & 1u))) * cells + (2 * x + (c & 1u))]);
			above[((size_t)z * half + y) * half + x] = next++;
		}
		level.swap(above);
	}
	uint32_t cubes[8];
	for (uint32_t c = 0; c < 8; c++) cubes[c] = level[c];         // (z * 2 + y) * 2 + x, the order topTrie expects
	std::vector<uint32_t> trie;
	const uint32_t rootAt = cbq::topTrie(k, origin, cubes, next, trie);
	const uint32_t root = next + rootAt;
	words.insert(words.end(), trie.begin(), trie.end());
	const uint64_t total = count + words.size() / 8;
	if (total > capacity) {
		uint8_t* grown = nullptr;
		CBQ_BRICK(poolAlloc(ctx, &grown, (size_t)total * 32));
		CBQ_BRICK(cudaMemcpyAsync(grown, dAll, (size_t)count * 32, cudaMemcpyDeviceToDevice, s));
		poolFree(ctx, dAll);
		dAll = grown; capacity = total;
	}
	CBQ_BRICK(cudaMemcpyAsync(dAll + (size_t)count * 32, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
	CBQ_BRICK(cudaStreamSynchronize(s));                          // `words` is ours
#undef CBQ_BRICK
	const int rc = bakeAndInstall(ctx, reinterpret_cast<const uint32_t*>(dAll), total, root, "cbq_build_dense");
	poolFree(ctx, dAll);
	return rc;
}

} // namespace

int cbq_build_dense_device(cbq_context* ctx, const uint8_t* d_voxels, uint32_t size_log2, const int32_t origin[3],
	const float* colours_rgb, uint64_t* node_count, uint32_t* root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_voxels || !origin) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	if (size_log2 < 2 || size_log2 > 12) return fail(CBQ_ERROR_INVALID_ARGUMENT, "size_log2 %u out of range (2..12: 4^3 .. 4096^3 voxels)", size_log2);
	const uint32_t side = 1u << size_log2;
	for (int a = 0; a < 3; a++) {
		if (((uint32_t)origin[a] & (side / 2 - 1u)) != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "origin[%d] = %d is not a multiple of half the grid side (%u)", a, origin[a], side / 2);
		if ((int64_t)origin[a] + (int64_t)side > 0x80000000ll) return fail(CBQ_ERROR_INVALID_ARGUMENT, "the grid leaves the volume along axis %d", a);
	}
	const bool hadVolume = ctx->volume != nullptr;
	// In one piece up to 1024^3; in bricks of 512^3 beyond, or of 2^dense_brick_log2 when that option asks for it.
	const uint32_t brickLog2 = ctx->denseBrickLog2 ? (uint32_t)ctx->denseBrickLog2 : 9u;
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	if (size_log2 > 10 || (ctx->denseBrickLog2 && brickLog2 < size_log2)) {
		rc = buildDenseBricked(ctx, d_voxels, size_log2, origin, brickLog2);
		if (rc) return rc;
	} else {
		uint64_t n = cbq::denseNodeCount(size_log2);
		uint8_t* tree = nullptr;
		CBQ_CUDA(poolAlloc(ctx, &tree, (size_t)n * 32));
		uint32_t root = 0;
		cudaError_t e = cbq::launchBuildDense(d_voxels, size_log2, origin, reinterpret_cast<uint32_t*>(tree), &root, &n, ctx->cfg.smCount, ctx->stream, &ctx->launches);
		if (e != cudaSuccess) { poolFree(ctx, tree); return fail(CBQ_ERROR_CUDA, "cbq_build_dense: %s", cudaGetErrorString(e)); }
		rc = bakeAndInstall(ctx, reinterpret_cast<const uint32_t*>(tree), n, root, "cbq_build_dense");
		poolFree(ctx, tree);
		if (rc) return rc;
	}
	if (colours_rgb || !hadVolume) { rc = cbq_set_colours(ctx, colours_rgb); if (rc) return rc; }
	if (node_count) *node_count = ctx->nodeCount;
	if (root_index) *root_index = ctx->root;
	return CBQ_OK;
}

int cbq_build_dense(cbq_context* ctx, const uint8_t* voxels, uint32_t size_log2, const int32_t origin[3],
	const float* colours_rgb, uint64_t* node_count, uint32_t* root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!voxels) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	if (size_log2 < 2 || size_log2 > 12) return fail(CBQ_ERROR_INVALID_ARGUMENT, "size_log2 %u out of range (2..12)", size_log2);
	const size_t bytes = (size_t)1 << (3 * size_log2);
	uint8_t* d = nullptr;
	CBQ_CUDA(poolAlloc(ctx, &d, bytes));
	cudaError_t e = cudaMemcpyAsync(d, voxels, bytes, cudaMemcpyHostToDevice, ctx->stream);
	if (e != cudaSuccess) { poolFree(ctx, d); return fail(CBQ_ERROR_CUDA, "cbq_build_dense: %s", cudaGetErrorString(e)); }
	ctx->bytesH2D += bytes;
	rc = cbq_build_dense_device(ctx, d, size_log2, origin, colours_rgb, node_count, root_index);
	poolFree(ctx, d);
	return rc;
}

int cbq_voxelize(cbq_context* ctx, const float* triangles, const uint8_t* materials, uint64_t triangle_count,
	uint8_t fill, uint8_t background, int thin, uint32_t size_log2, const int32_t origin[3],
	const float* colours_rgb, cbq_mesh_info* info_out, uint64_t* node_count, uint32_t* root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!triangles || !materials || !origin || triangle_count == 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null or empty mesh");
	if (size_log2 < 2 || size_log2 > 11) return fail(CBQ_ERROR_INVALID_ARGUMENT, "size_log2 %u out of range (2..11: 4^3 .. 2048^3 voxels; the work space is 5 bytes per voxel)", size_log2);
	if (background != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "background must be material 0: everything outside the grid is empty");
	if (triangle_count > 0x7fffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "too many triangles");
	const uint32_t S = 1u << size_log2;
	bool aligned = true;
	for (int a = 0; a < 3; a++) {
		if (((uint32_t)origin[a] & (S / 2 - 1u)) != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "origin[%d] = %d is not a multiple of half the grid side (%u)", a, origin[a], S / 2);
		if ((int64_t)origin[a] + (int64_t)S > 0x80000000ll) return fail(CBQ_ERROR_INVALID_ARGUMENT, "the grid leaves the volume along axis %d", a);
		aligned = aligned && (((uint32_t)origin[a] & (S - 1u)) == 0);
	}
	const cbq::Tri* tris = reinterpret_cast<const cbq::Tri*>(triangles);
	cbq_mesh_info info;
	cbq::analyseMesh(tris, triangle_count, &info);            // Mesh::build (voxelization.cpp:765-823)
	if (info_out) *info_out = info;
	for (int a = 0; a < 3; a++) {
		if (!(info.lower[a] - 2.0f >= (float)origin[a]) || !(info.upper[a] + 2.0f <= (float)((int64_t)origin[a] + S - 1)))
			return fail(CBQ_ERROR_INVALID_ARGUMENT, "the mesh (dilated by 2 voxels) does not fit the grid along axis %d: [%g, %g] vs [%d, %lld]", a,
				info.lower[a], info.upper[a], origin[a], (long long)origin[a] + S - 1);
	}
	if (info.is_inside_out) return fail(CBQ_ERROR_INVALID_ARGUMENT, "the mesh is inside-out (exclusively negative winding numbers): flip its triangles");
	const bool solid = info.is_closed && fill != background;

	// drawLargeTriangle's subdivision, on the host: the pieces are what gets scan-converted, in the user's order.
	std::vector<cbq::Tri> pieces;
	std::vector<uint8_t> pieceMaterials;
	cbq::splitTriangles(tris, materials, triangle_count, pieces, pieceMaterials);
	if (pieces.size() > 0x7ffffff0ull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "too many triangle pieces");

	const size_t voxels = (size_t)1 << (3 * size_log2);
	size_t pyramid = 0;
	for (uint32_t l = 1; l <= size_log2; l++) pyramid += (size_t)1 << (3 * (size_log2 - l));
	uint8_t *dVoxels = nullptr, *dOrder = nullptr, *dPyramid = nullptr, *dPieces = nullptr, *dTris = nullptr, *dMisc = nullptr, *dLeaves = nullptr, *dInside = nullptr;
	cudaStream_t s = ctx->stream;
	CBQ_CUDA(cudaStreamSynchronize(s));
	auto cleanup = [&]() { poolFree(ctx, dVoxels); poolFree(ctx, dOrder); poolFree(ctx, dPyramid); poolFree(ctx, dPieces); poolFree(ctx, dTris); poolFree(ctx, dMisc); poolFree(ctx, dLeaves); poolFree(ctx, dInside); };
#define CBQ_VOX(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return fail(e_ == cudaErrorMemoryAllocation ? CBQ_ERROR_OUT_OF_MEMORY : CBQ_ERROR_CUDA, "cbq_voxelize: %s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
	CBQ_VOX(poolAlloc(ctx, &dVoxels, voxels));
	CBQ_VOX(poolAlloc(ctx, &dOrder, voxels * sizeof(unsigned int)));
	CBQ_VOX(poolAlloc(ctx, &dPieces, pieces.size() * sizeof(cbq::Tri) + pieces.size()));
	CBQ_VOX(poolAlloc(ctx, &dMisc, 256));
	uint8_t* dPieceMaterials = dPieces + pieces.size() * sizeof(cbq::Tri);
	CBQ_VOX(cudaMemsetAsync(dVoxels, background, voxels, s));
	CBQ_VOX(cudaMemsetAsync(dOrder, 0, voxels * sizeof(unsigned int), s));
	CBQ_VOX(cudaMemcpyAsync(dPieces, pieces.data(), pieces.size() * sizeof(cbq::Tri), cudaMemcpyHostToDevice, s));
	CBQ_VOX(cudaMemcpyAsync(dPieceMaterials, pieceMaterials.data(), pieces.size(), cudaMemcpyHostToDevice, s));
	ctx->bytesH2D += pieces.size() * (sizeof(cbq::Tri) + 1);
	const int sm = ctx->cfg.smCount;
	uint64_t leafCount = 0;
	if (!solid) {
		// the shell only, in the user's order (voxelization.cpp:738-743)
		CBQ_VOX(cbq::launchShell(dPieces, (uint32_t)pieces.size(), dVoxels, size_log2, origin, 1, background, thin, reinterpret_cast<unsigned int*>(dOrder), sm, s));
		CBQ_VOX(cbq::launchResolve(dVoxels, reinterpret_cast<unsigned int*>(dOrder), dPieceMaterials, voxels, sm, s));
		ctx->launches += 2;
	} else {
		// (a) the checkerboard shell
		CBQ_VOX(cbq::launchShell(dPieces, (uint32_t)pieces.size(), dVoxels, size_log2, origin, 0, background, thin, reinterpret_cast<unsigned int*>(dOrder), sm, s));
		// (b) occupancy pyramid + bounds, leaves, classification, fill
		CBQ_VOX(poolAlloc(ctx, &dPyramid, pyramid));
		CBQ_VOX(poolAlloc(ctx, &dTris, triangle_count * sizeof(cbq::Tri)));
		CBQ_VOX(cudaMemcpyAsync(dTris, tris, triangle_count * sizeof(cbq::Tri), cudaMemcpyHostToDevice, s));
		ctx->bytesH2D += triangle_count * sizeof(cbq::Tri);
		int* dBounds = reinterpret_cast<int*>(dMisc);
		unsigned long long* dCounter = reinterpret_cast<unsigned long long*>(dMisc + 64);
		const int boundsInit[6] = { 0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000 };
		CBQ_VOX(cudaMemcpyAsync(dBounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, s));
		CBQ_VOX(cudaMemsetAsync(dCounter, 0, 8, s));
		std::vector<uint8_t*> level(size_log2 + 1, nullptr);      // level[l]: occupancy of the cells of 2^l voxels
		{
			size_t at = 0;
			for (uint32_t l = 1; l <= size_log2; l++) { level[l] = dPyramid + at; at += (size_t)1 << (3 * (size_log2 - l)); }
		}
		CBQ_VOX(cbq::launchOccupancy(dVoxels, size_log2, background, 1, level[1], dBounds, sm, s));
		for (uint32_t l = 2; l <= size_log2; l++) CBQ_VOX(cbq::launchOccupancy(level[l - 1], size_log2 - (l - 1), background, 0, level[l], dBounds, sm, s));
		ctx->launches += 1 + size_log2;
		// Leaves per level: level 0 (voxels of occupied 2x2x2 blocks) ... level size_log2 - 1 (the grid's eight half-side cubes, whose
		// octree parent is the whole grid only if the grid is aligned to its own size).
		auto collect = [&](void* out) -> cudaError_t {
			cudaError_t e = cudaMemsetAsync(dCounter, 0, 8, s);
			for (uint32_t l = 0; l < size_log2 && e == cudaSuccess; l++) {
				const uint8_t* parent = (l + 1 == size_log2 && !aligned) ? nullptr : level[l + 1];
				if (l == 0 && size_log2 == 1) parent = nullptr;
				e = cbq::launchCollect(l == 0 ? nullptr : level[l], l == 0 ? level[1] : parent, size_log2, l, dBounds, dCounter, out, sm, s);
			}
			return e;
		};
		CBQ_VOX(collect(nullptr));
		CBQ_VOX(cudaMemcpyAsync(&leafCount, dCounter, 8, cudaMemcpyDeviceToHost, s));
		CBQ_VOX(cudaStreamSynchronize(s));
		ctx->launches += 2 * size_log2;
		if (leafCount) {
			CBQ_VOX(poolAlloc(ctx, &dLeaves, leafCount * 16));
			CBQ_VOX(poolAlloc(ctx, &dInside, leafCount));
			CBQ_VOX(collect(dLeaves));
			CBQ_VOX(cbq::launchClassify(dLeaves, leafCount, dTris, (uint32_t)triangle_count, origin, dInside, s));
			CBQ_VOX(cbq::launchFill(dLeaves, dInside, leafCount, dVoxels, size_log2, fill, background, sm, s));
			ctx->launches += 2;
		}
		// (c) surface materials: the last triangle within distance 1 of every voxel that is not background (or of every voxel, thin)
		CBQ_VOX(cbq::launchShell(dPieces, (uint32_t)pieces.size(), dVoxels, size_log2, origin, 2, background, thin, reinterpret_cast<unsigned int*>(dOrder), sm, s));
		CBQ_VOX(cbq::launchResolve(dVoxels, reinterpret_cast<unsigned int*>(dOrder), dPieceMaterials, voxels, sm, s));
		ctx->launches += 3;
	}
	CBQ_VOX(cudaStreamSynchronize(s));
#undef CBQ_VOX
	ctx->voxelizeLeaves = leafCount;
	ctx->voxelizePieces = pieces.size();
	rc = cbq_build_dense_device(ctx, dVoxels, size_log2, origin, colours_rgb, node_count, root_index);
	cleanup();
	return rc;
}

int cbq_fill_sphere(cbq_context* ctx, float x, float y, float z, float radius, uint8_t material, uint32_t* root_index, uint64_t* node_count)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "cbq_fill_sphere before cbq_upload");
	if (!(radius >= 0.0f)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad radius");
	CBQ_CUDA(cudaDeviceSynchronize());   // header, sub-DAGs and root references change in place
	uint32_t listCapacity = 1u << 16;
	for (int attempt = 0; attempt < 8; attempt++) {
		if (ctx->nodeCapacity < ctx->nodeCount + 4096) { rc = growVolume(ctx, ctx->nodeCount + std::max<uint64_t>(ctx->nodeCount / 4, 1u << 16)); if (rc) return rc; }
		const size_t listBytes = (size_t)listCapacity * 32;
		uint8_t* work = nullptr;
		CBQ_CUDA(poolAlloc(ctx, &work, listBytes + 1024));
		unsigned int* dState = reinterpret_cast<unsigned int*>(work + listBytes);
		unsigned int state[8] = { (unsigned int)ctx->nodeCount, 0, 0, 0, 0, 0, 0, 0 };
		cudaError_t e = cudaMemsetAsync(dState, 0, cbq::fillSphereStateBytes(), ctx->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dState, state, sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream);
		uint32_t* nodes = reinterpret_cast<uint32_t*>(ctx->volume + cbq::kNodeOffset);
		const uint32_t capacity = (uint32_t)std::min<uint64_t>(ctx->nodeCapacity, 0xfffffff0ull);
		if (e == cudaSuccess) e = cbq::launchFillSphere(nodes, capacity, ctx->root, x, y, z, radius, material, work, listCapacity, dState,
			ctx->cfg.smCount, ctx->stream, &ctx->launches);
		if (e == cudaSuccess) e = cudaMemcpyAsync(state, dState, sizeof(state), cudaMemcpyDeviceToHost, ctx->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) { poolFree(ctx, work); return fail(CBQ_ERROR_CUDA, "cbq_fill_sphere failed: %s", cudaGetErrorString(e)); }
		ctx->bytesD2H += sizeof(state);
		if (state[2] != 0) {
			// Nothing that existed was written: forget the attempt, make room, start over from the same root.
			poolFree(ctx, work);
			if (state[2] & 2u) { if (listCapacity >= (1u << 25)) return fail(CBQ_ERROR_OUT_OF_MEMORY, "cbq_fill_sphere: the brush touches too many nodes"); listCapacity <<= 3; }
			if (state[2] & 1u) {
				if (ctx->nodeCapacity >= 0xfffffff0ull) return fail(CBQ_ERROR_OUT_OF_MEMORY, "cbq_fill_sphere: node indices are 32-bit");
				rc = growVolume(ctx, std::min<uint64_t>(ctx->nodeCapacity * 2, 0xfffffff0ull)); if (rc) return rc;
			}
			continue;
		}
		const uint64_t oldCount = ctx->nodeCount;
		const uint32_t oldRoot = ctx->root;
		ctx->nodeCount = state[0];
		ctx->root = state[4];
		rc = refreshSubdagsOnDevice(ctx, ctx->root, nullptr, work, "cbq_fill_sphere");
		poolFree(ctx, work);
		if (rc) { ctx->nodeCount = oldCount; ctx->root = oldRoot; return rc; }
		ctx->generation++;
		ctx->deviceDiverged = true;
		if (root_index) *root_index = ctx->root;
		if (node_count) *node_count = ctx->nodeCount;
		return writeHeaderAndSubdags(ctx);
	}
	return fail(CBQ_ERROR_OUT_OF_MEMORY, "cbq_fill_sphere: could not make room for the edit");
}

int cbq_set_root(cbq_context* ctx, uint32_t root_index)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "cbq_set_root before cbq_upload");
	if (root_index >= ctx->nodeCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "root %u is past the %llu nodes on the device", root_index, (unsigned long long)ctx->nodeCount);
	CBQ_CUDA(cudaDeviceSynchronize());   // sub-DAGs and root references change in place
	uint8_t* work = nullptr;
	CBQ_CUDA(poolAlloc(ctx, &work, 1024));
	const uint32_t oldRoot = ctx->root;
	ctx->root = root_index;
	rc = refreshSubdagsOnDevice(ctx, root_index, nullptr, work, "cbq_set_root");
	poolFree(ctx, work);
	if (rc) { ctx->root = oldRoot; return rc; }
	ctx->generation++;
	return writeHeaderAndSubdags(ctx);
}

int cbq_set_colours(cbq_context* ctx, const float* colours_rgb)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	std::vector<float> c4(256 * 4);
	for (int i = 0; i < 256; i++) {
		// Purple default, like Viewer (reference viewer.cpp:50-55).
		c4[4 * i + 0] = colours_rgb ? colours_rgb[3 * i + 0] : 1.0f;
		c4[4 * i + 1] = colours_rgb ? colours_rgb[3 * i + 1] : 0.0f;
		c4[4 * i + 2] = colours_rgb ? colours_rgb[3 * i + 2] : 1.0f;
		c4[4 * i + 3] = 0.0f;
	}
	CBQ_CUDA(cudaMemcpyAsync(ctx->volume + cbq::kColourOffset, c4.data(), c4.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->bytesH2D += c4.size() * sizeof(float);
	return CBQ_OK;
}

int cbq_get_subdags(cbq_context* ctx, cbq_subdag out[8])
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	// Read back what the kernels see, not the host copy.
	CBQ_CUDA(cudaMemcpyAsync(out, ctx->volume + cbq::kSubDagOffset, 256, cudaMemcpyDeviceToHost, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	return CBQ_OK;
}

int cbq_download_nodes(cbq_context* ctx, uint64_t begin, uint64_t count, uint32_t* out)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	if (begin + count > ctx->nodeCount) return fail(CBQ_ERROR_INVALID_ARGUMENT, "range past end");
	CBQ_CUDA(cudaMemcpyAsync(out, ctx->volume + cbq::kNodeOffset + (size_t)begin * 32, (size_t)count * 32, cudaMemcpyDeviceToHost, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	return CBQ_OK;
}

int cbq_node_count(cbq_context* ctx, uint64_t* out)
{
	if (!ctx || !out) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	*out = ctx->nodeCount;
	return CBQ_OK;
}

int cbq_trace_device(cbq_context* ctx, const cbq_ray* d_rays, uint64_t n, uint32_t flags, float max_footprint, cbq_hit* d_hits, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (n && (!d_rays || !d_hits)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null ray or hit buffer");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	return traceDevice(ctx, reinterpret_cast<const cbq::Ray*>(d_rays), n, flags, max_footprint, reinterpret_cast<cbq::Hit*>(d_hits), s, nullptr, 0, 0);
}

namespace {

// Host rays in, host results out: a three-stage pipeline over chunks -- H2D on copyIn, kernel on stream, D2H on copyOut,
// kStages staging buffers in flight. With pinned host memory the three overlap; with pageable memory the copies
// degrade to synchronous but the result is the same. The copy back is the longest leg for 40-byte records, so the first
// stages are short -- 2^15, 2^16, 2^17 rays -- to get it going early; after that every stage is `chunk` rays.
// recordBytes = 40 (cbq_hit) or 8 (cbq_hit_compact).
int tracePipeline(cbq_context* ctx, const cbq_ray* rays, uint64_t n, uint32_t flags, float max_footprint, void* results, size_t recordBytes)
{
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	if (n == 0) return CBQ_OK;
	if (!rays || !results) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null ray or hit buffer");
	const uint64_t chunk = std::min(kPipelineChunkMax, std::max(kPipelineChunkMin, n / 8));
	int rc = ensureStaging(ctx, chunk); if (rc) return rc;
	const bool compact = recordBytes == sizeof(cbq_hit_compact);
	uint8_t* out = static_cast<uint8_t*>(results);
	uint64_t begin = 0, stage = std::min<uint64_t>(chunk, 1u << 15);
	for (uint64_t c = 0; begin < n; c++) {
		const int b = (int)(c % kStages);
		const uint64_t len = std::min(stage, n - begin);
		if (c >= (uint64_t)kStages) CBQ_CUDA(cudaStreamWaitEvent(ctx->copyIn, ctx->evOut[b], 0));   // buffer b free again
		CBQ_CUDA(cudaMemcpyAsync(ctx->stageRays[b], rays + begin, len * sizeof(cbq_ray), cudaMemcpyHostToDevice, ctx->copyIn));
		CBQ_CUDA(cudaEventRecord(ctx->evIn[b], ctx->copyIn));
		CBQ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evIn[b], 0));
		if (c >= (uint64_t)kStages) CBQ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evOut[b], 0));
		rc = traceDevice(ctx, ctx->stageRays[b], len, flags, max_footprint, compact ? nullptr : ctx->stageHits[b], ctx->stream, nullptr, 0, 0,
			compact ? reinterpret_cast<cbq_hit_compact*>(ctx->stageHits[b]) : nullptr);
		if (rc) return rc;
		CBQ_CUDA(cudaEventRecord(ctx->evKernel[b], ctx->stream));
		CBQ_CUDA(cudaStreamWaitEvent(ctx->copyOut, ctx->evKernel[b], 0));
		CBQ_CUDA(cudaMemcpyAsync(out + begin * recordBytes, ctx->stageHits[b], len * recordBytes, cudaMemcpyDeviceToHost, ctx->copyOut));
		CBQ_CUDA(cudaEventRecord(ctx->evOut[b], ctx->copyOut));
		begin += len;
		stage = std::min<uint64_t>(stage * 2, chunk);
	}
	CBQ_CUDA(cudaStreamSynchronize(ctx->copyOut));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->bytesH2D += n * sizeof(cbq_ray);
	ctx->bytesD2H += n * recordBytes;
	return CBQ_OK;
}

} // namespace

int cbq_trace(cbq_context* ctx, const cbq_ray* rays, uint64_t n, uint32_t flags, float max_footprint, cbq_hit* hits)
{
	int rc = bind(ctx); if (rc) return rc;
	return tracePipeline(ctx, rays, n, flags, max_footprint, hits, sizeof(cbq_hit));
}

int cbq_trace_compact(cbq_context* ctx, const cbq_ray* rays, uint64_t n, uint32_t flags, float max_footprint, cbq_hit_compact* hits)
{
	int rc = bind(ctx); if (rc) return rc;
	return tracePipeline(ctx, rays, n, flags, max_footprint, hits, sizeof(cbq_hit_compact));
}

int cbq_trace_compact_device(cbq_context* ctx, const cbq_ray* d_rays, uint64_t n, uint32_t flags, float max_footprint, cbq_hit_compact* d_hits, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (n && (!d_rays || !d_hits)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null ray or hit buffer");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	return traceDevice(ctx, reinterpret_cast<const cbq::Ray*>(d_rays), n, flags, max_footprint, nullptr, s, nullptr, 0, 0, d_hits);
}

int cbq_camera_from_pose(const double position[3], double pitch, double yaw, double fov_degrees, cbq_camera* c)
{
	if (!position || !c) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	// Camera::forward / right / up and the fov scale (reference camera.cpp:24,40-66), evaluated with
	// the host libm exactly as the reference does; Pi is the FLOAT constant of camera.h:6.
	const float Pi = 3.14159265358979f;
	std::memset(c, 0, sizeof(*c));
	for (int a = 0; a < 3; a++) c->position[a] = position[a];
	c->forward[0] = std::cos(pitch) * std::sin(yaw);
	c->forward[1] = std::cos(pitch) * std::cos(yaw);
	c->forward[2] = std::sin(pitch);
	c->right[0] = std::sin(yaw + (Pi / 2));
	c->right[1] = std::cos(yaw + (Pi / 2));
	c->right[2] = 0;
	c->up[0] = c->right[1] * c->forward[2] - c->right[2] * c->forward[1];
	c->up[1] = c->right[2] * c->forward[0] - c->right[0] * c->forward[2];
	c->up[2] = c->right[0] * c->forward[1] - c->right[1] * c->forward[0];
	c->scale = (float)(std::tan(fov_degrees * 0.0174533f * 0.5f) * 2.0f);
	return CBQ_OK;
}

int cbq_primary_rays_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height, cbq_ray* d_rays, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!cam || !d_rays || !width || !height) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	CBQ_CUDA(cbq::launchPrimaryRays(*cam, width, height, reinterpret_cast<cbq::Ray*>(d_rays), s));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_primary_rays_tiled_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height, cbq_ray* d_rays,
	uint32_t* d_pixel_of, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!cam || !d_rays || !width || !height) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	if ((width % 8u) != 0 || (height % 4u) != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "tiled ray order needs width %% 8 == 0 and height %% 4 == 0");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	CBQ_CUDA(cbq::launchPrimaryRays(*cam, width, height, reinterpret_cast<cbq::Ray*>(d_rays), s, 1, d_pixel_of));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_random_rays_device(cbq_context* ctx, uint64_t seed, const float lower[3], const float upper[3], uint64_t n, cbq_ray* d_rays, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!lower || !upper || (n && !d_rays)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	if (n == 0) return CBQ_OK;
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	CBQ_CUDA(cbq::launchRandomRays(seed, lower, upper, n, reinterpret_cast<cbq::Ray*>(d_rays), s));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_raycast_frame_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height, uint32_t flags,
	float max_footprint, cbq_hit* d_hits, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!cam || !d_hits || !width || !height) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	if ((width % 8u) == 0 && (height % 4u) == 0) {
		// Generate the frame's rays once, in 8x4-pixel tile order, and trace them from the buffer: 7-17 % faster
		// than the row-major order and 20 % faster than generating rays inside the trace kernel (which pays for
		// the camera's double-precision maths twice per ray and for its registers): profiles/r01_analysis.md.
		if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
		const size_t n = (size_t)width * height;
		if (n > ctx->frameRayCapacity) {
			CBQ_CUDA(cudaDeviceSynchronize());
			cudaFree(ctx->frameRays); ctx->frameRays = nullptr; ctx->frameRayCapacity = 0;
			CBQ_CUDA(cudaMalloc(&ctx->frameRays, n * sizeof(cbq::Ray)));
			ctx->frameRayCapacity = n;
		}
		CBQ_CUDA(cbq::launchPrimaryRays(*cam, width, height, ctx->frameRays, s, 1, nullptr));
		ctx->launches++;
		cbq::TraceArgs a;
		std::memset(&a, 0, sizeof(a));
		a.volume = ctx->view();
		a.rays = ctx->frameRays; a.hits = reinterpret_cast<cbq::Hit*>(d_hits); a.count = n; a.maxFootprint = max_footprint;
		a.abandoned = ctx->abandonedPtr(); a.untileWidth = width;
		rc = nextQueue(ctx, s, &a.queue); if (rc) return rc;
		applyL2Window(ctx, s);
		cbq::LaunchConfig cfg = ctx->cfg;
		cfg.refillThreshold = 32;
		const bool feedback = orderBeforeTrace(ctx, a, cfg, s);
		CBQ_CUDA(cbq::launchTrace(a, (flags & CBQ_TRACE_SURFACE) != 0, cfg, s));
		ctx->launches++;
		ctx->raysTraced += n;
		const bool moved = std::memcmp(&ctx->lastFrameCamera, cam, sizeof(cbq_camera)) != 0;
		ctx->lastFrameCamera = *cam;
		if (feedback) return orderAfterTrace(ctx, a, s, moved);
		return CBQ_OK;
	}
	return traceDevice(ctx, nullptr, (uint64_t)width * height, flags, max_footprint, reinterpret_cast<cbq::Hit*>(d_hits), s, cam, width, height);
}

int cbq_render_device(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, float* d_accum, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!ctx->volume) return fail(CBQ_ERROR_NO_VOLUME, "no volume uploaded");
	if (!cam || !p || !d_accum) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	if (!p->width || !p->height || p->x1 > p->width || p->y1 > p->height || p->x0 > p->x1 || p->y0 > p->y1)
		return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad image rectangle");
	if (p->variant > 1 || p->bounces > 5) return fail(CBQ_ERROR_INVALID_ARGUMENT, "variant must be 0/1 and bounces <= 5 (the viewer's F2 limit, pathtracing_demo.cpp:280)");
	if (p->band_count > 1 && p->band_index >= p->band_count) return fail(CBQ_ERROR_INVALID_ARGUMENT, "band_index must be < band_count");
	if (p->tile_group_count > 1) {
		if (p->tile_group_index >= p->tile_group_count) return fail(CBQ_ERROR_INVALID_ARGUMENT, "tile_group_index must be < tile_group_count");
		if (p->band_count > 1 || p->x0 != 0 || p->y0 != 0 || p->x1 != p->width || p->y1 != p->height)
			return fail(CBQ_ERROR_INVALID_ARGUMENT, "tile groups need the whole image as the rectangle and no bands");
	}
	if (p->x0 == p->x1 || p->y0 == p->y1 || p->spp == 0) return CBQ_OK;
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	cbq::RenderArgs a;
	std::memset(&a, 0, sizeof(a));
	a.volume = ctx->view(); a.colours = ctx->coloursPtr();
	a.camera = *cam; a.params = *p; a.accum = d_accum; a.abandoned = ctx->abandonedPtr();
	applyL2Window(ctx, s);
	size_t pixels = (size_t)(p->x1 - p->x0) * cbq::bandedRowCount(p->y1 - p->y0, p->band_count, p->band_index);
	if (p->tile_group_count > 1) {
		// The tiles of this group, in row-major tile order (glsl/pathtracing.frag:792-799: group = bitMix(tile id) % count).
		std::vector<uint32_t> tiles;
		const uint32_t tilesX = (p->width + 63u) / 64u, tilesY = (p->height + 63u) / 64u;
		for (uint32_t ty = 0; ty < tilesY; ty++) for (uint32_t tx = 0; tx < tilesX; tx++)
			if (fmix32Host((tx << 16) | (ty & 0xffffu)) % p->tile_group_count == p->tile_group_index) tiles.push_back(tx | (ty << 16));
		if (tiles.empty()) return CBQ_OK;
		if (tiles.size() > ctx->tileListCapacity) {
			CBQ_CUDA(cudaDeviceSynchronize());
			cudaFree(ctx->tileList); ctx->tileList = nullptr; ctx->tileListCapacity = 0;
			const size_t cap = std::max<size_t>(1024, tiles.size() * 2);
			CBQ_CUDA(cudaMalloc(&ctx->tileList, cap * sizeof(uint32_t)));
			ctx->tileListCapacity = cap;
		}
		CBQ_CUDA(cudaStreamSynchronize(s));     // the previous call's kernels may still be reading the list
		CBQ_CUDA(cudaMemcpyAsync(ctx->tileList, tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
		CBQ_CUDA(cudaStreamSynchronize(s));     // `tiles` is on our stack
		a.tiles = ctx->tileList; a.tileCount = (uint32_t)tiles.size();
		pixels = tiles.size() * 4096;
	}
	if (pixels == 0) return CBQ_OK;
	// Samples traced together. 0 = auto: as many as fit ~32 M paths per wave (16 samples of a 1080p frame, ~12 GB of
	// path buffers), between 1 and 16: what the samples share (primary ray, depth-0 sun ray) is traced once per wave.
	cbq::LaunchConfig renderCfg = ctx->cfg;
	if (ctx->cfg.sampleGroup <= 0) renderCfg.sampleGroup = (int)std::max<size_t>(1, std::min<size_t>(16, (32u << 20) / pixels));
	else renderCfg.sampleGroup = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->cfg.sampleGroup, (32u << 20) / pixels));
	const size_t paths = pixels * std::min<size_t>(p->spp, (size_t)renderCfg.sampleGroup);
	if (paths > 0xffffffffull) return fail(CBQ_ERROR_INVALID_ARGUMENT, "rectangle too large for 32-bit path ids");
	if (paths > ctx->wavefront.pathCapacity || pixels > ctx->wavefront.pixelCapacity) CBQ_CUDA(cudaDeviceSynchronize());   // buffers may still be in use
	CBQ_CUDA((cudaError_t)cbq::wavefrontReserve(ctx->wavefront, paths, pixels));
	struct Adaptor { static int next(void* user, cudaStream_t st, unsigned long long** out) { return nextQueue(static_cast<cbq_context*>(user), st, out); } };
	CBQ_CUDA(cbq::launchRenderWavefront(a, ctx->wavefront, renderCfg, s, &Adaptor::next, ctx, &ctx->launches));
	return CBQ_OK;
}

int cbq_render(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, float* accum)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!cam || !p || !accum) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	const size_t bytes = (size_t)p->width * p->height * 3 * sizeof(float);
	if (bytes > ctx->stageAccumBytes) {
		CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->stageAccum); ctx->stageAccum = nullptr; ctx->stageAccumBytes = 0;
		CBQ_CUDA(cudaMalloc(&ctx->stageAccum, bytes));
		ctx->stageAccumBytes = bytes;
	}
	// accum is added to, so the caller's running image goes up first (mImage += pixel).
	CBQ_CUDA(cudaMemcpyAsync(ctx->stageAccum, accum, bytes, cudaMemcpyHostToDevice, ctx->stream));
	rc = cbq_render_device(ctx, cam, p, ctx->stageAccum, ctx->stream); if (rc) return rc;
	CBQ_CUDA(cudaMemcpyAsync(accum, ctx->stageAccum, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CBQ_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->bytesH2D += bytes; ctx->bytesD2H += bytes;
	return CBQ_OK;
}

int cbq_progressive_pass_device(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, uint32_t frame, float* d_rgba, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!cam || !p || !d_rgba || !p->width || !p->height) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	const size_t pixels = (size_t)p->width * p->height;
	if (pixels > ctx->progressiveScratchPixels) {
		CBQ_CUDA(cudaDeviceSynchronize());
		cudaFree(ctx->progressiveScratch); ctx->progressiveScratch = nullptr; ctx->progressiveScratchPixels = 0;
		CBQ_CUDA(cudaMalloc(&ctx->progressiveScratch, pixels * 3 * sizeof(float)));
		CBQ_CUDA(cudaMemset(ctx->progressiveScratch, 0, pixels * 3 * sizeof(float)));
		ctx->progressiveScratchPixels = pixels;
	}
	const uint32_t kGroups = 16;      // glsl/pathtracing.frag:796
	cbq_pt_params q = *p;
	q.x0 = 0; q.y0 = 0; q.x1 = p->width; q.y1 = p->height; q.band_count = 0; q.band_index = 0;
	q.tile_group_count = kGroups; q.tile_group_index = frame % kGroups;
	q.frame_id = frame * p->spp;
	rc = cbq_render_device(ctx, cam, &q, ctx->progressiveScratch, s); if (rc) return rc;
	CBQ_CUDA(cbq::launchProgressiveAdd(ctx->progressiveScratch, d_rgba, p->width, p->height, kGroups, q.tile_group_index, (float)p->spp, s));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_normalise_device(cbq_context* ctx, const float* d_rgba, uint32_t width, uint32_t height, float* d_rgb, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_rgba || !d_rgb || !width || !height) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	CBQ_CUDA(cbq::launchNormalise(d_rgba, width, height, d_rgb, stream ? (cudaStream_t)stream : ctx->stream));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_blur_device(cbq_context* ctx, float* d_rgba, uint32_t width, uint32_t height, float* d_scratch_rgba, int passes, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_rgba || !d_scratch_rgba || !width || !height || passes < 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	for (int i = 0; i < passes; i++) {
		CBQ_CUDA(cbq::launchBlur(d_rgba, d_scratch_rgba, width, height, 0, s));     // hBlurProgram: accTexture -> blurTexture
		CBQ_CUDA(cbq::launchBlur(d_scratch_rgba, d_rgba, width, height, 1, s));     // vBlurProgram: blurTexture -> accTexture
		ctx->launches += 2;
	}
	return CBQ_OK;
}

int cbq_upload_dag(cbq_context* ctx, const char* path, const float* colours_rgb)
{
	uint32_t* nodes = nullptr;
	uint64_t count = 0;
	uint32_t root = 0;
	int rc = cbq_dag_load(path, &nodes, &count, &root);
	if (rc) return fail(rc, "cbq_upload_dag: %s", cbq_dag_error());
	rc = cbq_upload(ctx, nodes, count, root, colours_rgb);
	cbq_dag_free(nodes);
	return rc;
}

void cbq_set_log_callback(void (*handler)(const char* message)) { g_logHandler = handler; }

int cbq_rng_points_device(cbq_context* ctx, const uint32_t* d_seeds, uint64_t n, int draws, float* d_points, uint32_t* d_states, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if ((n && (!d_seeds || !d_points || !d_states)) || draws < 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	if (n == 0 || draws == 0) return CBQ_OK;
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	CBQ_CUDA(cbq::launchRngPoints(d_seeds, n, draws, d_points, d_states, s));
	ctx->launches++;
	return CBQ_OK;
}

int cbq_shared_alloc(cbq_context* ctx, uint64_t bytes, void** d_ptr, cbq_ipc_handle* handle)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_ptr || !handle || bytes == 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	static_assert(sizeof(cbq_ipc_handle) == sizeof(cudaIpcMemHandle_t), "IPC handle size");
	void* p = nullptr;
	CBQ_CUDA(cudaMalloc(&p, bytes));            // its own allocation: an IPC handle exports a whole cudaMalloc block
	cudaError_t e = cudaMemset(p, 0, bytes);
	if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle), p);
	if (e != cudaSuccess) { cudaFree(p); return fail(CBQ_ERROR_CUDA, "cbq_shared_alloc: %s", cudaGetErrorString(e)); }
	*d_ptr = p;
	return CBQ_OK;
}

int cbq_shared_open(cbq_context* ctx, const cbq_ipc_handle* handle, uint64_t bytes, void** d_ptr)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_ptr || !handle || bytes == 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "bad argument");
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle, sizeof(h));
	CBQ_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	// Remember the range: a trace whose results land in it stores them a warp at a time (ParkedCompactSink).
	ctx->remote.push_back({ reinterpret_cast<uintptr_t>(*d_ptr), reinterpret_cast<uintptr_t>(*d_ptr) + bytes });
	return CBQ_OK;
}

int cbq_shared_close(cbq_context* ctx, void* d_ptr)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!d_ptr) return CBQ_OK;
	CBQ_CUDA(cudaDeviceSynchronize());
	for (size_t i = 0; i < ctx->remote.size(); i++) if (ctx->remote[i].first == reinterpret_cast<uintptr_t>(d_ptr)) { ctx->remote.erase(ctx->remote.begin() + i); break; }
	CBQ_CUDA(cudaIpcCloseMemHandle(d_ptr));
	return CBQ_OK;
}

int cbq_shared_free(cbq_context* ctx, void* d_ptr)
{
	int rc = bind(ctx); if (rc) return rc;
	if (d_ptr) { CBQ_CUDA(cudaDeviceSynchronize()); CBQ_CUDA(cudaFree(d_ptr)); }
	return CBQ_OK;
}

int cbq_copy_device(cbq_context* ctx, void* d_dst, const void* d_src, uint64_t bytes, void* stream)
{
	int rc = bind(ctx); if (rc) return rc;
	if (bytes == 0) return CBQ_OK;
	if (!d_dst || !d_src) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	CBQ_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, stream ? (cudaStream_t)stream : ctx->stream));
	return CBQ_OK;
}

int cbq_host_alloc(void** out, uint64_t bytes)
{
	if (!out) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	if (cbq_device_count() <= 0) return fail(CBQ_ERROR_NO_DEVICE, "no CUDA device is visible");
	CBQ_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
	return CBQ_OK;
}

int cbq_host_free(void* p)
{
	if (p) CBQ_CUDA(cudaFreeHost(p));
	return CBQ_OK;
}

int cbq_set_option(cbq_context* ctx, const char* key, int64_t value)
{
	if (!ctx || !key) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	const std::string k(key);
	if (k == "block_threads") {
		if (value < 32 || value > 256 || (value % 32) != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "block_threads must be a multiple of 32 in [32, 256]");
		ctx->cfg.blockThreads = (int)value;
	} else if (k == "blocks_per_sm") {
		if (value < 1 || value > 32) return fail(CBQ_ERROR_INVALID_ARGUMENT, "blocks_per_sm must be in [1, 32]");
		ctx->cfg.blocksPerSm = (int)value;
	} else if (k == "refill_threshold") {
		if (value < 1 || value > 32) return fail(CBQ_ERROR_INVALID_ARGUMENT, "refill_threshold must be in [1, 32]");
		ctx->cfg.refillThreshold = (int)value;
	} else if (k == "refill_quantum") {
		if (value < 1 || value > 32 || (value & (value - 1)) != 0) return fail(CBQ_ERROR_INVALID_ARGUMENT, "refill_quantum must be a power of two in [1, 32]");
		ctx->cfg.refillQuantum = (int)value;
	} else if (k == "pt_refill_threshold") {
		if (value < 1 || value > 32) return fail(CBQ_ERROR_INVALID_ARGUMENT, "pt_refill_threshold must be in [1, 32]");
		ctx->cfg.secondaryRefill = (int)value;
	} else if (k == "sample_group") {
		if (value < 0 || value > 16) return fail(CBQ_ERROR_INVALID_ARGUMENT, "sample_group must be 0 (auto) or in [1, 16]");
		ctx->cfg.sampleGroup = (int)value;
	} else if (k == "order_refresh") {
		if (value < 1 || value > 1024) return fail(CBQ_ERROR_INVALID_ARGUMENT, "order_refresh must be in [1, 1024]");
		ctx->orderRefresh = (int)value;
	} else if (k == "dense_brick_log2") {
		if (value != 0 && (value < 3 || value > 10)) return fail(CBQ_ERROR_INVALID_ARGUMENT, "dense_brick_log2 must be 0 (bricks of 512^3, past 1024^3 only) or in [3, 10]");
		ctx->denseBrickLog2 = (int)value;
	} else if (k == "park_results") {
		ctx->parkResults = value ? 1 : 0;
	} else if (k == "adaptive_order") {
		ctx->adaptiveOrder = value ? 1 : 0;
		ctx->orderTickets = 0;         // forget what was learnt
	} else if (k == "l2_persist") {
		ctx->l2Persist = value ? 1 : 0;
		ctx->windowUsed = 0;           // re-apply on next launch
	} else {
		return fail(CBQ_ERROR_INVALID_ARGUMENT, "unknown option '%s'", key);
	}
	return CBQ_OK;
}

int cbq_get_option(cbq_context* ctx, const char* key, int64_t* value)
{
	if (!ctx || !key || !value) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	const std::string k(key);
	if (k == "block_threads") *value = ctx->cfg.blockThreads;
	else if (k == "blocks_per_sm") *value = ctx->cfg.blocksPerSm;
	else if (k == "refill_threshold") *value = ctx->cfg.refillThreshold;
	else if (k == "refill_quantum") *value = ctx->cfg.refillQuantum;
	else if (k == "l2_persist") *value = ctx->l2Persist;
	else if (k == "adaptive_order") *value = ctx->adaptiveOrder;
	else if (k == "park_results") *value = ctx->parkResults;
	else if (k == "dense_brick_log2") *value = ctx->denseBrickLog2;
	else if (k == "order_refresh") *value = ctx->orderRefresh;
	else if (k == "sample_group") *value = ctx->cfg.sampleGroup;
	else if (k == "pt_refill_threshold") *value = ctx->cfg.secondaryRefill;
	else if (k == "sm_count") *value = ctx->cfg.smCount;
	else if (k == "stack_levels") *value = ctx->cfg.stackLevels;
	else if (k == "l2_bytes") *value = ctx->prop.l2CacheSize;
	else if (k == "l2_persist_max_bytes") *value = ctx->prop.persistingL2CacheMaxSize;
	else return fail(CBQ_ERROR_INVALID_ARGUMENT, "unknown option '%s'", key);
	return CBQ_OK;
}

int cbq_get_counter(cbq_context* ctx, const char* key, uint64_t* value)
{
	int rc = bind(ctx); if (rc) return rc;
	if (!key || !value) return fail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	const std::string k(key);
	if (k == "kernel_launches") *value = ctx->launches;
	else if (k == "rays_traced") *value = ctx->raysTraced;
	else if (k == "bytes_h2d") *value = ctx->bytesH2D;
	else if (k == "bytes_d2h") *value = ctx->bytesD2H;
	else if (k == "bake_reachable") *value = ctx->bakeReachable;
	else if (k == "voxelize_leaves") *value = ctx->voxelizeLeaves;
	else if (k == "voxelize_pieces") *value = ctx->voxelizePieces;
	else if (k == "abandoned_rays") {
		unsigned long long v = 0;
		CBQ_CUDA(cudaDeviceSynchronize());
		CBQ_CUDA(cudaMemcpy(&v, ctx->abandonedPtr(), sizeof(v), cudaMemcpyDeviceToHost));
		*value = v;
	} else return fail(CBQ_ERROR_INVALID_ARGUMENT, "unknown counter '%s'", key);
	return CBQ_OK;
}

int cbq_reset_counters(cbq_context* ctx)
{
	int rc = bind(ctx); if (rc) return rc;
	ctx->launches = ctx->raysTraced = ctx->bytesH2D = ctx->bytesD2H = 0;
	CBQ_CUDA(cudaDeviceSynchronize());
	CBQ_CUDA(cudaMemset(ctx->abandonedPtr(), 0, sizeof(unsigned long long)));
	return CBQ_OK;
}

} // extern "C"

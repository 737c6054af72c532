// Host-only half of the compact result path: cbq_expand_hits widens 8-byte cbq_hit_compact records into the
// 40-byte cbq_hit records cbq_trace writes -- every field of RayVolumeIntersection (reference
// src/library/raytracing.h:48-55). `position` is re-formed as the reference forms it (raytracing.cpp:463-466):
// origin + dir * (float)distance, one rounded multiply and one rounded add per component. This file is compiled
// with -ffp-contract=off so the compiler cannot fuse them, whatever the host ISA; the GPU kernels do the same
// arithmetic under -fmad=false, hence byte-identical records (tests/test_gpu_trace.py::test_compact_results_*).
#include "../../include/cubiquity_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

void expandRange(const cbq_ray* rays, const cbq_hit_compact* in, cbq_hit* out, uint64_t begin, uint64_t end)
{
	for (uint64_t i = begin; i < end; i++) {
		cbq_hit h;
		std::memset(&h, 0, sizeof(h));
		const uint32_t code = in[i].code;
		if (code & (1u << 14)) {
			h.hit = 1;
			h.distance = in[i].distance;
			h.material = code & 0xffu;
			for (int a = 0; a < 3; a++) {
				const uint32_t two = (code >> (8 + 2 * a)) & 3u;
				const uint32_t bits = ((two & 1u) ? 0x3f800000u : 0u) | ((two & 2u) ? 0x80000000u : 0u);
				std::memcpy(&h.normal[a], &bits, sizeof(bits));
				const float step = rays[i].dir[a] * h.distance;
				h.position[a] = rays[i].origin[a] + step;
			}
		}
		if (code & (1u << 15)) h.status = CBQ_HIT_ABANDONED;
		out[i] = h;
	}
}

} // namespace

extern "C" int cbq_expand_hits(const cbq_ray* rays, const cbq_hit_compact* compact, uint64_t n, cbq_hit* hits, int threads)
{
	if (n == 0) return CBQ_OK;
	if (!rays || !compact || !hits) return CBQ_ERROR_INVALID_ARGUMENT;
	unsigned t = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
	if (t == 0) t = 1;
	if (n < 65536 || t == 1) { expandRange(rays, compact, hits, 0, n); return CBQ_OK; }
	if (t > 64) t = 64;
	std::vector<std::thread> pool;
	const uint64_t per = (n + t - 1) / t;
	for (unsigned k = 0; k < t; k++) {
		const uint64_t b = (uint64_t)k * per, e = b + per < n ? b + per : n;
		if (b >= e) break;
		pool.emplace_back(expandRange, rays, compact, hits, b, e);
	}
	for (auto& th : pool) th.join();
	return CBQ_OK;
}

// ---- .dag files --------------------------------------------------------------------------------------------------------
// Volume::load / Volume::save and NodeStore::read / write (reference src/library/storage.cpp:505-542, 192-206): u32 root,
// u32 nodeCount, nodeCount x 32 bytes. The reference reads without checking anything; here every size is 64-bit and the
// file has to be exactly what its header says.

namespace {
thread_local std::string g_dagError;
int dagFail(int code, const std::string& why) { g_dagError = why; return code; }
}

extern "C" const char* cbq_dag_error(void) { return g_dagError.c_str(); }

extern "C" int cbq_dag_load(const char* path, uint32_t** nodes, uint64_t* node_count, uint32_t* root_index)
{
	if (!path || !nodes || !node_count || !root_index) return dagFail(CBQ_ERROR_INVALID_ARGUMENT, "null argument");
	*nodes = nullptr; *node_count = 0; *root_index = 0;
	FILE* f = std::fopen(path, "rb");
	if (!f) return dagFail(CBQ_ERROR_INVALID_ARGUMENT, std::string("cannot open ") + path);
	uint32_t head[2];
	if (std::fread(head, sizeof(uint32_t), 2, f) != 2) { std::fclose(f); return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": shorter than its 8-byte header"); }
	const uint64_t stored = head[1], total = stored + 256u;
	// the length must be exactly header + nodes (64-bit arithmetic: 2^27 nodes are already 4 GiB)
	if (std::fseek(f, 0, SEEK_END) != 0) { std::fclose(f); return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": cannot seek"); }
	const long long length = (long long)ftello(f);
	if (length < 0 || (uint64_t)length != 8u + stored * 32u) {
		std::fclose(f);
		return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": header says " + std::to_string(stored) + " nodes (" + std::to_string(8u + stored * 32u) +
			" bytes), the file has " + std::to_string(length));
	}
	if (head[0] >= total) { std::fclose(f); return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": root index " + std::to_string(head[0]) + " is past the " + std::to_string(total) + " nodes"); }
	std::fseek(f, 8, SEEK_SET);
	uint32_t* a = static_cast<uint32_t*>(std::malloc((size_t)total * 32u));
	if (!a) { std::fclose(f); return dagFail(CBQ_ERROR_OUT_OF_MEMORY, "out of memory for " + std::to_string(total) + " nodes"); }
	for (uint32_t m = 0; m < 256u; m++) for (int c = 0; c < 8; c++) a[(size_t)m * 8 + c] = m;      // the material nodes (storage.cpp:110-122)
	const size_t got = stored ? std::fread(a + 256u * 8u, 32u, (size_t)stored, f) : 0;
	std::fclose(f);
	if (got != stored) { std::free(a); return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": short read"); }
	uint32_t worst = 0;
	for (uint64_t i = 256u * 8u; i < total * 8u; i++) worst = a[i] > worst ? a[i] : worst;
	if (worst >= total) { std::free(a); return dagFail(CBQ_ERROR_CORRUPT_VOLUME, std::string(path) + ": child index " + std::to_string(worst) + " is past the " + std::to_string(total) + " nodes"); }
	*nodes = a; *node_count = total; *root_index = head[0];
	return CBQ_OK;
}

extern "C" void cbq_dag_free(uint32_t* nodes) { std::free(nodes); }

extern "C" int cbq_dag_save(const char* path, const uint32_t* nodes, uint64_t node_count, uint32_t root_index)
{
	if (!path || !nodes || node_count < 256u) return dagFail(CBQ_ERROR_INVALID_ARGUMENT, "the node array must include the 256 material nodes");
	if (node_count - 256u > 0xffffffffull || root_index >= node_count) return dagFail(CBQ_ERROR_INVALID_ARGUMENT, "node count or root index out of range");
	const std::string tmp = std::string(path) + ".partial";
	FILE* f = std::fopen(tmp.c_str(), "wb");
	if (!f) return dagFail(CBQ_ERROR_INVALID_ARGUMENT, "cannot create " + tmp);
	const uint32_t head[2] = { root_index, (uint32_t)(node_count - 256u) };
	bool ok = std::fwrite(head, sizeof(uint32_t), 2, f) == 2;
	const uint64_t stored = node_count - 256u;
	if (ok && stored) ok = std::fwrite(nodes + 256u * 8u, 32u, (size_t)stored, f) == stored;
	ok = (std::fclose(f) == 0) && ok;
	if (!ok || std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return dagFail(CBQ_ERROR_INVALID_ARGUMENT, std::string("cannot write ") + path); }
	return CBQ_OK;
}

// Host-only half of the compact result path: cbq_expand_hits widens 8-byte cbq_hit_compact records into the
// 40-byte cbq_hit records cbq_trace writes -- every field of RayVolumeIntersection (reference
// src/library/raytracing.h:48-55). `position` is re-formed as the reference forms it (raytracing.cpp:463-466):
// origin + dir * (float)distance, one rounded multiply and one rounded add per component. This file is compiled
// with -ffp-contract=off so the compiler cannot fuse them, whatever the host ISA; the GPU kernels do the same
// arithmetic under -fmad=false, hence byte-identical records (tests/test_gpu_trace.py::test_compact_results_*).
#include "../../include/cubiquity_b200.h"

#include <cstring>
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

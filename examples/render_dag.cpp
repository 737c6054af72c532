// Example host program: path-trace a Cubiquity .dag volume on a B200 through the C ABI.
//
//   g++ -std=c++17 -O2 examples/render_dag.cpp -Iinclude -Lcubiquity_b200/lib -lcubiquity_b200
//       -Wl,-rpath,$PWD/cubiquity_b200/lib -o render_dag            (one command line)
//   ./render_dag volume.dag out.ppm [width height spp bounces]
//
// It does what `cubiquity view volume.dag --mode=cpu-pathtracing` does per frame (reference
// src/application/commands/view/pathtracing_demo.cpp:214-260) minus the window: load the volume, place the
// camera like Viewer::onInitialise (viewer.cpp:59-90), accumulate `spp` samples, divide, write 8-bit RGB.
#include "../cubiquity_b200/host/cubiquity_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

// The reference's .dag reader (storage.cpp:192-199,505-528): u32 root, u32 count, count x 8 u32.
static bool loadDag(const char* path, std::vector<uint32_t>& nodes, uint32_t& root)
{
	std::ifstream f(path, std::ios::binary);
	uint32_t count = 0;
	if (!f.read(reinterpret_cast<char*>(&root), 4) || !f.read(reinterpret_cast<char*>(&count), 4)) return false;
	nodes.resize((256 + (size_t)count) * 8);
	for (uint32_t i = 0; i < 256; i++) for (int c = 0; c < 8; c++) nodes[i * 8 + c] = i;   // material nodes (storage.cpp:110-122)
	return (bool)f.read(reinterpret_cast<char*>(nodes.data() + 256 * 8), (std::streamsize)count * 32);
}

// Occupied bounds: walk the tree once, pruning empty space (what computeBounds does, utility.cpp).
static void bounds(const std::vector<uint32_t>& n, uint32_t node, int h, int64_t x, int64_t y, int64_t z, int64_t lo[3], int64_t hi[3])
{
	if (node == 0) return;
	const int64_t s = INT64_C(1) << h;
	if (node < 256 || h == 0) {
		lo[0] = std::min(lo[0], x); lo[1] = std::min(lo[1], y); lo[2] = std::min(lo[2], z);
		hi[0] = std::max(hi[0], x + s - 1); hi[1] = std::max(hi[1], y + s - 1); hi[2] = std::max(hi[2], z + s - 1);
		return;
	}
	if (x >= lo[0] && y >= lo[1] && z >= lo[2] && x + s - 1 <= hi[0] && y + s - 1 <= hi[1] && z + s - 1 <= hi[2]) return;
	const int64_t half = s / 2;
	for (int c = 0; c < 8; c++)
		bounds(n, n[(size_t)node * 8 + c], h - 1, x + ((c & 1) ? half : 0), y + ((c & 2) ? half : 0), z + ((c & 4) ? half : 0), lo, hi);
}

int main(int argc, char** argv)
{
	if (argc < 3) { std::fprintf(stderr, "usage: %s volume.dag out.ppm [width height spp bounces]\n", argv[0]); return 2; }
	const uint32_t width = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 640, height = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 360;
	const uint32_t spp = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 16, bounces = argc > 6 ? (uint32_t)std::atoi(argv[6]) : 1;

	std::vector<uint32_t> nodes;
	uint32_t root = 0;
	if (!loadDag(argv[1], nodes, root)) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }

	CubiquityGPU::GpuVolume gpu(0);
	if (!gpu.ok() || !gpu.upload(nodes.data(), nodes.size() / 8, root)) {
		std::fprintf(stderr, "GPU error: %s\n", CubiquityGPU::GpuVolume::lastError().c_str());
		return 1;
	}

	int64_t lo[3] = { INT64_MAX, INT64_MAX, INT64_MAX }, hi[3] = { INT64_MIN, INT64_MIN, INT64_MIN };
	bounds(nodes, root, 32, -(INT64_C(1) << 31), -(INT64_C(1) << 31), -(INT64_C(1) << 31), lo, hi);
	const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
	const double hd = 0.5 * std::sqrt(double(hi[0] - lo[0]) * (hi[0] - lo[0]) + double(hi[1] - lo[1]) * (hi[1] - lo[1]) + double(hi[2] - lo[2]) * (hi[2] - lo[2]));
	const double position[3] = { cx, cy - hd, cz + hd };            // viewer.cpp:71-79
	cbq_camera cam;
	cbq_camera_from_pose(position, -(3.14159265358979f / 4.0f), 0.0, 60.0, &cam);

	cbq_pt_params p = {};
	p.width = width; p.height = height; p.spp = spp; p.bounces = bounces;
	p.variant = bounces == 1 ? CBQ_VARIANT_ONE_BOUNCE : CBQ_VARIANT_RECURSIVE;
	p.include_sun = p.include_sky = p.add_noise = 1; p.max_footprint = 0.0035f;      // pathtracing_demo.h:80-84
	p.x1 = width; p.y1 = height;
	std::vector<float> image((size_t)width * height * 3, 0.0f);
	if (cbq_render(gpu.context(), &cam, &p, image.data()) != CBQ_OK) { std::fprintf(stderr, "render: %s\n", cbq_last_error()); return 1; }

	// pathtracing_demo.cpp:240-256: scale by 255 / frames, clamp, truncate.
	std::ofstream out(argv[2], std::ios::binary);
	out << "P6\n" << width << " " << height << "\n255\n";
	const float scale = 255.0f / (float)spp;
	for (float v : image) { float f = std::min(std::max(v * scale, 0.0f), 255.0f); out.put((char)(unsigned char)f); }
	std::printf("%u x %u, %u spp, %u bounce(s): mean %.4f\n", width, height, spp, bounces,
		(double)([&] { double s = 0; for (float v : image) s += v; return s; })() / (double)image.size() / spp);
	return 0;
}

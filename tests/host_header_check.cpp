// TEST: the C++ host mirror (cubiquity_b200/host/cubiquity_gpu.h) compiles without CUDA headers and
// behaves like the reference interface. argv[1] = a .dag file; prints "key value" lines for pytest.
#include "../cubiquity_b200/host/cubiquity_gpu.h"

#include <cstdio>
#include <fstream>
#include <vector>

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	std::ifstream f(argv[1], std::ios::binary);
	uint32_t root = 0, count = 0;
	f.read(reinterpret_cast<char*>(&root), 4);
	f.read(reinterpret_cast<char*>(&count), 4);
	std::vector<uint32_t> nodes((256 + (size_t)count) * 8);
	for (uint32_t i = 0; i < 256; i++) for (int c = 0; c < 8; c++) nodes[i * 8 + c] = i;   // storage.cpp:110-122
	f.read(reinterpret_cast<char*>(nodes.data() + 256 * 8), (std::streamsize)count * 32);

	CubiquityGPU::SubDAGArray sd;
	if (!CubiquityGPU::findSubDAGs(nodes.data(), 256 + count, root, sd)) { std::printf("subdags failed\n"); return 1; }
	for (int i = 0; i < 8; i++) std::printf("subdag %d %d %d %d %d %u\n", i, sd[i].lower[0], sd[i].lower[1], sd[i].lower[2], sd[i].height, sd[i].node);

	CubiquityGPU::GpuVolume vol(0);
	if (!vol.ok()) { std::printf("nogpu %s\n", CubiquityGPU::GpuVolume::lastError().c_str()); return 0; }
	if (!vol.upload(nodes.data(), 256 + count, root)) { std::printf("upload failed %s\n", CubiquityGPU::GpuVolume::lastError().c_str()); return 1; }
	// A ray straight down the z axis from above the volume, off the voxel boundaries.
	CubiquityGPU::RayVolumeIntersection r = CubiquityGPU::intersectVolume(vol, 0.25f, 0.125f, 100.0f, 0.001f, 0.002f, -1.0f, true);
	std::printf("hit %d %.9g %u %.9g %.9g %.9g %g %g %g\n", (int)r.hit, r.distance, r.material, r.position[0], r.position[1], r.position[2],
		r.normal[0], r.normal[1], r.normal[2]);
	// The volume's life cycle on the device through the same mirror: carve where the ray hit, trace again (the hit
	// moves away), undo by switching back to the old root (the hit is back), bake and read the merged array back.
	uint32_t carved = 0;
	if (!vol.fillSphere(r.position[0], r.position[1], r.position[2], 3.0f, 0, carved)) { std::printf("fill failed %s\n", CubiquityGPU::GpuVolume::lastError().c_str()); return 1; }
	CubiquityGPU::RayVolumeIntersection after = CubiquityGPU::intersectVolume(vol, 0.25f, 0.125f, 100.0f, 0.001f, 0.002f, -1.0f, true);
	if (!vol.setRoot(root)) return 1;
	CubiquityGPU::RayVolumeIntersection undone = CubiquityGPU::intersectVolume(vol, 0.25f, 0.125f, 100.0f, 0.001f, 0.002f, -1.0f, true);
	uint64_t bakedCount = 0; uint32_t bakedRoot = 0;
	if (!vol.bake(bakedCount, bakedRoot)) { std::printf("bake failed %s\n", CubiquityGPU::GpuVolume::lastError().c_str()); return 1; }
	std::vector<uint32_t> baked((size_t)bakedCount * 8);
	if (!vol.download(0, bakedCount, baked.data())) return 1;
	CubiquityGPU::RayVolumeIntersection rebaked = CubiquityGPU::intersectVolume(vol, 0.25f, 0.125f, 100.0f, 0.001f, 0.002f, -1.0f, true);
	std::printf("lifecycle %d %.9g %d %.9g %llu %u %d %.9g\n", (int)after.hit, after.distance, (int)undone.hit, undone.distance,
		(unsigned long long)bakedCount, baked[(size_t)bakedRoot * 8] | baked[(size_t)bakedRoot * 8 + 7], (int)rebaked.hit, rebaked.distance);
	return 0;
}

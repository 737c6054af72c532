// Device helpers shared by the ray-cast and path-tracing kernels: the per-pixel part of the camera,
// the reference's hashes / RNG and its surface shading. COMPILE WITH -fmad=false.
#pragma once

#include "cbq_internal.h"

namespace cbq {

// Camera::rayFromViewportPos (reference src/application/commands/view/camera.cpp:19-35) followed by
// static_cast<Ray3f> (pathtracing_demo.cpp:220). Mixed float/double exactly as written there; the
// trigonometry of camera.cpp:24,40-66 arrives precomputed in cbq_camera.
__device__ __forceinline__ void cameraRay(const cbq_camera& c, int x, int y, int width, int height, Ray& out)
{
	const double invWidth = (double)(1.0f / (float)width);
	const double invHeight = (double)(1.0f / (float)height);
	const float aspect = (float)width / (float)height;
	const float xOff = ((float)x - ((float)width / 2.0f)) + 0.5f;
	const float yOff = ((float)y - ((float)height / 2.0f)) + 0.5f;
	const double kx = ((invWidth * (double)xOff) * (double)aspect) * (double)c.scale;
	const double ky = (invHeight * (double)yOff) * (double)c.scale;
	double dir[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		double t = c.position[a] + c.forward[a];
		t += c.right[a] * kx;
		t -= c.up[a] * ky;
		dir[a] = t - c.position[a];
	}
	const double len = sqrt(((0.0 + dir[0] * dir[0]) + dir[1] * dir[1]) + dir[2] * dir[2]);
#pragma unroll
	for (int a = 0; a < 3; a++) {
		out.o[a] = (float)c.position[a];
		out.d[a] = (float)(dir[a] / len);
	}
}

__device__ __forceinline__ uint64_t bitMix64(uint64_t b)   // base.cpp:72-77
{
	b = ((b >> 32) ^ b) * 0x0e9846af9b1a615dull;
	b = ((b >> 32) ^ b) * 0x0e9846af9b1a615dull;
	return (b >> 28) ^ b;
}

__device__ __forceinline__ uint32_t fmix32(uint32_t h)   // glsl/pathtracing.frag:287-296
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}

// Per-sample RNG seed: hashRay(primary ray) ^ bitMix(frameId) (glsl/pathtracing.frag:770-780,786,816).
__device__ __forceinline__ uint32_t rayHash(const Ray& r)
{
	uint32_t h = 0;
	h ^= fmix32(__float_as_uint(r.o[0])); h ^= fmix32(__float_as_uint(r.o[1])); h ^= fmix32(__float_as_uint(r.o[2]));
	h ^= fmix32(__float_as_uint(r.d[0])); h ^= fmix32(__float_as_uint(r.d[1])); h ^= fmix32(__float_as_uint(r.d[2]));
	return h;
}
__device__ __forceinline__ uint32_t pixelSeed(const Ray& r, uint32_t sampleIndex) { return rayHash(r) ^ fmix32(sampleIndex); }

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
	return ((0.0f + ax * bx) + ay * by) + az * bz;   // linalg sum(a*b): fold from 0, left to right
}

// randomPointInUnitSphere, pathtracing_demo.cpp:62-79. The try cap is quirk Q8: the 32-bit
// truncation of the 64-bit mixer is not a bijection, a stream can fall into a cycle of rejected points
// (about one seed in 4e8) and the reference's loop would then spin for ever; after kMaxBallTries
// rejected candidates the last one is returned as it is. Same rule as the CPU checker.
constexpr int kMaxBallTries = 64;
// A coordinate takes one of 1024 values, (k - 511.5f) / 511.5f: the kernels keep them in a shared-memory table, filled
// with the reference's own expression, instead of three IEEE divisions per candidate (half of the shade kernel's
// instructions went there, profiles/r02_analysis.md).
constexpr int kBallLutSize = 1024;
__device__ __forceinline__ void fillBallLut(float* lut)   // every thread of the CTA; ends with a barrier
{
	for (int k = threadIdx.x; k < kBallLutSize; k += blockDim.x) lut[k] = ((float)k - 511.5f) / 511.5f;
	__syncthreads();
}
__device__ __forceinline__ void unitBallPoint(const float* __restrict__ lut, uint32_t& rng, float& x, float& y, float& z)
{
	int tries = 0;
	do {
		rng = (uint32_t)bitMix64((uint64_t)rng);
		x = lut[rng & 0x3FFu]; y = lut[(rng >> 10) & 0x3FFu]; z = lut[(rng >> 20) & 0x3FFu];
	} while (dot3(x, y, z, x, y, z) >= 1.0f && ++tries < kMaxBallTries);
}

// surfaceColour + positionBasedNoise, pathtracing_demo.cpp:36-60
__device__ __forceinline__ void surfaceColour(const float4* __restrict__ colours, uint32_t material, const float position[3],
	bool addNoise, float& r, float& g, float& b)
{
	const float4 base = __ldg(colours + material);
	r = base.x; g = base.y; b = base.z;
	if (addNoise) {
		// fnv1a (base.cpp:62-69) over the 12 bytes of ivec3(position + 0.499)
		const int cell[3] = { (int)(position[0] + 0.499f), (int)(position[1] + 0.499f), (int)(position[2] + 0.499f) };
		uint64_t hsh = 0xcbf29ce484222325ull;
#pragma unroll
		for (int a = 0; a < 3; a++) {
#pragma unroll
			for (int k = 0; k < 4; k++) { hsh ^= (uint64_t)(((uint32_t)cell[a] >> (8 * k)) & 0xffu); hsh *= 0x00000100000001B3ull; }
		}
		float noise = (float)((uint32_t)hsh & 0xffu) / 255.0f;
		noise = (float)(((double)noise * 0.1) + 0.9);
		r *= noise; g *= noise; b *= noise;
	}
}

// normalize(vec3(1, -2, 10)), pathtracing_demo.cpp:89: IEEE sqrt and divides, same bits as the host.
__device__ __forceinline__ void sunDirection(float& x, float& y, float& z)
{
	const float len = sqrtf(dot3(1.0f, -2.0f, 10.0f, 1.0f, -2.0f, 10.0f));
	x = 1.0f / len; y = -2.0f / len; z = 10.0f / len;
}

} // namespace cbq

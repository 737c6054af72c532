// The screen-space passes of the reference's GPU path-tracing viewer as CUDA kernels:
//   progressive accumulation  gpu_pathtracing_viewer.cpp:152-161 (additive blending GL_ONE, GL_ONE into an RGBA32F
//                             target) + glsl/pathtracing.frag:789-803 (frame f draws the 64x64-pixel tiles whose
//                             bitMix(tile id) % 16 == f % 16); alpha counts the samples a pixel has received
//   normalise                 glsl/normalise.frag: rgba / a
//   blur                      glsl/horz_blur.frag, glsl/vert_blur.frag: 9 taps, a tap counts only if its alpha is within
//                             0.5 of the centre's; texture coordinates wrap (the viewer leaves GL_TEXTURE_WRAP at its
//                             default, GL_REPEAT: gpu_pathtracing_viewer.cpp:100-104 sets the filters only)
// All three are bandwidth-trivial element-wise / short-stencil kernels: one thread per pixel, float4 accesses.
#include "cbq_internal.h"
#include "shading.cuh"

namespace cbq {

namespace {

// glsl/pathtracing.frag:792-799 with windowPos counted from the TOP-left pixel (gl_FragCoord counts from the bottom;
// the schedule is a pseudo-random partition of the tiles either way).
__device__ __forceinline__ uint32_t tileGroupOf(uint32_t x, uint32_t y, uint32_t groupCount)
{
	const uint32_t tx = x >> 6, ty = y >> 6;
	return fmix32((tx << 16) | (ty & 0xffffu)) % groupCount;
}

__global__ void __launch_bounds__(256)
progressiveAdd(float* __restrict__ scratch, float4* __restrict__ rgba, uint32_t width, uint32_t height, uint32_t groupCount, uint32_t groupIndex, float samples)
{
	const uint64_t n = (uint64_t)width * height;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
		if (tileGroupOf(x, y, groupCount) != groupIndex) continue;
		float* s = scratch + 3 * i;
		float4 v = rgba[i];
		v.x += s[0]; v.y += s[1]; v.z += s[2]; v.w += samples;
		rgba[i] = v;
		s[0] = s[1] = s[2] = 0.0f;      // the scratch image is all zeros again for the next pass
	}
}

__global__ void __launch_bounds__(256)
normalise(const float4* __restrict__ rgba, uint64_t n, float* __restrict__ rgb)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const float4 v = rgba[i];
		rgb[3 * i + 0] = v.x / v.w; rgb[3 * i + 1] = v.y / v.w; rgb[3 * i + 2] = v.z / v.w;     // col.rgba / col.a
	}
}

template <bool kVertical>
__global__ void __launch_bounds__(256)
blur(const float4* __restrict__ in, float4* __restrict__ out, uint32_t width, uint32_t height)
{
	const uint64_t n = (uint64_t)width * height;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const int x = (int)(i % width), y = (int)(i / width);
		const float4 c = in[i];
		float r = c.x, g = c.y, b = c.z, count = 1.0f;
		for (int tap = 1; tap <= 4; tap++) {          // taps = 9, halfTaps = 4; the minus tap first, then the plus tap (horz_blur.frag:19-26)
			for (int sgn = -1; sgn <= 1; sgn += 2) {
				int sx = x, sy = y;
				if (kVertical) sy = (y + sgn * tap + (int)height) % (int)height; else sx = (x + sgn * tap + (int)width) % (int)width;
				const float4 t = in[(uint64_t)sy * width + sx];
				if (fabsf(c.w - t.w) < 0.5f) { r += t.x; g += t.y; b += t.z; count += 1.0f; }
			}
		}
		out[i] = make_float4(r / count, g / count, b / count, c.w);
	}
}

int gridFor(uint64_t n) { uint64_t b = (n + 255) / 256; if (b > 148u * 8u) b = 148u * 8u; return (int)(b ? b : 1); }

} // namespace

cudaError_t launchProgressiveAdd(float* scratchRgb, float* rgba, uint32_t width, uint32_t height, uint32_t groupCount, uint32_t groupIndex, float samples, cudaStream_t stream)
{
	progressiveAdd<<<gridFor((uint64_t)width * height), 256, 0, stream>>>(scratchRgb, reinterpret_cast<float4*>(rgba), width, height, groupCount, groupIndex, samples);
	return cudaGetLastError();
}

cudaError_t launchNormalise(const float* rgba, uint32_t width, uint32_t height, float* rgb, cudaStream_t stream)
{
	const uint64_t n = (uint64_t)width * height;
	normalise<<<gridFor(n), 256, 0, stream>>>(reinterpret_cast<const float4*>(rgba), n, rgb);
	return cudaGetLastError();
}

cudaError_t launchBlur(const float* in, float* out, uint32_t width, uint32_t height, int vertical, cudaStream_t stream)
{
	const int grid = gridFor((uint64_t)width * height);
	if (vertical) blur<true><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), width, height);
	else blur<false><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), width, height);
	return cudaGetLastError();
}

} // namespace cbq

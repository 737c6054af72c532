// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Headless driver for the reference's CPU path tracer. The UNMODIFIED translation units
//   src/application/commands/view/pathtracing_demo.cpp   (the bounce loop: traceSingleRay,
//        traceSingleRayRecurse, gatherLighting, surfaceColour, randomPointInUnitSphere)
//   src/application/commands/view/camera.cpp             (Camera::rayFromViewportPos)
//   src/application/base/logging.cpp
// are compiled against oracle/stubs/SDL.h and linked with this file, which supplies
//   * the Window / Viewer member functions those units reference (the reference's own viewer.cpp and
//     window.cpp need a real SDL + OpenGL and are not built), and
//   * an extern "C" entry point that renders a pixel rectangle by calling the reference's private
//     traceSingleRay / traceSingleRayRecurse directly.
// The reference's RNG is ONE process-global u32 (pathtracing_demo.cpp:33, `nextPointInUnitSphere`,
// external linkage). To obtain an order-independent image we re-seed that global before every
// sample with the per-pixel rule documented in include/cubiquity_b200.h; everything downstream of
// the seed is the reference's code.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

// Reach the private bounce-loop members of PathtracingDemo. Standard headers are all included above,
// so only the reference's class definitions see these.
#define private public
#define protected public
#include "commands/view/pathtracing_demo.h"
#undef private
#undef protected

extern u32 nextPointInUnitSphere; // pathtracing_demo.cpp:33

namespace {
const float* g_colours = nullptr; // 256 x 3, consumed by the Viewer constructor below

uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }

uint32_t pixelSeed(const Ray3f& r, uint32_t sampleIndex)
{
	// glsl/pathtracing.frag:770-780 (hashRay) ^ bitMix(frameId) (:786,816)
	const float f[6] = { r.mOrigin.x, r.mOrigin.y, r.mOrigin.z, r.mDir.x, r.mDir.y, r.mDir.z };
	uint32_t bits[6], h = 0;
	std::memcpy(bits, f, sizeof(bits));
	for (int i = 0; i < 6; i++) h ^= fmix32(bits[i]);
	return h ^ fmix32(sampleIndex);
}
}

// ---- the pieces of window.cpp / viewer.cpp that pathtracing_demo.cpp links against ---------------
int Window::width() { return 0; }
int Window::height() { return 0; }
SDL_Surface* Window::surface() { return nullptr; }
MouseButtonState Window::mouseButtonState(Uint32) { return MouseButtonState::Up; }

Viewer::Viewer(const std::string& filename, WindowType windowType)
	: Window(windowType), mFrameNumber(0)
{
	mVolume.load(filename);
	// viewer.cpp:50-55: purple unless the metadata names a colour; here the table is handed in.
	for (int i = 0; i < 256; i++) {
		mColours[i] = g_colours ? vec3{ g_colours[3 * i], g_colours[3 * i + 1], g_colours[3 * i + 2] } : vec3{ 1.0f, 0.0f, 1.0f };
	}
}
void Viewer::onInitialise() {}
void Viewer::onUpdate(float) {}
void Viewer::onKeyUp(const SDL_KeyboardEvent&) {}
void Viewer::onMouseButtonDown(const SDL_MouseButtonEvent&) {}
void Viewer::onMouseButtonUp(const SDL_MouseButtonEvent&) {}
void Viewer::onMouseMotion(const SDL_MouseMotionEvent&) {}

extern "C" {
SDL_Surface* SDL_CreateRGBSurface(Uint32, int, int, int, Uint32, Uint32, Uint32, Uint32) { return nullptr; }
void SDL_FreeSurface(SDL_Surface*) {}
int SDL_BlitScaled(SDL_Surface*, const SDL_Rect*, SDL_Surface*, SDL_Rect*) { return 0; }

// params: width, height, spp, bounces, variant, sun, sky, noise, frame_id, x0, y0, x1, y1
// accum: width * height * 3 floats, ADDED to. Returns seconds. reseed = 0 leaves the reference's
// global stream alone (its as-is, scan-order-dependent behaviour).
double ref_pt_render(const char* dagPath, const float* colours768, const double camPos[3], double pitch, double yaw,
	const uint32_t* params, float maxFootprint, int reseed, float* accum)
{
	g_colours = colours768;
	PathtracingDemo demo{ std::string(dagPath) };
	demo.onVolumeModified();             // computes subDAGs exactly as the viewer does (pathtracing_demo.cpp:335-341)
	const uint32_t width = params[0], height = params[1], spp = params[2];
	demo.bounces = params[3];
	const uint32_t variant = params[4];
	demo.includeSun = params[5] != 0;
	demo.includeSky = params[6] != 0;
	demo.addNoise = params[7] != 0;
	demo.maxFootprint = maxFootprint;
	const uint32_t frameId = params[8], x0 = params[9], y0 = params[10], x1 = params[11], y1 = params[12];

	Camera camera;
	camera.position = dvec3{ camPos[0], camPos[1], camPos[2] };
	camera.pitch = pitch;
	camera.yaw = yaw;

	nextPointInUnitSphere = 17;
	auto t0 = std::chrono::steady_clock::now();
	for (uint32_t y = y0; y < y1; y++) {
		for (uint32_t x = x0; x < x1; x++) {
			Ray3f ray = static_cast<Ray3f>(camera.rayFromViewportPos((int)x, (int)y, (int)width, (int)height)); // pathtracing_demo.cpp:220
			for (uint32_t s = 0; s < spp; s++) {
				if (reseed) nextPointInUnitSphere = pixelSeed(ray, frameId + s);
				vec3 pixel = (variant == 0) ? demo.traceSingleRay(ray, 0) : demo.traceSingleRayRecurse(ray, 0);
				float* px = accum + 3 * ((size_t)y * width + x);
				px[0] += pixel.x; px[1] += pixel.y; px[2] += pixel.z;
			}
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// Camera::rayFromViewportPos for a whole frame (camera.cpp:12-38), cast to float.
void ref_camera_rays(const double camPos[3], double pitch, double yaw, int width, int height, float* rays6)
{
	Camera camera;
	camera.position = dvec3{ camPos[0], camPos[1], camPos[2] };
	camera.pitch = pitch;
	camera.yaw = yaw;
	for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {
		Ray3f r = static_cast<Ray3f>(camera.rayFromViewportPos(x, y, width, height));
		float* o = rays6 + 6 * ((size_t)y * width + x);
		o[0] = r.mOrigin.x; o[1] = r.mOrigin.y; o[2] = r.mOrigin.z; o[3] = r.mDir.x; o[4] = r.mDir.y; o[5] = r.mDir.z;
	}
}
}

// The geometry the reference's voxeliser evaluates per voxel and per triangle, restated for host and device with the
// reference's own operation order (compile with -fmad=false / -ffp-contract=off):
//   windingNumber          computeWindingNumber(queryPoint, triangles), src/library/voxelization.cpp:199-248 (the flat
//                          list version: Jacobson et al. 2013, generalized winding numbers)
//   pointTriangleDistance  distance(point, triangle), src/library/geometry.cpp:25-67
//   rayHitsTriangle        intersect(ray, triangle, t), src/library/geometry.cpp:71-108
// Vector arithmetic as in the reference's Vec3 (src/library/geometry.h:137-148): dot = x*x' + y*y' + z*z' left to right,
// cross component-wise, length = sqrt(dot).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CBQ_MM __host__ __device__ __forceinline__
#else
#define CBQ_MM inline
#endif

namespace cbq {

struct V3 { float x, y, z; };
CBQ_MM V3 operator-(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
CBQ_MM V3 operator+(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
CBQ_MM V3 operator*(V3 a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
CBQ_MM V3 operator/(V3 a, float s) { return V3{ a.x / s, a.y / s, a.z / s }; }
CBQ_MM float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CBQ_MM V3 cross(V3 a, V3 b) { return V3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
CBQ_MM float length(V3 v) { return sqrtf(dot(v, v)); }

struct Tri { V3 v[3]; };   // 9 floats, the layout of cbq_voxelize's triangle array

// One triangle's term of the winding number sum: 2 * atan2(numerator, denominator) or 0 (voxelization.cpp:203-239).
CBQ_MM float windingTerm(V3 q, const Tri& t)
{
	V3 qa = t.v[0] - q, qb = t.v[1] - q, qc = t.v[2] - q;
	const float al = length(qa), bl = length(qb), cl = length(qc);
	if (al != 0 && bl != 0 && cl != 0) {
		qa = qa / al; qb = qb / bl; qc = qc / cl;
		const float numerator = dot(qa, cross(qb - qa, qc - qa));
		const float denominator = 1.0f + dot(qa, qb) + dot(qa, qc) + dot(qb, qc);
		if (numerator != 0) return 2.0f * atan2f(numerator, denominator);
	}
	return 0.0f;
}

// Normalisation to [-1, 1]: `windingNumber /= 4.0 * pi` is a float divided by a double (voxelization.cpp:242).
CBQ_MM float windingNormalise(float sum) { return (float)((double)sum / (4.0 * 3.14159265358979323846)); }

// isInside's threshold (voxelization.cpp:356-367).
CBQ_MM bool windingInside(float w) { return fabsf(w) > 0.5f + 0.001f; }

CBQ_MM float clampStd(float n, float lower, float upper)   // geometry.cpp:20-22: std::max(lower, std::min(n, upper))
{
	const float m = (upper < n) ? upper : n;
	return (lower < m) ? m : lower;
}

CBQ_MM float pointTriangleDistance(V3 point, const Tri& t)
{
	V3 edges[3], toPoint[3];
	for (int i = 0; i < 3; i++) { edges[i] = t.v[(i + 1) % 3] - t.v[i]; toPoint[i] = point - t.v[i]; }
	const V3 normal = cross(edges[0], edges[2]);
	const bool inside = dot(cross(edges[0], normal), toPoint[0]) > 0.0f &&
	                    dot(cross(edges[1], normal), toPoint[1]) > 0.0f &&
	                    dot(cross(edges[2], normal), toPoint[2]) > 0.0f;
	float d2 = 3.402823466e+38f;
	if (inside) {
		d2 = dot(normal, toPoint[0]) * dot(normal, toPoint[0]) / dot(normal, normal);
	} else {
		for (int i = 0; i < 3; i++) {
			const V3 e = edges[i] * clampStd(dot(edges[i], toPoint[i]) / dot(edges[i], edges[i]), 0.0f, 1.0f) - toPoint[i];
			const float c = dot(e, e);
			d2 = (d2 < c) ? d2 : c;       // std::min(candidate, distanceSquared)
		}
	}
	return sqrtf(d2);
}

CBQ_MM bool rayHitsTriangle(V3 origin, V3 dir, const Tri& tri, float& t)
{
	const V3 normal = cross(tri.v[1] - tri.v[0], tri.v[2] - tri.v[0]);
	const float nd = dot(normal, dir);
	if (fabsf(nd) < 1e-6f) return false;
	const float d = -dot(normal, tri.v[0]);
	t = -(dot(normal, origin) + d) / nd;
	if (t < 0) return false;
	const V3 p = origin + (dir * t);
	for (int i = 0; i < 3; i++) {
		const V3 edge = tri.v[(i + 1) % 3] - tri.v[i];
		const V3 vp = p - tri.v[i];
		if (dot(normal, cross(edge, vp)) < 0) return false;
	}
	return true;
}

// "A Topological Approach to Voxelization": the triangle meets one of the voxel's three axis-aligned intersection
// targets (voxelization.cpp:470-483).
CBQ_MM bool touchesIntersectionTarget(int x, int y, int z, const Tri& tri)
{
	for (int axis = 0; axis < 3; axis++) {
		V3 o{ (float)x, (float)y, (float)z }, d{ 0.0f, 0.0f, 0.0f };
		if (axis == 0) { o.x -= 0.5f; d.x = 1.0f; } else if (axis == 1) { o.y -= 0.5f; d.y = 1.0f; } else { o.z -= 0.5f; d.z = 1.0f; }
		float t = 0.0f;
		if (rayHitsTriangle(o, d, tri, t) && t <= 1.0f) return true;
	}
	return false;
}

} // namespace cbq

/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for cubiquity_b200.
 *
 * A plain-C restatement of the reference's SVDAG ray cast and path-tracer bounce loop, used
 * as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg. It is
 * never linked into, loaded by or called from the product library (cubiquity_b200/).
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py compares every function here,
 * bit for bit, with the unmodified reference compiled from /root/reference into
 * oracle/_ref/libcbq_ref.so (oracle/Makefile), and tests/golden/ holds vectors generated
 * by that reference build (tests/golden/make_golden.py) so the pin also holds on machines
 * where /root/reference is absent. The reference's own golden fixture for this path
 * (data/tests/axis.dag, 124096 hits; test_rendering.cpp:100-117) is a network download
 * and is not available offline; its hash KATs (test_base.cpp:14-35) are checked.
 */
#ifndef CBQ_ORACLE_H
#define CBQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same record layouts as include/cubiquity_b200.h so buffers compare byte for byte. */
typedef struct { float o[3]; float d[3]; } cbqo_ray;           /* 24 B */
typedef struct {
	uint32_t hit;        /* RayVolumeIntersection::hit      (raytracing.h:50) */
	float    distance;   /* ::distance -- a float value the reference stores in a double (:51) */
	uint32_t material;   /* ::material (:52) */
	float    position[3];/* ::position (:53) */
	float    normal[3];  /* ::normal   (:54) */
	uint32_t pad;        /* 0; 1 when the traversal was abandoned at the iteration cap (quirk Q5) */
} cbqo_hit;                                                     /* 40 B */

/* Mirrors struct SubDAG (raytracing.h:57-65): 32 bytes, std430 compatible. */
typedef struct {
	int32_t  lower[3];
	int32_t  height;
	uint32_t pad0;
	uint32_t node;
	uint32_t pad1, pad2;
} cbqo_subdag;

/* Per-batch traversal statistics (the "instrumented oracle" of SURVEY 8d). */
typedef struct {
	uint64_t rays;
	uint64_t hits;
	uint64_t subdag_entries;   /* calls that passed the slab test of raytracing.cpp:232 */
	uint64_t iterations;       /* trips round the do/while of raytracing.cpp:253-367 */
	uint64_t descents;         /* PUSH  (raytracing.cpp:279-301) */
	uint64_t pops;             /* POP   (raytracing.cpp:339-364) */
	uint64_t material_steps;   /* node reads inside findNearestMaterial (raytracing.cpp:148-159) */
} cbqo_stats;

/* Camera with the trigonometry already applied (camera.cpp:40-66 evaluated once per frame on
 * the host); the per-pixel part of Camera::rayFromViewportPos (camera.cpp:19-35) is restated
 * in cbqo_camera_ray. */
typedef struct {
	double position[3];
	double forward[3];
	double up[3];
	double right[3];
	float  scale;       /* tan(fovInDegrees * 0.0174533f * 0.5f) * 2.0f  (camera.cpp:24) */
	float  pad;
} cbqo_camera;

typedef struct {
	uint32_t width, height;
	uint32_t spp;            /* samples per pixel accumulated by one call */
	uint32_t bounces;        /* PathtracingDemo::bounces (pathtracing_demo.h:80) */
	uint32_t variant;        /* 0 = traceSingleRay (pathtracing_demo.cpp:150-190), 1 = traceSingleRayRecurse (:120-148) */
	uint32_t include_sun, include_sky, add_noise; /* pathtracing_demo.h:81-83 */
	float    max_footprint;  /* pathtracing_demo.h:84 */
	uint32_t frame_id;       /* first sample index; sample s of the call uses frame_id + s */
	uint32_t x0, y0, x1, y1; /* pixel rectangle [x0,x1) x [y0,y1) to render */
	uint32_t pad;
} cbqo_pt_params;

uint64_t cbqo_iteration_cap(int subdag_height);

void cbqo_find_subdags(const uint32_t* nodes, uint32_t root, cbqo_subdag out[8]);

void cbqo_intersect(const uint32_t* nodes, const cbqo_subdag sd[8], const cbqo_ray* ray,
	int surface_props, float max_footprint, cbqo_hit* out, cbqo_stats* stats /* nullable */);

/* Traces n rays with `threads` host threads; returns seconds in the loop. hits/stats nullable. */
double cbqo_trace(const uint32_t* nodes, const cbqo_subdag sd[8], const cbqo_ray* rays, uint64_t n,
	int surface_props, float max_footprint, cbqo_hit* hits, int threads, cbqo_stats* stats);

void cbqo_camera_from_pose(const double position[3], double pitch, double yaw, double fov_degrees, cbqo_camera* out);
void cbqo_camera_ray(const cbqo_camera* cam, int x, int y, int width, int height, cbqo_ray* out);
void cbqo_camera_rays(const cbqo_camera* cam, int width, int height, cbqo_ray* out);

void cbqo_unit_ball_points(uint32_t* state, int draws, float* out /* draws x 3 */);
uint32_t cbqo_pixel_seed(const cbqo_ray* primary, uint32_t sample_index);

/* accum: width*height*3 floats, row-major, ADDED to (like mImage += pixel, pathtracing_demo.cpp:224).
 * colours: 256*3 floats. rays_out (nullable) receives the number of intersectVolume calls made.
 * Returns seconds. */
double cbqo_render(const uint32_t* nodes, const cbqo_subdag sd[8], const float* colours,
	const cbqo_camera* cam, const cbqo_pt_params* p, float* accum, int threads, uint64_t* rays_out);
void cbqo_last_render_stats(cbqo_stats* out);   /* node visits etc. of the last cbqo_render (instrumentation) */

uint64_t cbqo_bit_mix64(uint64_t x);
uint64_t cbqo_fnv1a(const void* data, int64_t len);
uint32_t cbqo_fmix32(uint32_t h);

#ifdef __cplusplus
}
#endif
/* EXPERIMENT (design study): see cbq_oracle.c. 0 = off (the default; every test runs with it off). */
#ifdef CBQO_EXPERIMENTS   /* oracle/experiments/, never part of the parity checker */
void cbqo_experiment_set_brick_height(int h);
void cbqo_experiment_set_grid_height(int h);
void cbqo_experiment_grid_counters(uint64_t out[2]);   /* sub-DAGs crossed by the grid walk, grid cells visited; resets */
#endif

/* Event strings per ray ('O' sub-DAG entered, 'D' descend, 'A' advance, 'P' advance + pop, 'H' hit): events is n * cap bytes. */
void cbqo_trace_events(const uint32_t* nodes, const cbqo_subdag sd[8], const cbqo_ray* rays, uint64_t n,
	int surf, float maxFootprint, uint8_t* events, uint32_t cap, uint32_t* counts);

/* As cbqo_trace_events, but each event byte is kind (0 D, 1 A, 2 P, 3 H, 4 O) + 8 * height of the node it happens in. */
void cbqo_trace_events_with_heights(const uint32_t* nodes, const cbqo_subdag sd[8], const cbqo_ray* rays, uint64_t n,
	int surf, float maxFootprint, uint8_t* events, uint32_t cap, uint32_t* counts);

/* Diagnostics: events (D, A, P, H, O) by the height of the node they happen in; hist has 5 * 34 counters. */
void cbqo_trace_event_heights(const uint32_t* nodes, const cbqo_subdag sd[8], const cbqo_ray* rays, uint64_t n,
	int surf, float maxFootprint, uint64_t* hist);

#endif

/* cubiquity_b200 -- C ABI of the Blackwell-native SVDAG ray cast and path tracer.
 *
 * This is the drop-in boundary for ONE path of Cubiquity (reference tree under
 * /root/reference, citations are relative to it): ray traversal of the Sparse Voxel DAG
 * and the path-tracer bounce loop built on it. The reference has no FFI of its own
 * ("there is no C API at present", README.md:136-137); each entry point below names the
 * C++ interface it stands in for. Plain pointers and sizes only, no exceptions cross the
 * boundary, every function returns a status code (CBQ_OK == 0) and cbq_last_error() gives
 * the message, in the manner of the existing C-style wrappers (src/library/cubiquity.h:22-26).
 *
 * One cbq_context per GPU. Calls on one context are serialised (stream ordered); different
 * contexts may be driven from different host threads.
 *
 * There is NO CPU fallback anywhere behind this header: without a CUDA device every compute
 * entry point fails with CBQ_ERROR_NO_DEVICE.
 */
#ifndef CUBIQUITY_B200_H
#define CUBIQUITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBQ_VERSION 100

enum {
	CBQ_OK = 0,
	CBQ_ERROR_INVALID_ARGUMENT = 1,
	CBQ_ERROR_NO_DEVICE = 2,
	CBQ_ERROR_CUDA = 3,
	CBQ_ERROR_OUT_OF_MEMORY = 4,
	CBQ_ERROR_NO_VOLUME = 5,
	CBQ_ERROR_CORRUPT_VOLUME = 6
};

/* ---- records ------------------------------------------------------------------------- */

/* A ray: the six floats every caller of intersectVolume passes (src/library/raytracing.h:72-75). */
typedef struct cbq_ray { float origin[3]; float dir[3]; } cbq_ray;                  /* 24 bytes */

/* Every field of struct RayVolumeIntersection (src/library/raytracing.h:48-55). `distance` is
 * the float the reference computes and then widens to double (raytracing.cpp:305); the C++
 * shim in cubiquity_b200/host/cubiquity_gpu.h widens it back. */
typedef struct cbq_hit {
	uint32_t hit;
	float    distance;
	uint32_t material;
	float    position[3];
	float    normal[3];
	uint32_t status;      /* 0, or CBQ_HIT_ABANDONED: see "degenerate rays" in DESIGN.md */
} cbq_hit;                                                                          /* 40 bytes */
#define CBQ_HIT_ABANDONED 1u

/* The same result in 8 bytes, for callers on the far side of PCIe (cbq_trace_compact): everything cbq_hit carries
 * except `position`, which is a function of the ray and `distance` alone -- origin + dir * (float)distance with an
 * un-fused multiply and add (raytracing.cpp:463-466) -- and is re-formed bit for bit by cbq_expand_hits() on the host.
 *   code bits 0..7   material
 *        bits 8..13  normal, two bits per axis (x: 8-9, y: 10-11, z: 12-13): bit 0 = the component is +-1 rather
 *                    than +-0, bit 1 = its sign bit (quirk Q3: zero components are signed, raytracing.cpp:317-318)
 *        bit  14     hit
 *        bit  15     CBQ_HIT_ABANDONED */
typedef struct cbq_hit_compact { float distance; uint32_t code; } cbq_hit_compact;            /* 8 bytes */

/* struct SubDAG verbatim (src/library/raytracing.h:57-65; GLSL mirror glsl/pathtracing.frag:118-126). */
typedef struct cbq_subdag {
	int32_t  lower[3];
	int32_t  height;
	uint32_t pad0;
	uint32_t node;
	uint32_t pad1, pad2;
} cbq_subdag;                                                                       /* 32 bytes */

/* class Camera (src/application/commands/view/camera.h:8-27) after forward()/right()/up() and
 * the fov scale have been evaluated on the host (camera.cpp:24,40-66: libm sin/cos/tan are not
 * bit-portable to the device; everything per-pixel in camera.cpp:19-35 is, and runs on the GPU). */
typedef struct cbq_camera {
	double position[3];
	double forward[3];
	double up[3];
	double right[3];
	float  scale;
	float  pad;
} cbq_camera;

/* The knobs of PathtracingDemo (src/application/commands/view/pathtracing_demo.h:80-84). */
typedef struct cbq_pt_params {
	uint32_t width, height;
	uint32_t spp;              /* samples accumulated per pixel by this call */
	uint32_t bounces;          /* pathtracing_demo.h:80 (used by variant 1 only, like the reference) */
	uint32_t variant;          /* 0: traceSingleRay (pathtracing_demo.cpp:150-190); 1: traceSingleRayRecurse (:120-148) */
	uint32_t include_sun, include_sky, add_noise;
	float    max_footprint;    /* 0.0035 in the reference; -1 disables LOD */
	uint32_t frame_id;         /* index of the first sample: sample s is seeded with frame_id + s */
	uint32_t x0, y0, x1, y1;   /* pixel rectangle [x0,x1) x [y0,y1) this call renders (tile sharding) */
	uint32_t band_count;       /* 0 or 1: every row of the rectangle. N > 1: the rectangle's rows are cut into */
	uint32_t band_index;       /* 64-row bands (the GLSL renderer's tile size, glsl/pathtracing.frag:789-803) and
	                              this call renders bands b with b % band_count == band_index -- one call per GPU
	                              for a round-robin tile-sharded frame */
	uint32_t tile_group_count; /* 0 or 1: every tile. N > 1: only the 64x64-pixel tiles whose fmix32(tile id) % N ==     */
	uint32_t tile_group_index; /* tile_group_index are rendered -- the GLSL viewer's progressive schedule (tile id =
	                              (tile x << 16) | tile y, glsl/pathtracing.frag:792-799; tile rows counted from the top
	                              of the image). Needs the rectangle to be the whole image and band_count <= 1. */
} cbq_pt_params;

#define CBQ_VARIANT_ONE_BOUNCE 0u
#define CBQ_VARIANT_RECURSIVE  1u

/* cbq_trace flags */
#define CBQ_TRACE_SURFACE 1u   /* computeSurfaceProperties = true (raytracing.h:75) */

#define CBQ_MAX_FOOTPRINT_DISABLED (-1.0f)   /* MAX_FOOTPRINT_DISABLED, raytracing.h:71 */

typedef struct cbq_context cbq_context;

/* ---- lifetime ------------------------------------------------------------------------ */

int  cbq_create(int device, cbq_context** out);
void cbq_destroy(cbq_context* ctx);
const char* cbq_last_error(void);              /* thread-local, never NULL */
int  cbq_device_count(void);                   /* 0 when no CUDA device / driver is present */
int  cbq_synchronize(cbq_context* ctx);

/* ---- DAG -> GPU serialisation ------------------------------------------- */

/* Replaces the one-off glBufferData upload of gpu_pathtracing_viewer.cpp:43-67.
 * nodes      : NodeStore::rawBytesPtr() (storage.h:101) -- node_count x 8 u32 INCLUDING the 256
 *              material nodes.
 * root_index : Internals::getRootNodeIndex(volume) (storage.cpp:573-576).
 * colours_rgb: 256 x 3 floats (Viewer::colours(), viewer.cpp:50-55) or NULL for all-purple.
 * The eight sub-DAGs (findSubDAGs, raytracing.cpp:89-97) are computed here and the whole lot is
 * flattened into ONE linear device buffer: [header | subDAG[8] | colours | nodes...]. */
int cbq_upload(cbq_context* ctx, const uint32_t* nodes, uint64_t node_count, uint32_t root_index,
               const float* colours_rgb);

/* cbq_upload for a node array that is already in device memory (e.g. landed there by an NCCL broadcast): one
 * device-to-device copy, sub-DAGs and the child-index check computed on the device. `stream` is the stream the
 * array was produced on (waited for), NULL = none. colours_rgb is a HOST pointer as in cbq_upload. */
int cbq_upload_device(cbq_context* ctx, const uint32_t* d_nodes, uint64_t node_count, uint32_t root_index,
                      const float* colours_rgb, void* stream);

/* Delta re-upload after a runtime edit (Viewer::onMouseButtonDown, viewer.cpp:152-172:
 * checkpoint -> fillBrush -> onVolumeModified). Copy-on-write (NodeStore::setNodeChild,
 * storage.cpp:152-167) only ever touches nodes at or above sharedNodesEnd(), so the device copy
 * is stale exactly on the tail [dirty_begin, node_count): pass dirty_begin = the
 * sharedNodesEnd() seen at the previous sync (or anything lower), the CURRENT base pointer and
 * node count, and the new root. Sub-DAGs are recomputed (pathtracing_demo.cpp:335-341).
 * After Volume::bake() everything moved: call cbq_upload again. Once the device copy has been changed ON the device
 * (cbq_fill_sphere, cbq_bake, cbq_build_dense) no host array is a delta of it any more: cbq_update then fails until
 * the next cbq_upload. */
int cbq_update(cbq_context* ctx, const uint32_t* nodes, uint64_t dirty_begin, uint64_t node_count,
               uint32_t root_index);

/* cbq_update with the dirty tail in device memory: d_tail holds nodes [dirty_begin, node_count) -- the tail only,
 * which is what a broadcast after an edit ships -- and is copied device to device. */
int cbq_update_device(cbq_context* ctx, const uint32_t* d_tail, uint64_t dirty_begin, uint64_t node_count,
                      uint32_t root_index, void* stream);

/* Volume::bake() (reference src/library/storage.cpp:388-395 -> NodeStore::merge / merge_node, :208-290) on the
 * DEVICE copy, without a host round trip: everything the root reaches is hash-consed bottom-up, a node whose eight
 * merged children are one material becomes that material (isMaterialNode(const Node&), :69-75), unreachable nodes
 * (undo history, copy-on-write garbage) are dropped and the children are renumbered. The result is the same
 * canonical minimal DAG the reference produces -- same node count, same voxels, isomorphic -- but not the same
 * array ORDER (the reference's order is its hash-table slot order; here distinct nodes keep the relative order
 * of their first occurrence in the input, so the output is deterministic and baking twice is the identity).
 * The device volume is replaced in place, sub-DAGs are refreshed on the device. *node_count (including the 256
 * material nodes) and *root_index describe the new array; read it back with cbq_download_nodes. Host-side
 * editors (cbq_editable) holding the old array must be re-created from the download: every index moved. */
int cbq_bake(cbq_context* ctx, uint64_t* node_count, uint32_t* root_index);

/* Build the volume from a dense grid of material ids ON THE DEVICE: what the reference does one voxel at a time
 * with Volume::setVoxel (storage.cpp:396-438) followed by Volume::bake (:388-395). voxels[(z * S + y) * S + x],
 * S = 2^size_log2 (2 <= size_log2 <= 12: up to 4096^3; past 1024^3 the grid is built in bricks of 512^3, same result), is the material of voxel origin + (x, y, z); origin must be a multiple
 * of S / 2 per axis (so a grid centred on 0 is fine); everything outside the grid is empty (material 0). The complete octree over the grid is written
 * level by level (no per-voxel inserts), chained up to the height-32 root and hash-consed with the same kernels
 * as cbq_bake; the result -- the canonical DAG, i.e. the same node count and content as the reference's bake of
 * the same voxels -- becomes the context's volume (read it back with cbq_download_nodes). colours_rgb as in
 * cbq_upload. The _device variant takes a device pointer to the grid. */
int cbq_build_dense(cbq_context* ctx, const uint8_t* voxels, uint32_t size_log2, const int32_t origin[3],
                    const float* colours_rgb, uint64_t* node_count, uint32_t* root_index);
int cbq_build_dense_device(cbq_context* ctx, const uint8_t* d_voxels, uint32_t size_log2, const int32_t origin[3],
                           const float* colours_rgb, uint64_t* node_count, uint32_t* root_index);

/* ---- mesh voxeliser ---- */

/* What Mesh::build decides about a triangle list (src/library/voxelization.cpp:765-823): bounds, and from 100 winding-number
 * samples whether the mesh is closed and whether it is inside-out. Host only. */
typedef struct cbq_mesh_info { float lower[3], upper[3]; uint32_t is_closed, is_inside_out; } cbq_mesh_info;
int cbq_mesh_analyse(const float* triangles, uint64_t triangle_count, cbq_mesh_info* info);

/* voxelize(volume, mesh, fill, background) (voxelization.cpp:692-744, with Mesh::build :765-823) on the device, into a grid of
 * S = 2^size_log2 voxels a side at `origin` (a multiple of S / 2 per axis, as for cbq_build_dense; 2 <= size_log2 <= 11: the work space is 5 bytes per voxel) that
 * becomes the context's volume, hash-consed like Volume::bake. triangles: triangle_count x 9 floats in voxel coordinates, user
 * order (later triangles win where several touch a voxel); materials: one id per triangle; thin = Mesh::isThin. A closed mesh is
 * filled: 6-separating shell by the topological test, every octree leaf next to or between the shells classified by the
 * generalized winding number at its centre (|w| > 0.501), surface materials from the triangles within distance 1. An open mesh
 * yields the shell only. background must be 0 (everything outside the grid is empty) and the mesh, dilated by 2 voxels, must fit
 * the grid; an inside-out mesh is refused (flip it: the reference would swap fill and background, i.e. fill the universe).
 * *info (nullable) receives Mesh::build's verdict. Voxel-for-voxel equal to the reference on the test meshes; the reference sums
 * the winding number through a patch hierarchy, here it is the flat sum -- the same number up to float rounding, which the
 * threshold's margin is there to absorb (voxelization.cpp:362-366). */
int cbq_voxelize(cbq_context* ctx, const float* triangles, const uint8_t* materials, uint64_t triangle_count,
                 uint8_t fill, uint8_t background, int thin, uint32_t size_log2, const int32_t origin[3],
                 const float* colours_rgb, cbq_mesh_info* info, uint64_t* node_count, uint32_t* root_index);

/* The reference's runtime edit -- Volume::checkpoint() + fillBrush(volume, SphereBrush(centre, radius), material)
 * (viewer.cpp:165-168; voxelization.cpp:825-915, voxelization.h:91-127) -- applied to the DEVICE copy: no host
 * editor, no PCIe transfer of a dirty tail. Same voxels change as in the reference (same contains() arithmetic and
 * box-overlap pruning); new nodes are appended copy-on-write, nothing that existed is written, so every earlier
 * root stays valid: *root_index is the new root, and cbq_set_root switches between roots (undo / redo,
 * storage.cpp:373-385). The appended tail is not word-for-word the reference's (breadth-first, copies made before
 * it is known that something below changes); the volume and its canonical DAG after a bake are. */
int cbq_fill_sphere(cbq_context* ctx, float x, float y, float z, float radius, uint8_t material,
                    uint32_t* root_index, uint64_t* node_count);
int cbq_set_root(cbq_context* ctx, uint32_t root_index);

int cbq_set_colours(cbq_context* ctx, const float* colours_rgb);
int cbq_get_subdags(cbq_context* ctx, cbq_subdag out[8]);

/* Host-only restatement of findSubDAGs for callers that want it without a device. */
int cbq_find_subdags(const uint32_t* nodes, uint64_t node_count, uint32_t root_index, cbq_subdag out[8]);

/* Read nodes [begin, begin + count) back from the device copy (testing / verification). */
int cbq_download_nodes(cbq_context* ctx, uint64_t begin, uint64_t count, uint32_t* out);
int cbq_node_count(cbq_context* ctx, uint64_t* out);

/* ---- ray cast: intersectVolume over a batch -------------------------------------------- */

/* hits[i] = intersectVolume(volume, subDAGs, rays[i]..., flags & CBQ_TRACE_SURFACE, max_footprint)
 * (raytracing.cpp:397-478) for every i. Host buffers; pinned memory (cbq_host_alloc) makes the
 * copies asynchronous and overlapped with the kernel. */
int cbq_trace(cbq_context* ctx, const cbq_ray* rays, uint64_t n, uint32_t flags, float max_footprint,
              cbq_hit* hits);

/* Same with device pointers, enqueued on `stream` (a cudaStream_t). NULL selects the context's own
 * non-blocking stream; to order the work on the CUDA default stream pass cudaStreamLegacy. */
int cbq_trace_device(cbq_context* ctx, const cbq_ray* d_rays, uint64_t n, uint32_t flags,
                     float max_footprint, cbq_hit* d_hits, void* stream);

/* cbq_trace with 8-byte results: 24 + 8 bytes per ray cross PCIe instead of 24 + 40. cbq_expand_hits (host only,
 * `threads` worker threads, 0 = all cores) widens them into exactly the records cbq_trace writes for the same rays. */
int cbq_trace_compact(cbq_context* ctx, const cbq_ray* rays, uint64_t n, uint32_t flags, float max_footprint,
                      cbq_hit_compact* hits);
int cbq_trace_compact_device(cbq_context* ctx, const cbq_ray* d_rays, uint64_t n, uint32_t flags,
                             float max_footprint, cbq_hit_compact* d_hits, void* stream);
int cbq_expand_hits(const cbq_ray* rays, const cbq_hit_compact* compact, uint64_t n, cbq_hit* hits, int threads);

/* Camera::rayFromViewportPos (camera.cpp:12-38) for every pixel, row-major, cast to float like
 * static_cast<Ray3f> in PathtracingDemo::raytrace (pathtracing_demo.cpp:220). */
int cbq_camera_from_pose(const double position[3], double pitch, double yaw, double fov_degrees,
                         cbq_camera* out);
int cbq_primary_rays_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height,
                            cbq_ray* d_rays, void* stream);

/* The same rays in 8x4-pixel tile order (width % 8 == 0, height % 4 == 0): ray i is pixel (i % 32) of tile
 * (i / 32), so every 32 consecutive rays -- what one warp claims from a buffer -- are one compact tile
 * (Morton-like locality for primary rays). d_pixel_of (nullable) receives y * width + x of each ray. */
int cbq_primary_rays_tiled_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height,
                                  cbq_ray* d_rays, uint32_t* d_pixel_of, void* stream);

/* Synthetic collision-query rays generated on the device (BASELINE config 3): origin uniform in
 * [lower, upper), direction uniform on the sphere, from a counter-based hash of (seed, ray index), so
 * ray i is the same whatever the batch it is generated in. cubiquity_b200/rays.py:counter_rays is the
 * identical host generator. */
int cbq_random_rays_device(cbq_context* ctx, uint64_t seed, const float lower[3], const float upper[3],
                           uint64_t n, cbq_ray* d_rays, void* stream);

/* Fused: generate the primary ray of every pixel (8x4-pixel tiles per warp for coherence) and
 * trace it; d_hits is row-major width x height. */
int cbq_raycast_frame_device(cbq_context* ctx, const cbq_camera* cam, uint32_t width, uint32_t height,
                             uint32_t flags, float max_footprint, cbq_hit* d_hits, void* stream);

/* ---- path tracer: PathtracingDemo::raytrace -------------------------------------------- */

/* accum (width x height x 3 floats, row-major) has p->spp samples ADDED to every pixel of the
 * rectangle, like `mImage[...] += pixel` (pathtracing_demo.cpp:224). Per-pixel RNG stream:
 * seed = hashRay(primary ray) ^ fmix32(frame_id + s) (the GLSL tracer's rule,
 * glsl/pathtracing.frag:770-780,816), advanced by (u32)bit_mix (pathtracing_demo.cpp:67). */
int cbq_render(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, float* accum);
int cbq_render_device(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, float* d_accum,
                      void* stream);

/* ---- the viewer's screen-space passes (GPUPathtracingViewer::onUpdate, gpu_pathtracing_viewer.cpp:121-222) ---- */

/* One progressive frame: `frame` is the viewer's frameId. The 64x64-pixel tiles of group frame % 16 (glsl/pathtracing.frag:
 * 789-803, groupCount 16) receive p->spp more samples (sample s is seeded with frame * spp + s), added into d_rgba --
 * width x height x 4 floats, alpha counting the samples of each pixel, as the viewer's additive blending into its RGBA32F
 * target does (gpu_pathtracing_viewer.cpp:152-161). p->frame_id, rectangle, bands and tile group fields are ignored. */
int cbq_progressive_pass_device(cbq_context* ctx, const cbq_camera* cam, const cbq_pt_params* p, uint32_t frame,
                                float* d_rgba, void* stream);
/* glsl/normalise.frag: rgb = rgba.rgb / rgba.a (a pixel that has no sample yet comes out 0/0 = NaN, as in the shader). */
int cbq_normalise_device(cbq_context* ctx, const float* d_rgba, uint32_t width, uint32_t height, float* d_rgb, void* stream);
/* glsl/horz_blur.frag then glsl/vert_blur.frag, `passes` times, in place (d_scratch_rgba is the viewer's blurTexture): 9 taps
 * along the axis, a tap counts only if its alpha differs from the centre's by less than 0.5, coordinates wrap (GL_REPEAT). The
 * viewer has these passes compiled but switched off (`blurPasses = 0`, gpu_pathtracing_viewer.cpp:168-192). */
int cbq_blur_device(cbq_context* ctx, float* d_rgba, uint32_t width, uint32_t height, float* d_scratch_rgba, int passes,
                    void* stream);

/* ---- .dag files (Volume::load / Volume::save, src/library/storage.cpp:505-542, NodeStore::read / write :192-206) ---- */

/* The file is u32 rootIndex, u32 nodeCount, nodeCount x 32 bytes (the non-material nodes; indices in the file already count the
 * 256 material nodes). The reference reads it unchecked and sizes it with 32-bit byte counts (storage.h:102); these do the
 * arithmetic in 64 bits and refuse a file whose length is not exactly 8 + 32 x nodeCount, whose root or any child index is past
 * the end, or that cannot be read in full. cbq_dag_load returns the array INCLUDING the 256 material nodes (what cbq_upload
 * takes), malloc'ed: release it with cbq_dag_free. cbq_dag_save writes next to the target and renames. */
int  cbq_dag_load(const char* path, uint32_t** nodes, uint64_t* node_count, uint32_t* root_index);
void cbq_dag_free(uint32_t* nodes);
int  cbq_dag_save(const char* path, const uint32_t* nodes, uint64_t node_count, uint32_t root_index);
int  cbq_upload_dag(cbq_context* ctx, const char* path, const float* colours_rgb);

/* Messages of failed calls also go to this callback (same shape as the reference's MessageHandlerPtr, base.h:102-107);
 * NULL switches it off. Process-wide. */
void cbq_set_log_callback(void (*handler)(const char* message));

/* Self-test hook: `draws` successive randomPointInUnitSphere results (pathtracing_demo.cpp:62-79) of the
 * stream that starts at each seed: d_points is n x draws x 3 floats, d_states the n final states. */
int cbq_rng_points_device(cbq_context* ctx, const uint32_t* d_seeds, uint64_t n, int draws, float* d_points,
                          uint32_t* d_states, void* stream);

/* ---- multi-GPU: results gathered over peer memory ----------------------------------- */

/* The reference is single-GPU; this is the one exchange the partitioned path needs: every GPU's results end up in ONE
 * buffer on one GPU. That buffer is allocated here (cudaMalloc, zero-filled) and exported as a CUDA IPC handle; the
 * other processes (one per GPU) open it and get a device pointer that addresses the owner's HBM over NVLink. A rank
 * then either passes that pointer (+ its slice offset) as the result buffer of cbq_trace_compact_device -- the ray-cast
 * kernel's stores ARE the transfer -- or ships finished chunks with cbq_copy_device (copy engines, no SM involved). */
typedef struct cbq_ipc_handle { unsigned char bytes[64]; } cbq_ipc_handle;
int cbq_shared_alloc(cbq_context* ctx, uint64_t bytes, void** d_ptr, cbq_ipc_handle* handle);
int cbq_shared_open(cbq_context* ctx, const cbq_ipc_handle* handle, uint64_t bytes, void** d_ptr);   /* bytes: as allocated */
int cbq_shared_close(cbq_context* ctx, void* d_ptr);     /* a pointer from cbq_shared_open */
int cbq_shared_free(cbq_context* ctx, void* d_ptr);      /* a pointer from cbq_shared_alloc */
int cbq_copy_device(cbq_context* ctx, void* d_dst, const void* d_src, uint64_t bytes, void* stream);

/* ---- pinned host memory, tuning, counters ---------------------------------------------- */

int cbq_host_alloc(void** out, uint64_t bytes);
int cbq_host_free(void* p);

/* Options: "block_threads", "blocks_per_sm", "refill_threshold" (idle lanes of a warp that trigger a
 * mid-flight refill, 1..32), "l2_persist" (0/1), "sample_group" (samples of a
 * pixel the wavefront tracer traces together, 1..16; 0 = choose from the size of the rectangle),
 * "pt_refill_threshold" (the refill threshold of the path tracer's shadow- and bounce-ray casts, 1..32, default 4),
 * "dense_brick_log2" (cbq_build_dense: 0 = one piece up to 1024^3 and bricks of 512^3 beyond; 3..10 = always bricks of that size),
 * "adaptive_order" (0/1, default 1: coherent batches -- refill_threshold 32, cbq_raycast_frame_device -- record how
 * long each 32-ray ticket took, and the next launch over the same ray buffer, size and stream deals the tickets
 * longest first; a scheduling hint only, results do not depend on it), "park_results" (0/1, default 0: cbq_trace_compact_device already coalesces a warp's result stores when the result
 * buffer lies in another GPU's memory, see cbq_shared_open; 1 forces that path for local buffers too -- a testing aid),
 * "order_refresh" (that order is rebuilt from
 * the recorded costs every this many launches, default 4; cbq_raycast_frame_device also rebuilds it after every
 * frame whose camera differs from the previous one). */
int cbq_set_option(cbq_context* ctx, const char* key, int64_t value);
int cbq_get_option(cbq_context* ctx, const char* key, int64_t* value);

/* Counters: "kernel_launches", "rays_traced", "bytes_h2d", "bytes_d2h", "abandoned_rays". */
int cbq_get_counter(cbq_context* ctx, const char* key, uint64_t* value);
int cbq_reset_counters(cbq_context* ctx);

/* ---- host-side copy-on-write edits (the viewer's pick -> checkpoint -> brush -> re-sync loop) ---- */

/* A growable copy of a node array with the reference's edit history. Restates Volume::checkpoint /
 * undo / redo (storage.cpp:358-385), NodeStore::setNodeChild (storage.cpp:152-167) and fillBrush with a
 * SphereBrush (voxelization.cpp:825-915) so that the resulting array is word for word the reference's. */
typedef struct cbq_editable cbq_editable;
int  cbq_editable_create(const uint32_t* nodes, uint64_t node_count, uint32_t root_index, cbq_editable** out);
void cbq_editable_destroy(cbq_editable* e);
int  cbq_editable_checkpoint(cbq_editable* e);
int  cbq_editable_undo(cbq_editable* e);
int  cbq_editable_redo(cbq_editable* e);
int  cbq_editable_fill_sphere(cbq_editable* e, float x, float y, float z, float radius, uint8_t material);
const uint32_t* cbq_editable_nodes(const cbq_editable* e, uint64_t* node_count);
uint32_t cbq_editable_root(const cbq_editable* e);
uint64_t cbq_editable_shared_end(const cbq_editable* e);
/* first_upload != 0: cbq_upload; otherwise cbq_update with the tail that changed since the last sync. */
int  cbq_editable_sync(cbq_editable* e, cbq_context* ctx, int first_upload, const float* colours_rgb);

#ifdef __cplusplus
}
#endif
#endif /* CUBIQUITY_B200_H */

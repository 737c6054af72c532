/* Procedural SVDAG scenes: the synthetic inputs of the tests and benchmarks (BASELINE.json "synthetic procedural
 * volumes of the named resolutions"). Host only, its own shared library (scenes/lib/libcbq_scenes.so): NOT part of the
 * product library, so that the reference arm of bench.py can build its input without mapping product code. */
#ifndef CBQ_SCENES_H
#define CBQ_SCENES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CBQ_OK = 0, CBQ_ERROR_INVALID_ARGUMENT = 1, CBQ_ERROR_OUT_OF_MEMORY = 4 };   /* same values as cubiquity_b200.h */

typedef struct cbq_scene cbq_scene;
/* kind: "sphere_noise" | "terrain" | "soup" | "city". The volume is 2^size_log2 voxels a side. The node array is the
 * reference's NodeStore::rawBytesPtr() layout (src/library/storage.h:101): 8 x u32 per node incl. the 256 material nodes. */
int cbq_scene_build(const char* kind, uint32_t size_log2, uint64_t seed, cbq_scene** out);
const uint32_t* cbq_scene_nodes(const cbq_scene* s, uint64_t* node_count);
uint32_t cbq_scene_root(const cbq_scene* s);
void cbq_scene_bounds(const cbq_scene* s, int32_t lower[3], int32_t upper[3]);
void cbq_scene_colours(const cbq_scene* s, float* rgb768);
void cbq_scene_voxels(const cbq_scene* s, const int32_t* xyz, uint64_t n, uint8_t* out);
void cbq_scene_free(cbq_scene* s);

#ifdef __cplusplus
}
#endif
#endif

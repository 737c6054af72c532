// Host-side copy-on-write edits of a node array -- just enough of the reference's mutation path to
// drive the runtime-edit protocol (pick -> checkpoint -> sphere brush -> delta re-upload) without the
// reference library: BASELINE config 5 ("runtime voxel edits and delta re-upload between frames").
//
// Follows, with the same arithmetic and the same visiting order so that the resulting node ARRAY is
// identical to the reference's (tests/test_edit.py compares them word for word):
//   NodeStore::getNodeChild / setNodeChild / cloneRoot   reference src/library/storage.cpp:143-167,298-303
//   Volume::checkpoint / undo / redo                      storage.cpp:358-385
//   fillBrush + SphereBrush::contains / bounds            src/library/voxelization.cpp:825-915, voxelization.h:91-127
//   overlaps(Box3f, Box3f)                                src/library/geometry.h:347-357
// Merging (bake) is NOT here: after a bake every node moves and the volume is uploaded afresh.
#include "../../include/cubiquity_b200.h"

#include <array>
#include <cstdint>
#include <new>
#include <vector>

namespace {

typedef std::array<uint32_t, 8> Node;

struct Sphere {
	float cx, cy, cz, radiusSquared;
	float lo[3], hi[3];
	bool contains(float x, float y, float z) const
	{
		const float dx = x - cx, dy = y - cy, dz = z - cz;
		const int64_t distSq = (int64_t)(dx * dx + dy * dy + dz * dz);   // truncation, as the reference does
		return (float)distSq < radiusSquared;
	}
};

} // namespace

struct cbq_editable {
	std::vector<Node> nodes;
	uint32_t sharedEnd = 0;           // nodes below this index may have several parents: copy before writing
	std::vector<uint32_t> roots;      // edit history (undo / redo)
	int current = 0;
	uint64_t syncedSharedEnd = 0;     // sharedEnd when a device copy was last brought up to date

	uint32_t root() const { return roots[(size_t)current]; }

	uint32_t child(uint32_t node, uint32_t slot) const { return node < 256u ? node : nodes[node][slot]; }

	uint32_t setChild(uint32_t node, uint32_t slot, uint32_t value)
	{
		if (node < sharedEnd) {            // includes the material nodes
			nodes.push_back(nodes[node]);
			node = (uint32_t)(nodes.size() - 1);
		}
		nodes[node][slot] = value;
		return node;
	}

	uint32_t fill(const Sphere& b, uint8_t mat, uint32_t node, int height, int32_t lx, int32_t ly, int32_t lz)
	{
		const uint32_t childHeight = (uint32_t)height - 1u;
		const uint32_t side = 1u << childHeight;
		for (uint32_t cz = 0; cz <= 1; cz++) for (uint32_t cy = 0; cy <= 1; cy++) for (uint32_t cx = 0; cx <= 1; cx++) {
			const uint32_t slot = cz << 2 | cy << 1 | cx;
			const int32_t x0 = (int32_t)((uint32_t)lx + side * cx), y0 = (int32_t)((uint32_t)ly + side * cy), z0 = (int32_t)((uint32_t)lz + side * cz);
			const int32_t x1 = (int32_t)((uint32_t)x0 + (side - 1u)), y1 = (int32_t)((uint32_t)y0 + (side - 1u)), z1 = (int32_t)((uint32_t)z0 + (side - 1u));
			const float fl[3] = { (float)x0, (float)y0, (float)z0 }, fu[3] = { (float)x1, (float)y1, (float)z1 };
			bool apart = false;
			for (int a = 0; a < 3; a++) if (b.hi[a] < fl[a] || b.lo[a] > fu[a]) apart = true;
			if (apart) continue;
			bool inside = true;
			for (int k = 0; k < 8; k++) {
				if (!b.contains((k & 4) ? fu[0] : fl[0], (k & 2) ? fu[1] : fl[1], (k & 1) ? fu[2] : fl[2])) inside = false;
			}
			const uint32_t old = child(node, slot);
			if (old == mat) continue;
			uint32_t fresh = old;
			if (height >= 2 && !inside) fresh = fill(b, mat, old, height - 1, x0, y0, z0);
			else if (inside) fresh = mat;
			if (fresh != old) node = setChild(node, slot, fresh);
		}
		return node;
	}
};

extern "C" {

int cbq_editable_create(const uint32_t* nodes, uint64_t node_count, uint32_t root_index, cbq_editable** out)
{
	if (!nodes || !out || node_count < 256 || node_count > 0xffffffffull || root_index >= node_count) return CBQ_ERROR_INVALID_ARGUMENT;
	cbq_editable* e = new (std::nothrow) cbq_editable();
	if (!e) return CBQ_ERROR_OUT_OF_MEMORY;
	try {
		e->nodes.resize((size_t)node_count);
		for (uint64_t i = 0; i < node_count; i++) for (int c = 0; c < 8; c++) e->nodes[(size_t)i][c] = nodes[i * 8 + c];
	} catch (const std::bad_alloc&) { delete e; return CBQ_ERROR_OUT_OF_MEMORY; }
	e->sharedEnd = (uint32_t)node_count;          // like a freshly loaded volume (storage.cpp:198)
	e->roots.assign(1, root_index);
	e->current = 0;
	e->syncedSharedEnd = node_count;
	*out = e;
	return CBQ_OK;
}

void cbq_editable_destroy(cbq_editable* e) { delete e; }

int cbq_editable_checkpoint(cbq_editable* e)
{
	if (!e) return CBQ_ERROR_INVALID_ARGUMENT;
	e->roots.resize((size_t)e->current + 1);               // drop any redo entries
	e->sharedEnd = (uint32_t)e->nodes.size();              // everything so far becomes potentially shared
	e->nodes.push_back(e->nodes[e->root()]);               // unshared copy of the root
	e->roots.push_back((uint32_t)(e->nodes.size() - 1));
	e->current++;
	return CBQ_OK;
}

int cbq_editable_undo(cbq_editable* e) { if (!e) return CBQ_ERROR_INVALID_ARGUMENT; if (e->current > 0) e->current--; return CBQ_OK; }
int cbq_editable_redo(cbq_editable* e) { if (!e) return CBQ_ERROR_INVALID_ARGUMENT; if ((size_t)e->current + 1 < e->roots.size()) e->current++; return CBQ_OK; }

int cbq_editable_fill_sphere(cbq_editable* e, float x, float y, float z, float radius, uint8_t material)
{
	if (!e) return CBQ_ERROR_INVALID_ARGUMENT;
	Sphere b;
	b.cx = x; b.cy = y; b.cz = z; b.radiusSquared = radius * radius;
	b.lo[0] = x - radius; b.lo[1] = y - radius; b.lo[2] = z - radius;
	b.hi[0] = x + radius; b.hi[1] = y + radius; b.hi[2] = z + radius;
	try {
		const int32_t lowest = INT32_MIN;
		e->roots[(size_t)e->current] = e->fill(b, material, e->root(), 32, lowest, lowest, lowest);
	} catch (const std::bad_alloc&) { return CBQ_ERROR_OUT_OF_MEMORY; }
	return CBQ_OK;
}

const uint32_t* cbq_editable_nodes(const cbq_editable* e, uint64_t* node_count)
{
	if (node_count) *node_count = e->nodes.size();
	return reinterpret_cast<const uint32_t*>(e->nodes.data());
}

uint32_t cbq_editable_root(const cbq_editable* e) { return e->root(); }
uint64_t cbq_editable_shared_end(const cbq_editable* e) { return e->sharedEnd; }

// Bring a device copy up to date: the first call uploads, later calls ship the dirty tail only.
int cbq_editable_sync(cbq_editable* e, cbq_context* ctx, int first_upload, const float* colours_rgb)
{
	if (!e || !ctx) return CBQ_ERROR_INVALID_ARGUMENT;
	uint64_t n = 0;
	const uint32_t* p = cbq_editable_nodes(e, &n);
	int rc;
	if (first_upload) rc = cbq_upload(ctx, p, n, e->root(), colours_rgb);
	else rc = cbq_update(ctx, p, e->syncedSharedEnd < n ? e->syncedSharedEnd : n, n, e->root());
	if (rc == CBQ_OK) e->syncedSharedEnd = e->sharedEnd;
	return rc;
}

} // extern "C"

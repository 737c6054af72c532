"""Reader / writer for the reference's `.dag` volume file.

Layout (reference src/library/storage.cpp:192-206, 505-542):
    u32 rootIndex            absolute index into the node array (already includes the 256 material nodes)
    u32 nodeCount            number of non-material nodes that follow
    nodeCount x 8 x u32      entries [256, 256 + nodeCount) of the node array
The 256 self-referencing material nodes (storage.cpp:110-122) are not stored; `read_dag` recreates
them so the returned array is exactly what `NodeStore::rawBytesPtr()` exposes (storage.h:101).
"""
import numpy as np

MATERIAL_COUNT = 256


def material_nodes():
    return np.repeat(np.arange(MATERIAL_COUNT, dtype=np.uint32)[:, None], 8, axis=1)


def write_dag(path, nodes, root):
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8)
    if len(nodes) < MATERIAL_COUNT:
        raise ValueError("node array must include the 256 material nodes")
    body = nodes[MATERIAL_COUNT:]
    with open(path, "wb") as f:
        np.array([root, len(body)], dtype="<u4").tofile(f)
        body.astype("<u4").tofile(f)


def read_dag(path):
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype="<u4", count=2)
        if len(head) != 2:
            raise IOError("truncated .dag header: %s" % path)
        root, count = int(head[0]), int(head[1])
        body = np.fromfile(f, dtype="<u4", count=count * 8)
        if len(body) != count * 8:
            raise IOError("truncated .dag body: %s" % path)
    nodes = np.concatenate([material_nodes(), body.reshape(-1, 8)])
    return nodes, root

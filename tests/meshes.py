"""Small procedural triangle meshes for the voxeliser tests: (n, 9) float32 triangles + (n,) uint8 materials."""
import numpy as np


def icosphere(centre, radius, subdivisions=2, material=3):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    tris = v[np.array(f)]
    for _ in range(subdivisions):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    tris = tris / np.linalg.norm(tris, axis=2, keepdims=True) * radius + np.asarray(centre, dtype=np.float64)
    return tris.reshape(-1, 9).astype(np.float32), np.full(len(tris), material, dtype=np.uint8)


def box(lower, upper, material=5):
    x0, y0, z0 = lower
    x1, y1, z1 = upper
    p = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], dtype=np.float64)
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]     # outward normals
    tris = []
    for q in quads:
        tris.append([p[q[0]], p[q[1]], p[q[2]]])
        tris.append([p[q[0]], p[q[2]], p[q[3]]])
    return np.array(tris).reshape(-1, 9).astype(np.float32), np.full(12, material, dtype=np.uint8)


def flipped(tris):
    t = tris.reshape(-1, 3, 3).copy()
    t[:, [1, 2]] = t[:, [2, 1]]
    return t.reshape(-1, 9)


def soup(offset=(0.0, 0.0, 0.0)):
    """Two spheres, a long box (its faces are split by drawLargeTriangle) and a small box with integer corners."""
    o = np.asarray(offset, dtype=np.float64)
    parts = [icosphere(o + [40.3, 44.1, 50.7], 21.4, 2, 3), icosphere(o + [78.2, 70.9, 60.2], 17.8, 2, 7),
             box(o + [20.25, 80.5, 30.75], o + [100.6, 95.1, 41.2], 5), box(o + [60, 20, 20], o + [76, 36, 44], 9)]
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

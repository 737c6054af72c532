"""Multi-GPU plumbing: one process per GPU, DAG replicated, work partitioned, results gathered.

The path shards without any exchange inside it: rays, pixels and samples are
independent and the DAG is read-only during a frame. So the only collectives are
  * a broadcast of the node array (and, after an edit, of its dirty tail) from rank 0, and
  * one gather / reduce of the finished image or hit buffer.
`torch.distributed` supplies both (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def split_range(n, world_size, rank):
    """Contiguous 1/world_size slice of n items (ray batches)."""
    base, rem = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def tile_rows(height, world_size, rank, tile=64):
    """Rows of 64-pixel-high tile bands dealt round-robin to ranks (the GLSL progressive renderer's tile
    size, reference glsl/pathtracing.frag:789-803). Returns a list of (y0, y1)."""
    bands = [(y, min(y + tile, height)) for y in range(0, height, tile)]
    return [b for i, b in enumerate(bands) if i % world_size == rank]


def broadcast_volume(dist, nodes, root, device=None, src=0):
    """Replicate the DAG: rank `src` passes (nodes, root); the others pass (None, None)."""
    import torch
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        meta[0] = len(nodes)
        meta[1] = int(root)
    dist.broadcast(meta, src=src)
    count, root = int(meta[0].item()), int(meta[1].item())
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.uint32).view(np.int32).reshape(-1))
    else:
        buf = torch.empty(count * 8, dtype=torch.int32)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy().view(np.uint32).reshape(-1, 8), root


def broadcast_tail(dist, nodes, dirty_begin, root, device=None, src=0):
    """After an edit on rank `src`: ship only nodes[dirty_begin:] and the new root.
    Other ranks pass their current (stale) array as `nodes`; returns the refreshed (array, root)."""
    import torch
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        meta[0], meta[1], meta[2] = len(nodes), int(dirty_begin), int(root)
    dist.broadcast(meta, src=src)
    count, dirty_begin, root = (int(v) for v in meta.tolist())
    tail = count - dirty_begin
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(nodes[dirty_begin:], dtype=np.uint32).view(np.int32).reshape(-1))
    else:
        buf = torch.empty(tail * 8, dtype=torch.int32)
    if device is not None:
        buf = buf.to(device)
    if tail:
        dist.broadcast(buf, src=src)
    if dist.get_rank() == src:
        return np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 8), root, dirty_begin
    fresh = np.concatenate([np.asarray(nodes, dtype=np.uint32).reshape(-1, 8)[:dirty_begin],
                            buf.cpu().numpy().view(np.uint32).reshape(-1, 8)])
    return fresh, root, dirty_begin


def reduce_image(dist, accum, dst=0):
    """Disjoint tiles or disjoint sample ranges: a sum onto rank `dst` assembles the frame.
    `accum` is a torch tensor (H, W, 3) float32 holding zeros outside this rank's share."""
    dist.reduce(accum, dst=dst)
    return accum

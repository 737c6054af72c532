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


def replicate_volume(dist, ctx, nodes, root, colours, device, src=0):
    """broadcast_volume straight into the ray caster: the NCCL broadcast lands in device memory and
    cbq_upload_device copies it device to device into the volume buffer (no trip through the host: the
    reference's upload is the one-off glBufferData of gpu_pathtracing_viewer.cpp:46-50). Rank `src` passes the
    host arrays, the others None. Returns (node_count, root, colours)."""
    import torch
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        meta[0], meta[1] = len(nodes), int(root)
    dist.broadcast(meta, src=src)
    count, root = int(meta[0].item()), int(meta[1].item())
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.uint32).view(np.int32).reshape(-1)).to(device)
        col = torch.from_numpy(np.ascontiguousarray(colours, dtype=np.float32).reshape(256, 3).copy()).to(device)
    else:
        buf = torch.empty(count * 8, dtype=torch.int32, device=device)
        col = torch.empty(256, 3, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    dist.broadcast(col, src=src)
    colours = col.cpu().numpy()
    torch.cuda.synchronize(device)
    ctx.upload_device(buf.data_ptr(), count, root, colours)
    return count, root, colours


def replicate_tail(dist, ctx, nodes, dirty_begin, root, device, src=0):
    """broadcast_tail straight into the ray caster: rank `src` passes its current host array, the first dirty node
    and the new root; every rank (src included) applies the tail with cbq_update_device. Returns (node_count, root,
    tail_bytes)."""
    import torch
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        meta[0], meta[1], meta[2] = len(nodes), int(dirty_begin), int(root)
    dist.broadcast(meta, src=src)
    count, dirty_begin, root = (int(v) for v in meta.tolist())
    tail = count - dirty_begin
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(nodes[dirty_begin:], dtype=np.uint32).view(np.int32).reshape(-1)).to(device)
    else:
        buf = torch.empty(tail * 8, dtype=torch.int32, device=device)
    if tail:
        dist.broadcast(buf, src=src)
    torch.cuda.synchronize(device)
    ctx.update_device(buf.data_ptr() if tail else 0, dirty_begin, count, root)
    return count, root, tail * 32


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

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


def job_chunks(units, unit_items, world_size, rank, chunk_units):
    """A job of `units` equal units (frames) of `unit_items` items (rays) each: rank's contiguous share of the units, cut
    into chunks of `chunk_units` units. Returns [(begin, end)] in GLOBAL item indices. A chunk is what one launch traces
    and one message carries to rank 0."""
    u0, u1 = split_range(units, world_size, rank)
    step = max(1, int(chunk_units))
    return [(u * unit_items, min(u + step, u1) * unit_items) for u in range(u0, u1, step)]


def post_receives(dist, out, plans, words_per_item, dst=0):
    """On rank `dst`: the receives of the whole exchange, straight into the job-sized result tensor `out`
    (`words_per_item` elements per item), posted before its own work starts. plans[r] = job_chunks(..., rank=r).
    Round j = chunk j of every other rank, ONE batched (grouped) operation per round, so that the ranks' transfers
    of a round run side by side instead of queueing behind each other. Returns the work handles."""
    works = []
    rounds = max((len(p) for r, p in enumerate(plans) if r != dst), default=0)
    for j in range(rounds):
        ops = []
        for src, plan in enumerate(plans):
            if src != dst and j < len(plan) and plan[j][1] > plan[j][0]:
                b, e = plan[j]
                ops.append(dist.P2POp(dist.irecv, out[b * words_per_item:e * words_per_item], src))
        if ops:
            works += dist.batch_isend_irecv(ops)
    return works


def send_chunk(dist, piece, dst=0):
    """On the other ranks: ship one finished chunk (the matching half of a post_receives round); the transfer runs
    beside the trace of the next chunk."""
    return dist.batch_isend_irecv([dist.P2POp(dist.isend, piece, dst)])


class SharedResults:
    """The job's result buffer: allocated on rank `dst`'s GPU (cbq_shared_alloc), opened by every other rank as a
    device pointer into that GPU's memory (CUDA IPC; the accesses travel over NVLink). `ptr` is the buffer in THIS
    process's address space -- pass ptr + offset as the result pointer of a trace (the kernel's stores are the
    transfer) or as the destination of cbq_copy_device (copy engines)."""

    def __init__(self, dist, ctx, nbytes, device, dst=0):
        import torch
        self.ctx, self.owner = ctx, dist.get_rank() == dst
        h = torch.zeros(64, dtype=torch.uint8, device=device)
        if self.owner:
            self.ptr, handle = ctx.shared_alloc(nbytes)
            h.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
        dist.broadcast(h, src=dst)
        if not self.owner:
            self.ptr = ctx.shared_open(bytes(h.cpu().numpy().tobytes()), nbytes)
        self.nbytes = nbytes

    def close(self, dist):
        dist.barrier()                    # nobody may still be writing when the owner frees
        if self.owner:
            self.ctx.shared_free(self.ptr)
        else:
            self.ctx.shared_close(self.ptr)
        self.ptr = 0


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

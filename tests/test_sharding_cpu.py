"""World-size-2 `gloo` test of the multi-GPU plumbing (cubiquity_b200/sharding.py) on CPU.

No GPU here, so each rank renders its share with the ORACLE (allowed in tests); what is under test is the
replication of the DAG, the delta broadcast after an edit, the work partition and the final reduction."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import torch
    from cubiquity_b200 import api, sharding
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = pyoracle.Port()
    PI_F = float(np.float32(3.14159265358979))

    # rank 0 owns the volume; everyone gets a replica
    sc = api.Scene("sphere_noise", 6, seed=4) if rank == 0 else None
    nodes, root = sharding.broadcast_volume(dist, sc.nodes if rank == 0 else None, sc.root if rank == 0 else None)
    meta = torch.zeros(6, dtype=torch.int64)
    if rank == 0:
        meta[:3] = torch.from_numpy(sc.lower.astype(np.int64))
        meta[3:] = torch.from_numpy(sc.upper.astype(np.int64))
    dist.broadcast(meta, src=0)
    lower, upper = meta[:3].numpy(), meta[3:].numpy()
    colours = np.tile(np.array([[0.7, 0.6, 0.5]], dtype=np.float32), (256, 1))

    cam = api.default_camera(lower, upper)
    ocam = oracle.camera(list(cam.position), -(PI_F / 4.0), 0.0)
    w, h = 96, 80
    sd = oracle.find_subdags(nodes, root)

    # tile-sharded frame: each rank renders its 16-row bands, then one reduce assembles the image
    accum = np.zeros((h, w, 3), dtype=np.float32)
    for y0, y1 in sharding.tile_rows(h, world, rank, tile=16):
        p = pyoracle.PtParams(w, h, 2, 2, 1, 1, 1, 1, 0.0035, 0, 0, y0, w, y1, 0)
        oracle.render(nodes, sd, colours, ocam, p, accum=accum)
    t = torch.from_numpy(accum)
    sharding.reduce_image(dist, t, dst=0)

    # ray batch: contiguous slices, gathered
    rays = oracle.camera_rays(ocam, w, h)
    b, e = sharding.split_range(len(rays), world, rank)
    hits, _, _ = oracle.trace(nodes, sd, rays[b:e], True, -1.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, hits.tobytes())

    # the bench's job exchange: a job of 5 "frames" (16-row strips of the image), contiguous shares, one message per chunk
    # to rank 0 while the next chunk is computed. Results travel as 8-byte compact records.
    from test_host_shim import compact_of
    strip = 16 * w
    plans = [sharding.job_chunks(5, strip, world, r, 1) for r in range(world)]
    job_out = torch.zeros(5 * strip * 2, dtype=torch.int32)
    works = sharding.post_receives(dist, job_out, plans, 2) if rank == 0 else []
    for b0, b1 in plans[rank]:
        piece, _, _ = oracle.trace(nodes, sd, rays[b0:b1], True, -1.0)
        words = torch.from_numpy(compact_of(piece, api).view(np.int32).reshape(-1).copy())
        if rank == 0:
            job_out[b0 * 2:b1 * 2] = words
        else:
            works += sharding.send_chunk(dist, words)
    for wk in works:
        wk.wait()
    # sample-sharded path tracing: every rank renders its share of the samples of the whole frame, one reduce
    s0, s1 = sharding.split_range(4, world, rank)
    part = np.zeros((h, w, 3), dtype=np.float32)
    oracle.render(nodes, sd, colours, ocam, pyoracle.PtParams(w, h, s1 - s0, 2, 1, 1, 1, 1, 0.0035, 10 + s0, 0, 0, w, h, 0), accum=part)
    tp = torch.from_numpy(part)
    sharding.reduce_image(dist, tp, dst=0)

    # an "edit" on rank 0: append a copy of the root with one octant cleared, ship only the tail
    if rank == 0:
        new_root = nodes[root].copy()
        new_root[int(np.nonzero(new_root)[0][0])] = 0
        edited = np.concatenate([nodes, new_root[None, :]])
        fresh, froot, dirty = sharding.broadcast_tail(dist, edited, len(nodes), len(edited) - 1)
    else:
        fresh, froot, dirty = sharding.broadcast_tail(dist, nodes, 0, 0)
    sd2 = oracle.find_subdags(fresh, froot)
    hits2, _, _ = oracle.trace(fresh, sd2, rays[b:e], True, -1.0)
    np.save(os.path.join(out_dir, "edit_hits_%d.npy" % rank), hits2)
    np.save(os.path.join(out_dir, "fresh_%d.npy" % rank), fresh)

    if rank == 0:
        np.save(os.path.join(out_dir, "image.npy"), t.numpy())
        np.save(os.path.join(out_dir, "job.npy"), job_out.numpy())
        np.save(os.path.join(out_dir, "samples.npy"), tp.numpy())
        np.save(os.path.join(out_dir, "hits.npy"), np.frombuffer(b"".join(gathered), dtype=pyoracle.HIT_DTYPE))
        np.save(os.path.join(out_dir, "nodes.npy"), nodes)
        np.save(os.path.join(out_dir, "meta.npy"), np.array([root, froot, dirty]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_result(tmp_path, port, api):
    import torch.multiprocessing as mp
    from oracle import pyoracle
    world = 2
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)

    nodes = np.load(tmp_path / "nodes.npy")
    root, froot, dirty = (int(v) for v in np.load(tmp_path / "meta.npy"))
    sc = api.Scene("sphere_noise", 6, seed=4)
    assert np.array_equal(nodes, sc.nodes) and root == sc.root          # replica == original
    PI_F = float(np.float32(3.14159265358979))
    cam = api.default_camera(sc.lower, sc.upper)
    ocam = port.camera(list(cam.position), -(PI_F / 4.0), 0.0)
    w, h = 96, 80
    sd = port.find_subdags(nodes, root)
    colours = np.tile(np.array([[0.7, 0.6, 0.5]], dtype=np.float32), (256, 1))
    whole, _, _ = port.render(nodes, sd, colours, ocam, pyoracle.PtParams(w, h, 2, 2, 1, 1, 1, 1, 0.0035, 0, 0, 0, w, h, 0))
    assert np.array_equal(np.load(tmp_path / "image.npy"), whole)        # tile sharding + reduce is exact
    rays = port.camera_rays(ocam, w, h)
    want, _, _ = port.trace(nodes, sd, rays, True, -1.0)
    assert np.load(tmp_path / "hits.npy").tobytes() == want.tobytes()
    # the chunked job exchange assembled exactly the single-process result on rank 0
    from test_host_shim import compact_of
    job = np.load(tmp_path / "job.npy").view(api.COMPACT_DTYPE).reshape(-1)
    assert job.tobytes() == compact_of(want[:5 * 16 * w], api).tobytes()
    assert api.expand_hits(rays[:len(job)], job).tobytes() == want[:len(job)].tobytes()
    # sample sharding: the same 4 samples per pixel whoever renders them (float sums regrouped: equal to rounding)
    four, _, _ = port.render(nodes, sd, colours, ocam, pyoracle.PtParams(w, h, 4, 2, 1, 1, 1, 1, 0.0035, 10, 0, 0, w, h, 0))
    np.testing.assert_allclose(np.load(tmp_path / "samples.npy"), four, rtol=2e-6, atol=1e-6)
    # the delta broadcast left both ranks with the same edited array
    f0, f1 = np.load(tmp_path / "fresh_0.npy"), np.load(tmp_path / "fresh_1.npy")
    assert np.array_equal(f0, f1) and len(f0) == len(nodes) + 1 and dirty == len(nodes) and froot == len(nodes)
    sd2 = port.find_subdags(f0, froot)
    want2, _, _ = port.trace(f0, sd2, rays, True, -1.0)
    got2 = np.concatenate([np.load(tmp_path / "edit_hits_0.npy"), np.load(tmp_path / "edit_hits_1.npy")])
    assert got2.tobytes() == want2.tobytes() and want2.tobytes() != want.tobytes()


def test_partitions_cover_everything_once():
    from cubiquity_b200 import sharding
    for n, world in [(0, 3), (7, 8), (2073600, 8), (100, 1)]:
        spans = [sharding.split_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for h, world in [(1080, 8), (2160, 4), (70, 3)]:
        rows = sorted(sum([sharding.tile_rows(h, world, r) for r in range(world)], []))
        assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))

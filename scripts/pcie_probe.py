"""What the host link allows for one 1080p frame of cbq_trace: 49.8 MB of rays in, 82.9 MB of hits out, pinned memory.
Prints the time of each copy alone, of both at once on two streams, and of the same bytes in 2^18-ray pieces."""
import json
import time

import torch


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    n = 1920 * 1080
    h_in = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n * 24, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    def pieces():
        step = 1 << 18
        for b in range(0, n, step):
            e = min(n, b + step)
            with torch.cuda.stream(s1):
                d_in[b * 24:e * 24].copy_(h_in[b * 24:e * 24], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[b * 40:e * 40].copy_(d_out[b * 40:e * 40], non_blocking=True)

    r = {"h2d_49.8MB_ms": timed(h2d), "d2h_82.9MB_ms": timed(d2h), "both_ms": timed(both), "both_in_2^18_ray_pieces_ms": timed(pieces)}
    r["h2d_GBps"] = n * 24 / r["h2d_49.8MB_ms"] / 1e6
    r["d2h_GBps"] = n * 40 / r["d2h_82.9MB_ms"] / 1e6
    print(json.dumps(r))


if __name__ == "__main__":
    main()

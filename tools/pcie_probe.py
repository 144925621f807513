#!/usr/bin/env python
"""PCIe rates of the box, to put the end-to-end (host buffers) numbers of bench.py in context: pinned host <-> device copies of
2 GB on one stream, split over two streams, and both directions at once."""
import json
import time

import torch

n = 2 * 1024 ** 3
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    d.copy_(h, non_blocking=True)


def d2h():
    h.copy_(d, non_blocking=True)


def h2d_split():
    with torch.cuda.stream(s1):
        d[:n // 2].copy_(h[:n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        d[n // 2:].copy_(h[n // 2:], non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


out = {"bytes": n, "h2d_GBs": n / timed(h2d) / 1e9, "d2h_GBs": n / timed(d2h) / 1e9, "h2d_two_streams_GBs": n / timed(h2d_split) / 1e9,
       "duplex_each_GBs": n / timed(both) / 1e9}
print(json.dumps(out))

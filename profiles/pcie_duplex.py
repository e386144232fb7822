"""Pinned-memory PCIe rates on this box: H2D alone, D2H alone, both at once (48 MB each, as the e2e path moves per step)."""
import json

import torch

n = 48_000_000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize()
    s2.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


import time


def wall(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


res = {k: wall(f) for k, f in (("h2d_ms", h2d), ("d2h_ms", d2h), ("both_ms", both))}
res.update({"h2d_GBs": n / res["h2d_ms"] / 1e6, "d2h_GBs": n / res["d2h_ms"] / 1e6, "duplex_aggregate_GBs": 2 * n / res["both_ms"] / 1e6})
print(json.dumps(res))

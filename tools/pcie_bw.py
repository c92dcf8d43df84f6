#!/usr/bin/env python3
"""Pinned-memory copy bandwidth of this box (the ceiling of bench.py's e2e leg): H2D alone, D2H alone, both at once."""
import json, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
out = {"bytes": n, "h2d_GBps": n / timed(h2d) / 1e9, "d2h_GBps": n / timed(d2h) / 1e9}
t = timed(both); out["both_h2d_GBps"] = n / t / 1e9; out["both_d2h_GBps"] = n / t / 1e9
print(json.dumps(out))

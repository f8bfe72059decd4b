#!/usr/bin/env python3
"""Is the step-time noise host-side or GPU-side?  (a) a pure GPU chain timed with events, no host involvement
inside; (b) the same chain with a host round trip (tiny D2H + sync) between the kernels."""
import time, torch
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
def chain(n, roundtrip):
    y = x
    for _ in range(n):
        y = y @ x
        y = y / y.abs().max()
        if roundtrip:
            float(y[0, 0])           # D2H + sync: the host must answer before the next kernel is enqueued
    return y
for rt in (False, True):
    chain(4, rt); torch.cuda.synchronize()
    ts = []
    for _ in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); chain(40, rt); e1.record(); torch.cuda.synchronize()
        ts.append((round(e0.elapsed_time(e1), 1), round((time.perf_counter() - t0) * 1e3, 1)))
    ev = sorted(t[0] for t in ts)
    print("roundtrip" if rt else "pure-gpu ", "median", ev[len(ev) // 2], "max", ev[-1], "outliers(>1.15x)", sum(1 for v in ev if v > 1.15 * ev[len(ev) // 2]), "of", len(ev), flush=True)

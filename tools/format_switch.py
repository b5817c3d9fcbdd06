#!/usr/bin/env python3
"""One engine, two formats in a row (4K 8-bit random access, then 4K Main10): value-mode frames/s of the second format, to check that
nothing of the first format's state (surface pool, staging, stream bookkeeping) slows the second down."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from libde265_b200 import synth
from libde265_b200.engine import Engine


def run(eng, seq, key_slot, bd, steps=4):
    ref0 = synth.random_planes(3840, 2160, bd, 7)
    eng.upload_slot(key_slot, seq[0].params, ref0)
    prepared = [eng.prepare(p) for p in seq]
    stream = torch.cuda.ExternalStream(eng.stream())
    n = [0]

    def step():
        v = n[0] % bench.STEP_VARIANTS
        n[0] += 1
        for h in prepared[32 * v:32 * (v + 1)]:
            eng.run_prepared(h)

    for _ in range(4):
        step()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    eng.join()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    eng.sync()
    for h in prepared:
        eng.free_prepared(h)
    return 32 * steps / (ms / 1e3)


s8, k8, _ = bench.build_workload(3840, 2160, 8, seed0=1000)
s10, k10, _ = bench.build_workload(3840, 2160, 10, seed0=1000)
eng = Engine(0)
print("fresh engine, Main10:", round(run(eng, s10, k10, 10), 1), flush=True)
eng.close()
eng = Engine(0)
print("8-bit first:", round(run(eng, s8, k8, 8), 1), flush=True)
print("then Main10 on the same engine:", round(run(eng, s10, k10, 10), 1), flush=True)
print("then 8-bit again:", round(run(eng, s8, k8, 8), 1), flush=True)
eng.close()

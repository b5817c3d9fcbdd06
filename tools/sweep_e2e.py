#!/usr/bin/env python3
"""Measures the e2e leg (b200_engine_submit_picture from host records + D2H of every picture into pinned memory) of the headline
workload for a list of engine settings (environment assignments), with the engine's host-side profile (B200_HOST_PROF)."""
import os, sys, time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from libde265_b200 import capi, synth
from libde265_b200.engine import Engine

os.environ["B200_HOST_PROF"] = "1"
steps = 4
W, H = 3840, 2160
seq, key_slot, gen_s = bench.build_workload(W, H, 8, seed0=1000)
ref0 = synth.random_planes(W, H, 8, 7)
outs = [[torch.empty((H, W), dtype=torch.uint8).pin_memory(), torch.empty((H // 2, W // 2), dtype=torch.uint8).pin_memory(),
         torch.empty((H // 2, W // 2), dtype=torch.uint8).pin_memory()] for _ in range(8)]
if not os.environ.get("B200_E2E_PAGEABLE"):
    print("records pinned:", bench.pin_records(seq, torch), flush=True)
for setting in (sys.argv[1:] or [""]):
    env = dict(kv.split("=", 1) for kv in setting.split() if "=" in kv)
    for k, v in env.items():
        os.environ[k] = v
    eng = Engine(0)
    eng.upload_slot(key_slot, seq[0].params, ref0)
    n = [0]

    def step():
        v = n[0] % bench.STEP_VARIANTS
        n[0] += 1
        for i, p in enumerate(seq[32 * v:32 * (v + 1)]):
            (eng.submit if os.environ.get('B200_E2E_SYNC') else eng.submit_async)(p)
            o = outs[i & 7]
            capi.check(eng.lib.b200_engine_read_slot_async(eng.handle, p.params.dst_slot, capi.PlaneArray(*[t.data_ptr() for t in o]),
                                                           capi.StrideArray(*[t.stride(0) for t in o])), "read")

    for _ in range(5):
        step()
    eng.sync()
    t0 = time.time()
    for _ in range(steps):
        step()
    eng.sync()
    dt = time.time() - t0
    print(f"[{setting or 'default'}] e2e {32 * steps / dt:8.1f} frames/s  ({1e3 * dt / (32 * steps):.3f} ms per picture)", flush=True)
    eng.close()
    for k in env:
        os.environ.pop(k, None)

#!/usr/bin/env python3
"""Builds the headline workload once and measures `value` (records resident, pictures pipelined) for a list of engine
settings given as environment assignments, e.g.:  sweep_bench.py "B200_IND_STREAMS=1" "B200_IND_STREAMS=2 B200_INTRA_I_GRID=64"
Optional first argument --timeline=FILE records a B200_TIMELINE for the LAST setting."""
import os, sys, time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA initialises: one hardware queue per engine stream (see capi.py)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from libde265_b200 import synth
from libde265_b200.engine import Engine

args = sys.argv[1:]
tl = None
if args and args[0].startswith("--timeline="):
    tl = args.pop(0).split("=", 1)[1]
steps = 6
seq, key_slot, gen_s = bench.build_workload(3840, 2160, 8, seed0=1000)
ref0 = synth.random_planes(3840, 2160, 8, 7)
print(f"workload built in {gen_s:.1f} s", flush=True)
for i, setting in enumerate(args or [""]):
    env = dict(kv.split("=", 1) for kv in setting.split() if "=" in kv)
    for k, v in env.items():
        os.environ[k] = v
    if tl and i == len(args) - 1:
        os.environ["B200_TIMELINE"] = tl
    eng = Engine(0)
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
    print(f"[{setting or 'default'}] {32 * steps / (ms / 1e3):8.1f} frames/s  {ms / steps:7.3f} ms/step", flush=True)
    eng.sync()
    for h in prepared:
        eng.free_prepared(h)
    eng.close()
    for k in env:
        os.environ.pop(k, None)

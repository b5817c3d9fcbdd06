#!/usr/bin/env python3
"""Runs one synthetic 4K picture (B by default, or I) through the whole pipeline on ONE stream a few times — for ncu captures
of every kernel of a picture (`ncu --set full -k regex:... python tools/one_picture.py B 3`)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libde265_b200 import synth
from libde265_b200.engine import Engine
W, H = 3840, 2160
kind = sys.argv[1] if len(sys.argv) > 1 else "B"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine(0)
eng.set_streams(1)
pic = synth.make_picture(W, H, kind, seed=1002 if kind == "B" else 1000, dst_slot=2, ref_slots=(0, 1) if kind != "I" else ())
for s in (0, 1):
    eng.upload_slot(s, pic.params, synth.random_planes(W, H, 8, s + 1))
h = eng.prepare(pic)
eng.enable_timing(True)
for _ in range(reps):
    eng.run_prepared(h)
eng.sync()
ms, n = eng.timing_sum(reset=True)
print({k: round(v / n, 4) for k, v in ms.items()}, "ms per picture over", n)
eng.free_prepared(h)
eng.close()

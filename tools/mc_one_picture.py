#!/usr/bin/env python3
"""Runs the inter-prediction stage of one synthetic 4K B picture a few times (for ncu captures of the MC kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libde265_b200 import capi, synth
from libde265_b200.engine import Engine
W, H = 3840, 2160
bd = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine(0)
eng.set_streams(1)
b = synth.make_picture(W, H, "B", seed=1002, dst_slot=2, ref_slots=(0, 1), bit_depth=bd)
for s in (0, 1):
    eng.upload_slot(s, b.params, synth.random_planes(W, H, bd, s + 1))
b.c.params.stop_after_stage = capi.STAGE_INTER_PRED
h = eng.prepare(b)
for _ in range(reps):
    eng.run_prepared(h)
eng.sync()
eng.enable_timing(True)
for _ in range(10):
    eng.run_prepared(h)
ms, n = eng.timing_sum(reset=True)
print("inter_pred ms per picture:", ms["inter_pred"] / n, "tiles/units:", len(b.pus))
eng.free_prepared(h)
eng.close()

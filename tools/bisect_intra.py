#!/usr/bin/env python3
"""Debug: find which intra TU first differs from the oracle (stage recon) for several synthetic configurations."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libde265_b200 import capi, synth
from libde265_b200.engine import Engine
import oracle_lib
eng, orc = Engine(0), oracle_lib.Oracle()
W, H = 256, 128
cfgs = {"nores": dict(cbf_prob=0.0, special_frac=0.0), "res": dict(cbf_prob=1.0, special_frac=0.0),
        "big": dict(cbf_prob=0.5, special_frac=0.0, size_area=(0.5, 0.5, 0.0, 0.0)),
        "small": dict(cbf_prob=0.5, special_frac=0.0, size_area=(0.0, 0.0, 0.0, 1.0)),
        "special": dict(cbf_prob=0.9, special_frac=0.2)}
for name, kw in cfgs.items():
    p = synth.make_picture(W, H, "I", seed=3, dst_slot=0, deblock=False, sao=False, **kw)
    p.c.params.stop_after_stage = capi.STAGE_RECON
    eng.submit(p); g = eng.read_slot(0, p.params)
    orc.reconstruct(p); o = orc.read_slot(0, p.params)
    tot = sum(int((a != b).sum()) for a, b in zip(g, o))
    print(f"== {name}: {tot} samples differ")
    if not tot: continue
    # attribute: per TU (in decode order) does its block match?
    shown = 0
    for i, tu in enumerate(p.tus):
        c, nT = int(tu["cidx"]), 1 << int(tu["log2_size"])
        x, y = int(tu["x"]), int(tu["y"])
        a, b = g[c][y:y + nT, x:x + nT], o[c][y:y + nT, x:x + nT]
        if (a != b).any():
            print(f"  first bad TU #{i}: plane {c} ({x},{y}) nT={nT} flags={int(tu['flags']):#x} mode={int(tu['intra_mode'])} ncoef={int(tu['n_coeff'])} avail={int(tu['avail']):#x}")
            print("   gpu\n", a[:4, :8], "\n   oracle\n", b[:4, :8])
            shown += 1
            if shown >= 3: break

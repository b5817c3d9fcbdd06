#!/usr/bin/env python3
"""Which TUs are wrong?  Runs a synthetic sequence on the engine and the oracle up to the recon stage and
classifies mismatching samples by the TU that covers them.  Usage: python tools/diag_tu.py W H bitdepth [kw=val ...]"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libde265_b200 import capi, synth  # noqa: E402
from libde265_b200.engine import Engine  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    W, H, bd = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    kw = {}
    for a in sys.argv[4:]:
        k, v = a.split("=")
        kw[k] = eval(v)
    eng, orc = Engine(0), oracle_lib.Oracle()
    ref = synth.random_planes(W, H, bd, 99)
    pics = [synth.make_picture(W, H, "I", seed=11, dst_slot=0, bit_depth=bd, **kw)]
    eng.upload_slot(5, pics[0].params, ref)
    orc.upload_slot(5, pics[0].params, ref)
    pics.append(synth.make_picture(W, H, "P", seed=12, dst_slot=1, ref_slots=(0, 5), bit_depth=bd, **kw))
    pics.append(synth.make_picture(W, H, "B", seed=13, dst_slot=2, ref_slots=(0, 1, 5), bit_depth=bd, **kw))
    for i, p in enumerate(pics):
        for st in (capi.STAGE_INTER_PRED, capi.STAGE_RECON, capi.STAGE_ALL):
            p.c.params.stop_after_stage = st
            eng.submit(p)
            orc.reconstruct(p)
            g, o = eng.read_slot(p.params.dst_slot, p.params), orc.read_slot(p.params.dst_slot, p.params)
            bad = [(a != b) for a, b in zip(g, o)]
            n = sum(int(b.sum()) for b in bad)
            print(f"pic {i} stage {st}: {n} samples differ")
            if n and st == capi.STAGE_RECON:
                cnt = collections.Counter()
                tot = collections.Counter()
                for t in p.tus:
                    c, x, y, s = int(t["cidx"]), int(t["x"]), int(t["y"]), 1 << int(t["log2_size"])
                    key = (int(t["log2_size"]), c > 0, hex(int(t["flags"])))
                    tot[key] += 1
                    if bad[c][y:y + s, x:x + s].any():
                        cnt[key] += 1
                for k in sorted(tot):
                    if cnt[k]:
                        print(f"   log2={k[0]} chroma={k[1]} flags={k[2]}: {cnt[k]} of {tot[k]} TUs wrong")
                return
        p.c.params.stop_after_stage = 0


main()

#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle diagnosis (the stage dump of SURVEY §5): for every picture runs the engine
and the CPU oracle with stop_after_stage = inter-pred / recon / deblock / all and reports the first
stage, plane and sample that differ.  Usage: python tools/diag_gpu.py [girlshy|synth] [max_pictures]"""
import ctypes as C
import gzip
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libde265_b200 import capi, synth  # noqa: E402
from libde265_b200.engine import Engine  # noqa: E402
import oracle_lib  # noqa: E402

STAGES = [(capi.STAGE_INTER_PRED, "inter_pred"), (capi.STAGE_RECON, "recon"), (capi.STAGE_DEBLOCK, "deblock"), (capi.STAGE_ALL, "all")]


def load_girlshy():
    lib = capi.load()
    raw = gzip.open(os.path.join(ROOT, "tests/golden/girlshy_records.bin.gz"), "rb").read()
    pics, pos = [], 0
    keep = []
    while pos < len(raw):
        (n,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        buf = C.create_string_buffer(raw[pos:pos + n], n)
        pos += n
        p = capi.Picture()
        assert lib.b200_picture_deserialize(buf, n, C.byref(p)) > 0
        keep.append(buf)
        pics.append(p)
    return pics, keep


def compare(tag, a, b):
    bad = False
    for c, (x, y) in enumerate(zip(a, b)):
        d = np.argwhere(x != y)
        if len(d):
            bad = True
            yy, xx = d[0]
            print(f"  MISMATCH {tag} plane {c}: {len(d)} samples differ, first at x={xx} y={yy}: gpu={x[yy, xx]} oracle={y[yy, xx]}; "
                  f"bbox x[{d[:,1].min()},{d[:,1].max()}] y[{d[:,0].min()},{d[:,0].max()}]")
            ys, xs = max(0, yy - 1), max(0, xx - 2)
            print("   gpu   :", x[ys:ys + 3, xs:xs + 12].tolist())
            print("   oracle:", y[ys:ys + 3, xs:xs + 12].tolist())
    return bad


def run(pics, eng, orc, max_pics):
    first_bad = None
    for i, pic in enumerate(pics[:max_pics]):
        cp = getattr(pic, "c", pic)
        for st, name in STAGES:
            cp.params.stop_after_stage = st
            eng.submit(cp)
            g = eng.read_slot(cp.params.dst_slot, cp.params)
            orc.reconstruct(cp)
            o = orc.read_slot(cp.params.dst_slot, cp.params)
            if compare(f"pic {i} stage {name}", g, o) and first_bad is None:
                first_bad = (i, name)
        cp.params.stop_after_stage = 0
        print(f"pic {i}: n_pu={cp.n_pu} n_tu={cp.n_tu} {'OK' if first_bad is None else 'first bad: ' + str(first_bad)}", flush=True)
        if first_bad is not None and i > first_bad[0] + 1:
            break
    return first_bad


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "girlshy"
    max_pics = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    eng, orc = Engine(0), oracle_lib.Oracle()
    if what == "girlshy":
        pics, keep = load_girlshy()
        bad = run(pics, eng, orc, max_pics)
    else:
        W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (416, 240)
        bd = int(sys.argv[5]) if len(sys.argv) > 5 else 8
        ref = synth.random_planes(W, H, bd, 99)
        p0 = synth.make_picture(W, H, "I", seed=1, dst_slot=0, bit_depth=bd)
        eng.upload_slot(5, p0.params, ref)
        orc.upload_slot(5, p0.params, ref)
        pics = [p0,
                synth.make_picture(W, H, "P", seed=2, dst_slot=1, ref_slots=(0, 5), bit_depth=bd),
                synth.make_picture(W, H, "B", seed=3, dst_slot=2, ref_slots=(0, 1, 5), bit_depth=bd),
                synth.make_picture(W, H, "B", seed=4, dst_slot=3, ref_slots=(0, 1, 2), weighted=True, bit_depth=bd),
                synth.make_picture(W, H, "B", seed=5, dst_slot=4, ref_slots=(1, 2, 3), n_slices=3, scaling_list=True, bit_depth=bd)]
        bad = run(pics, eng, orc, max_pics)
    print("RESULT:", "all pictures bit-exact" if bad is None else f"first mismatch at {bad}")
    return 0 if bad is None else 1


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""CPU-only analysis of the intra dependency DAG the engine's k_intra executes: tasks as the host planner forms them (the
small TUs of one plane inside a 16x16-luma / 8x8-chroma region, or one larger TU), level = 1 + max level of the neighbour
units a task may read (availability masks).  Prints the DAG depth (the number of dependent task latencies an intra picture
costs at least) and the tasks per level (how many warps can work at the same time).
Usage: python tools/dag_depth.py [synthetic | intra4k | intra1080]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libde265_b200 import synth  # noqa: E402


def analyse(tus, W, H, name):
    intra = (tus["flags"] & 1) != 0
    idx = np.nonzero(intra)[0]
    cur_key, cur_task, ntask = [None] * 3, [0] * 3, 0
    task_of = np.zeros(len(idx), np.int64)
    for k, i in enumerate(idx):
        t = tus[i]
        c, nT = int(t["cidx"]), 1 << int(t["log2_size"])
        G = 16 >> (1 if c else 0)
        key = ("L", i) if nT >= G else (int(t["y"]) // G, int(t["x"]) // G)
        if key != cur_key[c]:
            cur_key[c], cur_task[c] = key, ntask
            ntask += 1
        task_of[k] = cur_task[c]
    lvl = [np.zeros(((H >> (1 if c else 0)) // 4 + 2, (W >> (1 if c else 0)) // 4 + 2), np.int32) for c in range(3)]
    tl = np.zeros(ntask, np.int32)
    order = np.argsort(task_of, kind="stable")
    bounds = np.searchsorted(task_of[order], np.arange(ntask + 1))
    for t in range(ntask):
        ks = order[bounds[t]:bounds[t + 1]]
        best = 0
        for k in ks:
            tu = tus[idx[k]]
            c, x4, y4, n4, av = int(tu["cidx"]), int(tu["x"]) // 4, int(tu["y"]) // 4, (1 << int(tu["log2_size"])) // 4, int(tu["avail"])
            L = lvl[c]
            for g in range(2 * n4):
                if (av >> g) & 1 and x4 > 0:
                    best = max(best, L[y4 + g, x4 - 1])
                if (av >> (17 + g)) & 1 and y4 > 0:
                    best = max(best, L[y4 - 1, x4 + g])
            if (av >> 16) & 1 and x4 > 0 and y4 > 0:
                best = max(best, L[y4 - 1, x4 - 1])
        tl[t] = best + 1
        for k in ks:
            tu = tus[idx[k]]
            c, x4, y4, n4 = int(tu["cidx"]), int(tu["x"]) // 4, int(tu["y"]) // 4, (1 << int(tu["log2_size"])) // 4
            lvl[c][y4:y4 + n4, x4:x4 + n4] = tl[t]
    # weighted critical path with the per-task latencies measured on the B200 (B200_TRACE_INTRA): dependent part of a region task
    # = 3.0 us + 0.43 us per TU, a large luma / chroma TU 5.5 / 3.2 us
    fin = [np.zeros_like(lvl[c], dtype=np.float64) for c in range(3)]
    end = 0.0
    for t in range(ntask):
        ks = order[bounds[t]:bounds[t + 1]]
        start = 0.0
        for k in ks:
            tu = tus[idx[k]]
            c, x4, y4, n4, av = int(tu["cidx"]), int(tu["x"]) // 4, int(tu["y"]) // 4, (1 << int(tu["log2_size"])) // 4, int(tu["avail"])
            F = fin[c]
            for g in range(2 * n4):
                if (av >> g) & 1 and x4 > 0:
                    start = max(start, F[y4 + g, x4 - 1])
                if (av >> (17 + g)) & 1 and y4 > 0:
                    start = max(start, F[y4 - 1, x4 + g])
            if (av >> 16) & 1 and x4 > 0 and y4 > 0:
                start = max(start, F[y4 - 1, x4 - 1])
        tu0 = tus[idx[ks[0]]]
        large = len(ks) == 1 and (1 << int(tu0["log2_size"])) > 8
        cost = (5.5 if int(tu0["cidx"]) == 0 else 3.2) if large else 3.0 + 0.43 * len(ks)
        e = start + cost
        for k in ks:
            tu = tus[idx[k]]
            c, x4, y4, n4 = int(tu["cidx"]), int(tu["x"]) // 4, int(tu["y"]) // 4, (1 << int(tu["log2_size"])) // 4
            fin[c][y4:y4 + n4, x4:x4 + n4] = e
        end = max(end, e)
    widths = np.bincount(tl)[1:]
    print(f"{name}: modelled critical path {end / 1000:.2f} ms (measured task latencies)")
    print(f"{name}: {len(idx)} intra TUs in {ntask} tasks; DAG depth {int(tl.max())} levels; tasks per level mean {widths.mean():.1f}, "
          f"p10 {int(np.percentile(widths, 10))}, p50 {int(np.percentile(widths, 50))}, p90 {int(np.percentile(widths, 90))}, max {int(widths.max())}")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
    if what == "synthetic":
        p = synth.make_picture(3840, 2160, "I", seed=1000)
        analyse(p.tus, 3840, 2160, "synthetic 4K I picture of bench.py (CTB 64)")
        return
    import oracle_lib
    from libde265_b200 import de265
    dec = de265.Decoder(oracle_lib.ref_path("libde265_hooked.so"))
    store = []

    def sink(pic, planes, strides):
        store.append((np.ctypeslib.as_array(C.cast(pic.tus, C.POINTER(C.c_uint8)), shape=(pic.n_tu * 24,)).view(synth.TU_DT).copy(), pic.params.width,
                      pic.params.height))
        return 0

    dec.attach(sink)
    dec.decode_stream(open(os.path.join(ROOT, "tests", "golden", what + ".h265"), "rb").read(), lambda img: None)
    dec.close()
    for n, (tus, W, H) in enumerate(store):
        analyse(tus, W, H, f"{what} picture {n}")


if __name__ == "__main__":
    main()

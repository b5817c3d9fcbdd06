#!/usr/bin/env python3
"""Reads a B200_TIMELINE file (engine.cu: one CUDA-event interval per launch: `stream kernel poc t0_ms t1_ms`) and prints
   * per kernel: launches, mean duration, share of the covered time,
   * concurrency: how much of the covered wall time has 1, 2, 3, ... launches in flight, how many PICTURES overlap,
   * an ASCII timeline of a window (one row per stream).
Usage: timeline.py FILE [t_start_ms t_end_ms]"""
import sys
from collections import defaultdict


def main():
    rows = []
    for line in open(sys.argv[1]):
        p = line.split()
        if len(p) == 5:
            rows.append((int(p[0]), p[1], int(p[2]), float(p[3]), float(p[4])))
    if not rows:
        print("empty timeline")
        return
    t_lo = float(sys.argv[2]) if len(sys.argv) > 3 else min(r[3] for r in rows)
    t_hi = float(sys.argv[3]) if len(sys.argv) > 3 else max(r[4] for r in rows)
    rows = [r for r in rows if r[4] > t_lo and r[3] < t_hi]
    span = t_hi - t_lo
    per = defaultdict(lambda: [0, 0.0])
    for s, k, poc, a, b in rows:
        per[k][0] += 1
        per[k][1] += b - a
    print(f"window {t_lo:.3f} .. {t_hi:.3f} ms ({span:.3f} ms), {len(rows)} launches on {len({r[0] for r in rows})} streams, {len({(r[0], r[2]) for r in rows})} pictures")
    print(f"{'kernel':12s} {'launches':>8s} {'mean us':>9s} {'sum ms':>8s} {'sum/window':>10s}")
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:12s} {n:8d} {1e3 * t / n:9.1f} {t:8.3f} {t / span:10.2f}")
    # concurrency profile (launch intervals and picture intervals)
    inst, last = {}, {}
    for i, r in enumerate(sorted(rows, key=lambda r: r[3])):  # picture instance = consecutive launches of one poc on one stream
        if last.get(r[0], (None, None))[0] != r[2]:
            last[r[0]] = (r[2], i)
        inst[(r[0], r[1], r[2], r[3])] = (r[0], last[r[0]][1])
    for what, key in (("launches", lambda r: (r[0], r[1], r[2], r[3])), ("pictures", lambda r: inst[(r[0], r[1], r[2], r[3])])):
        iv = {}
        for r in rows:
            k = key(r)
            a, b = iv.get(k, (1e30, -1e30))
            iv[k] = (min(a, r[3]), max(b, r[4]))
        ev = sorted([(a, 1) for a, b in iv.values()] + [(b, -1) for a, b in iv.values()])
        hist, cur, last = defaultdict(float), 0, ev[0][0]
        for t, d in ev:
            hist[cur] += t - last
            cur, last = cur + d, t
        tot = sum(hist.values())
        print(f"{what} in flight: " + "  ".join(f"{n}: {100 * hist[n] / tot:.0f}%" for n in sorted(hist) if hist[n] / tot >= 0.005))
    # ASCII timeline
    cols = 160
    glyph = {"mc": "M", "residual": "r", "mark": ".", "intra": "I", "deblock_v": "d", "deblock_h": "d", "sao_prep": ".", "sao": "s", "extend": "x"}
    for s in sorted({r[0] for r in rows}):
        line = [" "] * cols
        for st, k, poc, a, b in rows:
            if st != s:
                continue
            i0, i1 = int((max(a, t_lo) - t_lo) / span * cols), int((min(b, t_hi) - t_lo) / span * cols)
            for i in range(i0, min(cols, max(i1, i0 + 1))):
                line[i] = glyph.get(k, "?")
        print(f"s{s} |" + "".join(line) + "|")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total GPU time of the last `--last-frac` of the
launches (the timed step after warm-up) and the top kernels by total time, optionally split by grid size."""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--last-frac", type=float, default=0.5)
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--by-grid", action="store_true")
    a = ap.parse_args()
    with open(a.path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    rows = []
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        if row[ui] == "ns":
            v /= 1000.0
        rows.append((re.sub(r"\(.*", "", row[ki]), row[gi], v))
    sel = rows[int(len(rows) * (1 - a.last_frac)):]
    tot = sum(v for _, _, v in sel)
    print("%d launches in the file; last %.0f%%: %d launches, %.3f ms GPU time" % (len(rows), 100 * a.last_frac, len(sel), tot / 1e3))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, v in sel:
        key = (k, g) if a.by_grid else k
        agg[key][0] += 1
        agg[key][1] += v
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:a.top]:
        name = ("%s grid %s" % k) if a.by_grid else k
        print("%9.1f us %5.1f%%  %4d x %8.1f  %s" % (v, 100 * v / tot, n, v / n, name[-110:]))


if __name__ == "__main__":
    main()

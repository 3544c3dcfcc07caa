#!/usr/bin/env python3
"""Summarise a tools/trace/libqpzktrace.so CSV: per-kernel busy time and launch count, GPU busy fraction (union of
kernel intervals), average number of kernels running at once, and per-stream gaps.  usage: summarize.py trace.csv [skip_frac]
skip_frac (default 0.3): ignore the first part of the trace (warm-up / graph capture)."""
import csv, collections, re, sys

def main():
    path = sys.argv[1]
    skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    recs = []
    for r in csv.reader(open(path)):
        if not r or r[0] not in "KCS" or len(r) < 10:
            continue
        recs.append((r[0], re.sub(r"\(.*", "", r[1]).replace("void ", "").replace("qpzk::", ""), int(r[2]), int(r[3]), int(r[4]), r[5], r[6], r[7], r[8]))
    t0 = min(r[3] for r in recs); t1 = max(r[4] for r in recs)
    cut = t0 + (t1 - t0) * skip
    recs = [r for r in recs if r[3] >= cut]
    t0 = min(r[3] for r in recs); span = t1 - t0
    agg = collections.OrderedDict()
    for k, name, st, s, e, grid, block, regs, smem in recs:
        a = agg.setdefault(name, [0, 0, grid, block, regs, smem]); a[0] += 1; a[1] += e - s
    ev = []
    for r in recs:
        if r[0] == "K":
            ev.append((r[3], 1)); ev.append((r[4], -1))
    ev.sort()
    busy = 0; conc_int = 0; depth = 0; last = ev[0][0]
    for t, d in ev:
        if depth > 0: busy += t - last; conc_int += depth * (t - last)
        depth += d; last = t
    streams = collections.defaultdict(list)
    for r in recs: streams[r[2]].append((r[3], r[4]))
    print("window %.3f ms, %d records, %d streams; GPU busy (>=1 kernel) %.1f%%; mean kernels in flight while busy %.2f" %
          (span / 1e6, len(recs), len(streams), 100.0 * busy / span, conc_int / max(busy, 1)))
    tot = sum(a[1] for a in agg.values())
    print("%-44s %7s %10s %7s %10s  grid/block/regs/smem" % ("kernel", "count", "sum ms", "share", "avg us"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-44s %7d %10.3f %6.1f%% %10.1f  %s/%s/%s/%s" % (name[:44], a[0], a[1] / 1e6, 100.0 * a[1] / tot, a[1] / a[0] / 1e3, a[2], a[3], a[4], a[5]))
    print("sum of kernel durations %.3f ms = %.2f x window" % (tot / 1e6, tot / span))

main()

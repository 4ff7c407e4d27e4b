#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): share per kernel and the
launch sequence of one pass.  usage: python tools/launch_summary.py launches.csv [note]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    gs, bs = hdr.index('Grid Size'), hdr.index('Block Size')
    agg, seq = collections.OrderedDict(), []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        v *= {'us': 1e-3, 'ns': 1e-6, 'ms': 1.0, 's': 1e3}.get(r[mu], 1.0)
        a = agg.setdefault(r[kn], [0, 0.0, r[gs], r[bs]])
        a[0] += 1
        a[1] += v
        seq.append((r[kn], v))
    tot = sum(v[1] for v in agg.values())
    if len(sys.argv) > 2:
        print("#", sys.argv[2])
    print("# per-launch times under ncu are cold-cache and serialised: read the shares, not the absolutes")
    print(f"# launches {len(seq)}  total {tot:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1] / tot * 100:5.1f}%  {v[1]:9.3f} ms  n={v[0]:4d}  grid {v[2]:>14s} block {v[3]:>12s}  {k}")
    idx = [i for i, (n, v) in enumerate(seq) if 'k_trace' in n and n.rstrip().endswith('1>(spb::WaveArgs, unsigned int)')]
    if idx:
        s = idx[len(idx) // 2]
        print("\n# one pass in launch order, ms:")
        for n, v in seq[s:s + 16]:
            print(f"  {v:8.3f}  {n}")


if __name__ == "__main__":
    main()

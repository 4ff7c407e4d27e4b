#!/usr/bin/env python
"""Per-CUDA-source-line instruction / lane / stall-sample shares of the kernels in an .ncu-rep
(captured with --import-source on from a -lineinfo build).
usage: python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top-N]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = func = hdr = None
    launch = -1
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name" or r[0] == "Function Name":
            if r[0] == "Function Name" and r[1] != func:
                launch += 1
            func = r[1]
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].isdigit() and len(r) > 10 and r[2] == "-":
            try:
                inst, thr, samp = int(r[7]), int(r[8]), int(r[6])
            except ValueError:
                continue
            key = (func, cur_file, int(r[0]))
            a = agg.setdefault(key, [r[1], 0, 0, 0])
            a[1] += inst
            a[2] += thr
            a[3] += samp
    funcs = collections.OrderedDict()
    for (f, fi, ln), a in agg.items():
        funcs.setdefault(f, []).append((fi, ln, a))
    for f, lines in funcs.items():
        if want not in f:
            continue
        T = sum(a[1] for _, _, a in lines) or 1
        S = sum(a[3] for _, _, a in lines) or 1
        TT = sum(a[2] for _, _, a in lines)
        print("=" * 100)
        print(f, "warp-inst", T, "threads/inst %.2f" % (TT / T), "samples", S)
        for fi, ln, a in sorted(lines, key=lambda x: -x[2][1])[:top]:
            print(f"  {fi[:18]:18s}:{ln:5d} inst {a[1] / T * 100:5.2f}%  thr/inst {a[2] / max(a[1], 1):5.1f}  "
                  f"stall {a[3] / S * 100:5.2f}%  {a[0].strip()[:80]}")


if __name__ == "__main__":
    main()

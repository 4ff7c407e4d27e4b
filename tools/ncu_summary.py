#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per captured kernel the headline counters
and a per-SASS-segment table of where instructions and stall samples go.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--segments]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(run([rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("=" * 100)
        print(d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:85s} {d[k]:>16s} {u.get(k, '')}")
    if "--segments" not in sys.argv:
        return
    rows = list(csv.reader(run([rep, "--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for k in kernels:
        hdr, body = k["rows"][0], k["rows"][1:]
        ia, it, isamp, isrc = (hdr.index(x) for x in ("Instructions Executed", "Thread Instructions Executed",
                                                        "# Samples", "Source"))
        ti = sum(int(r[ia]) for r in body)
        tt = sum(int(r[it]) for r in body)
        ts = max(1, sum(int(r[isamp]) for r in body))
        print("=" * 100)
        print(k["name"], "SASS lines", len(body), "warp-inst", ti, "threads/inst %.2f" % (tt / max(1, ti)))
        chunk = 40
        for c in range(0, len(body), chunk):
            seg = body[c:c + chunk]
            si = sum(int(r[ia]) for r in seg)
            st = sum(int(r[it]) for r in seg)
            ss = sum(int(r[isamp]) for r in seg)
            if si / max(1, ti) < 0.004 and ss / ts < 0.004:
                continue
            ops = {}
            for r in seg:
                tok = r[isrc].split()
                op = tok[1] if tok and tok[0].startswith('@') and len(tok) > 1 else (tok[0] if tok else '')
                ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
            top = ",".join(f"{a}{b}" for a, b in sorted(ops.items(), key=lambda x: -x[1])[:6])
            print(f"  @{c:5d} inst {si / ti * 100:5.1f}%  thr/inst {st / max(si, 1):5.1f}  stall-samples {ss / ts * 100:5.1f}%  {top}")


if __name__ == "__main__":
    main()

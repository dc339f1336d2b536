#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) for one kernel: key counters and per-region instruction counts."""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sass__inst_executed_local_stores', 'sm__cycles_elapsed.avg']
STALLS = 'smsp__average_warps_issue_stalled_'


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, 'raw')
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {}
    for i, h in enumerate(hdr):
        if h in KEEP:
            m[h] = f"{vals[i]} {units[i]}".strip()
        elif h.startswith(STALLS) and h.endswith('_per_issue_active.ratio'):
            v = float(vals[i])
            if v >= 0.2:
                m['stall_' + h[len(STALLS):-len('_per_issue_active.ratio')]] = round(v, 2)
    src = page(rep, 'source')
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = src[2:]
    grid = int(float(m.get('launch__grid_size', '1').split()[0]))
    rows = [(int(r[ix['Instructions Executed']]), int(r[ix['# Samples']]), r[ix['Source']].strip(),
             r[ix['Avg. Threads Executed']]) for r in data]
    tot = sum(r[0] for r in rows)
    regions = []
    start = 0
    for k in range(1, len(rows) + 1):
        if k == len(rows) or abs(rows[k][0] - rows[start][0]) > 0.15 * max(rows[start][0], 1) + 1000:
            n = sum(r[0] for r in rows[start:k])
            if n / max(tot, 1) >= 0.004:
                ops = collections.Counter((r[2].split()[1] if r[2].startswith('@') else r[2].split()[0]) for r in rows[start:k])
                regions.append({"sass": [start, k], "warp_inst_per_cta": round(n / grid, 1), "pct": round(100 * n / tot, 1),
                                "execs_per_cta": round(rows[start][0] / grid, 2), "avg_threads": rows[start][3],
                                "stall_samples": sum(r[1] for r in rows[start:k]), "top_ops": ops.most_common(6)})
            start = k
    out = {"report": rep, "kernel": src[0][1] if len(src[0]) > 1 else "", "metrics": m,
           "warp_inst_per_cta": round(tot / grid, 1), "regions": regions}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()

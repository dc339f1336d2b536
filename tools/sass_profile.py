#!/usr/bin/env python
"""Per-instruction view of an `ncu --set full --import-source on` capture, read on the CPU box.

    python tools/sass_profile.py gpurun_out/prof.ncu-rep [--kernel decode_kernel] [--tiles N] [--listing]

Prints, for the first matching kernel launch, warp instructions executed by opcode (with the
pipe each opcode issues on, from tools/ubench*.cu), stall samples by opcode, and optionally the
annotated SASS listing (executions per tile, samples).  This is how DESIGN.md attributes issue
slots and ALU-pipe cycles to regions of the decode kernel.
"""
from __future__ import annotations

import argparse
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

# pipe cycles per warp instruction (tools/ubench*.cu on B200): ALU pipe and FMA pipe are both 16 lanes / clk / SMSP
# for most integer work; 2-input VIMNMX / HMNMX2 and FFMA / HFMA2 run at 32 lanes / clk / SMSP.
ALU2 = {"VIMNMX3", "PRMT", "LOP3", "SHF", "IADD3", "IADD", "ISETP", "SEL", "LEA", "IABS", "FLO", "POPC", "BREV", "LOP", "VABSDIFF",
        "ICMP", "FSEL", "PLOP3", "P2R", "R2P", "SGXT", "BMSK", "FMNMX", "IMNMX", "VIADD", "VIADDMNMX"}
ALU1 = {"VIMNMX", "HMNMX2"}
FMA2 = {"IMAD", "IDP", "IDP4A"}
FMA1 = {"FFMA", "HFMA2", "HSET2", "FADD", "FMUL", "HADD2", "HMUL2", "MOV"}


def pipe_of(op: str):
    base = op.split(".")[0]
    if base in ALU1:
        return "alu", 1
    if base in ALU2:
        return "alu", 2
    if base in FMA2:
        return "fma", 2
    if base in FMA1:
        return "fma", 1
    return "other", 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="decode_kernel")
    ap.add_argument("--tiles", type=float, default=0.0, help="divide executions by this (tiles per launch)")
    ap.add_argument("--listing", action="store_true")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    # the output is one CSV block per kernel launch, each introduced by a "Kernel Name" line
    blocks = re.split(r'(?m)^"Kernel Name",', txt)
    block = next((b for b in blocks[1:] if a.kernel in b.split("\n", 1)[0]), None)
    if block is None:
        sys.exit(f"no kernel matching {a.kernel}")
    name, body = block.split("\n", 1)
    rows = list(csv.DictReader(io.StringIO(body)))
    print("kernel:", name.strip().strip('",'))
    by_op = defaultdict(lambda: [0, 0])
    total = samples = 0
    pipes = defaultdict(float)
    for r in rows:
        src = r["Source"].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else src
        n = int(r["Instructions Executed"] or 0)
        s = int(r["# Samples"] or 0)
        base = op.split(".")[0] + (".U16x2" if "U16x2" in op else "")
        by_op[base][0] += n
        by_op[base][1] += s
        total += n
        samples += s
        p, c = pipe_of(op)
        pipes[p] += n * c
    div = a.tiles or 1.0
    print(f"warp instructions: {total}  ({total / div:.1f} per tile)   stall samples: {samples}")
    print(f"pipe cycles per tile: ALU {pipes['alu'] / div:.1f}  FMA {pipes['fma'] / div:.1f}")
    print(f"{'opcode':<18}{'executed':>14}{'per tile':>10}{'%':>7}{'samples%':>9}")
    for op, (n, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
        if n == 0 and s == 0:
            continue
        print(f"{op:<18}{n:>14}{n / div:>10.2f}{100.0 * n / total:>7.2f}{100.0 * s / max(samples, 1):>9.2f}")
    if a.listing:
        print()
        for k, r in enumerate(rows):
            n = int(r["Instructions Executed"] or 0)
            s = int(r["# Samples"] or 0)
            print(f"{k:5d} {n / div:9.3f} {s:7d}  {r['Source'].strip()}")


if __name__ == "__main__":
    main()

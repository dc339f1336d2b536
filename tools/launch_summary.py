#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel, launches, total ms, largest launches."""
import collections
import csv
import json
import sys


def main():
    path, command = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    per = collections.defaultdict(list)
    for r in rows:
        name = r[4].split("(")[0].strip()
        per[name].append(float(r[14]) / 1e6)
    out = {"command": command,
           "note": "ncu times are cold-cache and serialised: compare shares, not absolutes. The few largest decode / "
                   "group_scan / gather launches are the warm-up + timed steps on the 1 h capture; the many small ones "
                   "are the 32 Mi-sample pieces of the e2e (host buffer) leg; compose / scatter are the synthetic generator.",
           "kernels": {k: {"launches": len(v), "total_ms": round(sum(v), 3), "largest_ms": [round(x, 4) for x in sorted(v)[-8:]]}
                       for k, v in per.items()}}
    dec = [v for k, v in per.items() if "decode_kernel" in k]
    fin = [v for k, v in per.items() if "gather_kernel" in k or "group_scan" in k]
    if dec:
        big = sorted(dec[0])[-8:]
        rest = sum(sorted(v)[-8:][-1] for v in fin)
        out["step_share"] = {"decode_kernel_ms": round(big[-1], 4), "group_scan_plus_gather_ms": round(rest, 4),
                             "decode_share_of_step": round(big[-1] / (big[-1] + rest), 4)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

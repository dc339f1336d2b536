#!/usr/bin/env python
"""parity of the WHOLE 1 h capture (8.64 G samples u8), once.
The GPU decodes the capture resident in HBM in one call; the chunk-parallel fast oracle (proven equal to the
literal one in tests/test_oracle.py) decodes it on the host cores slice by slice (each slice carries its
240-sample halo, so the concatenation is the single-buffer result); records and the gate-pass counter must
be identical.  Prints one JSON object.  The oracle is the checker here, nothing else."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402  (workload definition only)
from air_rs_b200 import synth  # noqa: E402
from air_rs_b200.decoder import AdsbDecoder  # noqa: E402
from air_rs_b200.native import FMT_U8, FRAME_DTYPE  # noqa: E402
from oracle import oracle_c  # noqa: E402


def main():
    total = bench.TOTAL_SAMPLES
    slice_n = int(os.environ.get("AIRGPU_PARITY_SLICE", 540_000_000))
    threads = os.cpu_count() or 1
    gen = synth.DeviceSynth(bench.traffic_table())
    iq = gen.render(bench.SEED, 0, total, FMT_U8, bench.SIGMA, period=bench.PERIOD)
    torch.cuda.synchronize()
    dec = AdsbDecoder(fmt=FMT_U8)
    cap = total // 240
    out = torch.empty((cap, FRAME_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    dec.decode_device(iq.data_ptr(), total, out.data_ptr(), cap, 0, 0, 0, s.cuda_stream)   # count kept by the library
    n_gpu = dec.sync_count()                                                                # waits, mirrors the counters
    gpu = out[:n_gpu].cpu().numpy().view(FRAME_DTYPE).reshape(-1)
    gpu_gate = dec.stats()["gate_passes"]

    L = oracle_c.lib()
    parts, cpu_gate, t_cpu = [], 0, 0.0
    a = 0
    while a < total - 240:
        n = min(slice_n + 240, total - a)
        host = iq[2 * a: 2 * (a + n)].cpu().numpy()
        buf = np.zeros(max(1 << 16, n // 200), dtype=FRAME_DTYPE)
        gp = C.c_uint64(0)
        t0 = time.perf_counter()
        got = L.oracle_decode_fast(host.ctypes.data, n, FMT_U8, 0, a, buf.ctypes.data, len(buf), C.byref(gp), threads)
        t_cpu += time.perf_counter() - t0
        assert got <= len(buf), "oracle capacity"
        parts.append(buf[:got].copy())
        cpu_gate += int(gp.value)
        a += slice_n
    cpu = np.concatenate(parts)
    equal = cpu.shape == gpu.shape and cpu.tobytes() == gpu.tobytes()
    first_diff = None
    if not equal:
        m = min(len(cpu), len(gpu))
        d = np.nonzero(cpu[:m].view(np.uint8).reshape(m, -1) != gpu[:m].view(np.uint8).reshape(m, -1))[0]
        first_diff = int(d[0]) if len(d) else m
    print(json.dumps({
        "workload": bench.workload_config(1, total)["workload"], "samples": total,
        "gpu_frames": n_gpu, "cpu_frames": int(len(cpu)), "records_identical": bool(equal), "first_difference": first_diff,
        "gpu_gate_passes": int(gpu_gate), "cpu_gate_passes": cpu_gate, "gate_passes_identical": int(gpu_gate) == cpu_gate,
        "repaired_frames": int((gpu["fixed_bit"] != 0xFF).sum()),
        "oracle": f"oracle_decode_fast, {threads} threads, slices of {slice_n} samples + 240 halo",
        "oracle_seconds": round(t_cpu, 2), "oracle_msamples_per_s": round(total / t_cpu / 1e6, 1)}))
    return 0 if equal and int(gpu_gate) == cpu_gate else 1


if __name__ == "__main__":
    sys.exit(main())

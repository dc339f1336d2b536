#!/usr/bin/env python
"""CS16 decode of a dense 1.2 G-sample capture, a few passes (profiling target)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from air_rs_b200 import synth
from air_rs_b200.decoder import AdsbDecoder
from air_rs_b200.native import FMT_CS16
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_200_000_000
tab = synth.make_traffic(1090, 24_000_000, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30), sigma=300.0)
iq = synth.DeviceSynth(tab).render(1090, 0, n, FMT_CS16, 300.0, period=24_000_000)
dec = AdsbDecoder(fmt=FMT_CS16)
cap = n // 240
out = torch.empty((cap, 24), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
s = torch.cuda.Stream()
for _ in range(4):
    dec.decode_device(iq.data_ptr(), n, out.data_ptr(), cap, 0, 0, cnt.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    print(dec.stats()["decode_ms"], int(cnt.item()))

"""CPU: the packed two-stream preamble gate of the decode kernel (csrc/airgpu_scan.cuh), compiled for the
host, against a direct evaluation of the reference gate (demod.rs:17-44) on random level arrays -- the bf16
compare of U8 levels and the generic 16-bit compare of CS16 levels -- plus the hit-bit layout, the shared-memory bank-conflict pattern, and the
addresses the DF test and the bit slicer read (the same index helpers the kernel compiles)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_gate_scan_host_emulation(tmp_path):
    exe = tmp_path / "emu_scan"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tools" / "emu_scan.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gate_scan<bf16>" in r.stdout and "gate_scan<u16>" in r.stdout and r.stdout.count("all equal") == 2
    assert "64 distinct of 64" in r.stdout
    assert "phase-1 stores: worst 1-way" in r.stdout and "phase-2 loads: worst 1-way" in r.stdout
    assert "level reads at the right address" in r.stdout

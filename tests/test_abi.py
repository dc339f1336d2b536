"""CPU tests of the boundary: the shared library builds for sm_100a, loads, exports
every symbol the headers declare, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from air_rs_b200 import build, native
from air_rs_b200.decoder import AdsbDecoder

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(airgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert path.exists() and path.parent == ROOT / "air_rs_b200" / "_lib"
    lib = native.lib()
    assert b"sm_100a" in lib.airgpu_version()


@pytest.mark.parametrize("header", ["airgpu.h", "airgpu_synth.h"])
def test_every_declared_symbol_is_exported(header):
    names = declared_symbols(ROOT / "include" / header)
    assert len(names) >= 4
    raw = C.CDLL(str(build.build()))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in {header} but not exported: {missing}"
    bound = set(native.SYMBOLS) | set(native.SYNTH_SYMBOLS)
    assert set(names) <= bound, f"not bound in native.py: {sorted(set(names) - bound)}"


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers must compile as strict C99 and as C++ with nothing else included."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "include/airgpu.h"\n#include "include/airgpu_synth.h"\n'
                   'int main(void) { airgpu_config c; airgpu_frame f; (void)c; (void)f; return sizeof(airgpu_frame) == 24 ? 0 : 1; }\n')
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + [f"-I{ROOT}", str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_frame_record_layout():
    assert native.FRAME_DTYPE.itemsize == 24
    assert native.FRAME_DTYPE.fields["fixed_bit"][1] == 14
    assert native.FRAME_DTYPE.fields["offset"][1] == 16
    assert C.sizeof(native.Config) == 32


def test_sass_is_sm100_and_uses_packed_minmax():
    out = subprocess.run(["cuobjdump", "-sass", str(build.build())], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "VIMNMX3.U16x2" in out and "IDP.4A" in out
    assert "airgpu" in out
    # the U8 decode kernel: compare and hit-bit gathering on the FMA pipe, 16-byte streaming loads, no local memory
    k = re.search(r"Function : \S*decode_kernelILi1ELb1E.*?(?=Function :|\Z)", out, flags=re.S).group(0)
    assert "HSET2.BF16_V2" in k and "IDP.4A.S8.U8" in k and "LDG.E.NA.128" in k
    assert " STL" not in k and " LDL" not in k
    assert k.count("VIMNMX.U16x2") > 200 and "REDUX" in k


def test_no_cpu_fallback(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present; the fallback check is for GPU-less machines")
    with pytest.raises(native.AirgpuError) as ei:
        AdsbDecoder()
    assert ei.value.code == native.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = ROOT / "air_rs_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "adsb_oracle" not in text and "oracle_c" not in text and "oracle_np" not in text, f


def test_c_consumer_links_and_runs(tmp_path):
    """A C99 program compiled against include/airgpu.h and linked with libairgpu.so runs: on a box without a GPU it
    sees AIRGPU_ERR_NO_DEVICE (there is no CPU fallback), on a B200 it decodes through the ABI."""
    lib = build.build()
    exe = tmp_path / "c_consumer"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "tests" / "c_consumer.c"), "-o", str(exe), f"-L{lib.parent}", "-lairgpu",
                    f"-Wl,-rpath,{lib.parent}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi 2" in r.stdout

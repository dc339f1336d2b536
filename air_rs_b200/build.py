"""In-tree nvcc build of libairgpu.so (sm_100a only)."""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
from contextlib import contextmanager
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
LIB = LIBDIR / os.environ.get("AIRGPU_LIB", "libairgpu.so")   # AIRGPU_LIB: A/B-test a prebuilt variant

SOURCES = ["airgpu_kernels.cu", "airgpu_api.cu", "airgpu_synth.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libairgpu.so cannot be built")
    return exe


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = ([CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) +
            [PKG.parent / "include" / "airgpu.h", PKG.parent / "include" / "airgpu_synth.h"])
    return any(d.exists() and d.stat().st_mtime > t for d in deps)


@contextmanager
def _build_lock():
    """One builder at a time: every rank of a torchrun job imports the package (and may find the library stale) at
    the same moment.  The others wait here and then find it fresh."""
    LIBDIR.mkdir(exist_ok=True)
    with open(LIBDIR / ".build.lock", "w") as fh:
        fcntl.flock(fh, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(fh, fcntl.LOCK_UN)


def _compile(cmd, out: Path) -> subprocess.CompletedProcess:
    """Run a compiler that writes to a temporary file next to `out`, then move it into place atomically: a process
    that is loading the library never sees a half-written file."""
    tmp = out.with_name(f".{out.name}.{os.getpid()}.tmp")
    full = [c if c != "@OUT@" else str(tmp) for c in cmd]
    r = subprocess.run(full, capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"{full[0]} failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, out)
    return r


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source into air_rs_b200/_lib/libairgpu.so."""
    if (not force and not is_stale()) or "AIRGPU_LIB" in os.environ:
        return LIB
    with _build_lock():
        if not force and not is_stale():           # another process built it while this one waited
            return LIB
        srcs = [str(CSRC / s) for s in SOURCES]
        cmd = [nvcc(), *NVCC_FLAGS, "-o", "@OUT@", *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = _compile(cmd, LIB)
        if verbose:
            print(r.stderr)
    return LIB


HOST_BIN = LIBDIR / "airgpu_playback"


def build_host(force: bool = False) -> Path:
    """Compile the C++ three-thread harness (csrc/host) against libairgpu.so."""
    src = CSRC / "host" / "airgpu_playback.cpp"
    hdr = CSRC / "host" / "adsb_host.hpp"
    build()
    stale = (not HOST_BIN.exists()) or HOST_BIN.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime,
                                                                      LIB.stat().st_mtime)
    if force or stale:
        with _build_lock():
            _compile(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", str(src), "-o", "@OUT@",
                      f"-L{LIBDIR}", "-lairgpu", "-Wl,-rpath,$ORIGIN"], HOST_BIN)
    return HOST_BIN


HOST_LIB = LIBDIR / "libadsb_host.so"


def build_host_lib(force: bool = False) -> Path:
    """Compile the CUDA-free host mirror of the tracker / CPR / JSON summary (csrc/host) with its C shims."""
    srcs = [CSRC / "host" / "adsb_host_c.cpp", CSRC / "host" / "adsb_track.hpp", CSRC / "host" / "adsb_host.hpp"]
    LIBDIR.mkdir(exist_ok=True)
    stale = (not HOST_LIB.exists()) or HOST_LIB.stat().st_mtime < max(f.stat().st_mtime for f in srcs)
    if force or stale:
        with _build_lock():
            _compile(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", "@OUT@", str(srcs[0])], HOST_LIB)
    return HOST_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_host(force=True))
    print(build_host_lib(force=True))

"""Synthetic 1090 MHz captures: integer-only, counter-based, host == device bit for bit.

This is the workload generator SURVEY.md section 8(d) specifies (the reference
ships no IQ fixture: its author's capture is git-ignored, reference .gitignore:4).
Everything that decides a sample value is integer arithmetic on a 64-bit
counter hash, so the numpy renderer here and the CUDA renderer in
csrc/airgpu_synth.cu produce identical bytes for the same (table, seed, range).

Timing follows the reference demodulator, which hard-codes 2 samples per
microsecond (reference src/adsb/demod.rs:20-24, src/adsb.rs:98-106): preamble
pulses at samples 0, 2, 7, 9 and data bit k at sample 16+2k (bit 1) or 17+2k
(bit 0).  "2.4 MS/s" in BASELINE.json therefore fixes sample COUNTS only.

Sample model (per component c in {I, Q}, sample index j, all integers):
    g   = sum of the 8 bytes of mix64(seed * GOLDEN + 2*j + c)        in [0, 2040]
    n   = ((g - 1020) * K) >> 16            (arithmetic shift; Irwin-Hall(8) noise)
    s   = sum over pulses landing on j of the frame's amplitude component
    U8  : clip(128 + n + s, 0, 255)         (symmetric about 127.5)
    CS16: clip(n + s, -32768, 32767)
with K = round(sigma * 65536 / sqrt(43690)).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
SIGMA_UNIT = 43690.0 ** 0.5  # std of the sum of 8 uniform bytes

FMT_CS16 = 0
FMT_U8 = 1

PREAMBLE_PULSES = (0, 2, 7, 9)
FRAME_SAMPLES_LONG = 16 + 112 * 2
FRAME_SAMPLES_SHORT = 16 + 56 * 2

# The seven CRC-valid DF17 frames the reference's own tests carry
# (reference src/adsb/aircraft.rs:188-261, src/adsb/demod.rs:339-344).
GOLDEN_FRAMES = (
    "8d7c6b3020293532d70820fc8090",
    "8d7c6b30581304f388bb4455896f",
    "8D40621D58C386435CC412692AD6",
    "8D40621D58C382D690C8AC2863A7",
    "8d7c6b30580d107903b3cabf62ab",
    "8d7c6b30580d24eeaebb2dfea5bb",
    "8D406B902015A678D4D220AA4BDA",
)


def noise_gain(sigma: float) -> int:
    """K in the sample model for a target noise std of `sigma` LSB."""
    return int(round(sigma * 65536.0 / SIGMA_UNIT))


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= _M1
    x ^= x >> np.uint64(27)
    x *= _M2
    x ^= x >> np.uint64(31)
    return x


def noise(seed: int, j0: int, n: int, gain: int) -> np.ndarray:
    """(n, 2) int32 noise for samples [j0, j0+n)."""
    with np.errstate(over="ignore"):
        ctr = (np.uint64(seed) * GOLDEN) + (np.uint64(2) * (np.uint64(j0) + np.arange(n, dtype=np.uint64)))[:, None] \
            + np.arange(2, dtype=np.uint64)[None, :]
        h = mix64(ctr)
    g = h.view(np.uint8).reshape(n, 2, 8).sum(axis=2, dtype=np.int64) - 1020
    return ((g * gain) >> 16).astype(np.int32)


def crc24(data: bytes) -> int:
    """Mode S parity, generator 0x1FFF409 (same polynomial as reference crc.rs:11)."""
    crc = 0
    for b in data:
        crc ^= b << 16
        for _ in range(8):
            crc = ((crc << 1) ^ 0xFFF409) & 0xFFFFFF if crc & 0x800000 else (crc << 1) & 0xFFFFFF
    return crc


@dataclass
class FrameTable:
    """Injected transmissions, sorted by start sample."""

    start: np.ndarray  # int64 [F]  first preamble sample (absolute index)
    nbits: np.ndarray  # int32 [F]  112 or 56
    payload: np.ndarray  # uint8 [F, 14] MSB-first bits (short frames use 7 bytes)
    amp_i: np.ndarray  # int32 [F]  per-pulse I amplitude (LSB of the output format)
    amp_q: np.ndarray  # int32 [F]
    smear: np.ndarray  # uint8 [F]  1 = half-sample late: each pulse splits over 2 samples
    kind: np.ndarray = field(default=None)  # uint8 [F] downlink format, informational

    def __post_init__(self):
        self.start = np.ascontiguousarray(self.start, dtype=np.int64)
        self.nbits = np.ascontiguousarray(self.nbits, dtype=np.int32)
        self.payload = np.ascontiguousarray(self.payload, dtype=np.uint8).reshape(-1, 14)
        self.amp_i = np.ascontiguousarray(self.amp_i, dtype=np.int32)
        self.amp_q = np.ascontiguousarray(self.amp_q, dtype=np.int32)
        self.smear = np.ascontiguousarray(self.smear, dtype=np.uint8)
        if self.kind is None:
            self.kind = (self.payload[:, 0] >> 3).astype(np.uint8)

    def __len__(self) -> int:
        return int(self.start.size)

    @staticmethod
    def empty() -> "FrameTable":
        z = np.zeros(0, dtype=np.int64)
        return FrameTable(z, z, np.zeros((0, 14), np.uint8), z, z, z)

    @staticmethod
    def concat(tables) -> "FrameTable":
        tables = [t for t in tables if len(t)]
        if not tables:
            return FrameTable.empty()
        t = FrameTable(
            np.concatenate([t.start for t in tables]),
            np.concatenate([t.nbits for t in tables]),
            np.concatenate([t.payload for t in tables]),
            np.concatenate([t.amp_i for t in tables]),
            np.concatenate([t.amp_q for t in tables]),
            np.concatenate([t.smear for t in tables]),
            np.concatenate([t.kind for t in tables]),
        )
        order = np.argsort(t.start, kind="stable")
        return t.take(order)

    def take(self, idx) -> "FrameTable":
        return FrameTable(self.start[idx], self.nbits[idx], self.payload[idx], self.amp_i[idx],
                          self.amp_q[idx], self.smear[idx], self.kind[idx])

    def pulses(self):
        """(frame index [P], sample offset within the frame [P]) of every pulse."""
        f = len(self)
        bits = np.unpackbits(self.payload, axis=1)[:, :112]  # [F,112]
        k = np.arange(112)
        data_pos = 16 + 2 * k[None, :] + (1 - bits)  # bit 1 -> first half, bit 0 -> second half
        valid = k[None, :] < self.nbits[:, None]
        fi = np.repeat(np.arange(f), 4)
        po = np.tile(np.array(PREAMBLE_PULSES), f)
        fd, kd = np.nonzero(valid)
        return np.concatenate([fi, fd]), np.concatenate([po, data_pos[fd, kd]])


def signal(table: FrameTable, j0: int, n: int) -> np.ndarray:
    """(n, 2) int64 summed pulse amplitudes for samples [j0, j0+n)."""
    sig = np.zeros((n, 2), dtype=np.int64)
    if len(table) == 0:
        return sig
    sel = np.flatnonzero((table.start < j0 + n) & (table.start + FRAME_SAMPLES_LONG + 1 > j0))
    if sel.size == 0:
        return sig
    t = table.take(sel)
    fi, po = t.pulses()
    pos = t.start[fi] + po - j0
    for comp, amp in ((0, t.amp_i), (1, t.amp_q)):
        a = amp[fi].astype(np.int64)
        sm = t.smear[fi].astype(bool)
        late = np.where(sm, a // 2, 0)  # floor division, same as the device's arithmetic
        early = a - late
        for p, v in ((pos, early), (pos[sm] + 1, late[sm])):
            ok = (p >= 0) & (p < n)
            # bincount on float64 weights is exact for these small integers
            sig[:, comp] += np.rint(np.bincount(p[ok], weights=v[ok].astype(np.float64), minlength=n)).astype(np.int64)
    return sig


def render(table: FrameTable, seed: int, j0: int, n: int, fmt: int = FMT_U8, sigma: float = 2.0,
           period: int = 0) -> np.ndarray:
    """Interleaved IQ for samples [j0, j0+n): uint8 (U8) or int16 (CS16), shape (2n,).

    period > 0 repeats the frame schedule every `period` samples (the noise does
    not repeat: it is hashed on the absolute sample index).
    """
    gain = noise_gain(sigma)
    out = noise(seed, j0, n, gain).astype(np.int64)
    if period:
        assert period >= FRAME_SAMPLES_LONG + 2
        first = j0 // period
        last = (j0 + n - 1) // period
        for rep in range(first - 1 if first > 0 else first, last + 1):
            out += signal(table, j0 - rep * period, n)
    else:
        out += signal(table, j0, n)
    if fmt == FMT_U8:
        return np.clip(out + 128, 0, 255).astype(np.uint8).reshape(-1)
    return np.clip(out, -32768, 32767).astype(np.int16).reshape(-1)


# --------------------------------------------------------------------------- #
# Traffic: what gets injected                                                 #
# --------------------------------------------------------------------------- #

def df17_frame(icao: int, me: bytes, ca: int = 5) -> bytes:
    body = bytes([(17 << 3) | ca, (icao >> 16) & 0xFF, (icao >> 8) & 0xFF, icao & 0xFF]) + me
    return body + crc24(body).to_bytes(3, "big")


def _overlaid(df: int, body: bytes, icao: int) -> bytes:
    """DF4/5/20/21-style frame: parity field = CRC xor ICAO address."""
    ap = crc24(body) ^ icao
    return body + ap.to_bytes(3, "big")


def make_traffic(seed: int, n_samples: int, df17_per_s: float = 200.0, decoy_per_s: float = 0.0,
                 snr_db=(20.0, 20.0), sigma: float = 2.0, n_icao: int = 500, sample_rate: float = 2.4e6,
                 include_golden: bool = True, smear_fraction: float = 0.0, amp_scale: float = 1.0) -> FrameTable:
    """Poisson arrivals of DF17 squitters plus rejected-by-design decoys.

    DF17 payloads: type codes 1-4 (identification), 9-18 (airborne position,
    alternating CPR parity) and 19 (velocity) from a fixed ICAO pool; the seven
    golden frames are cycled in verbatim.  Decoys: DF4/5 (56 bit), DF11 (56 bit),
    DF20/21 (112 bit) -- the reference gate only passes first-five-bits 10001.
    Amplitude A = sigma * 10^(snr/20) at a uniform random carrier phase.
    """
    rng = np.random.default_rng(seed)
    icaos = rng.integers(0x100000, 0xFFFFFF, size=n_icao)
    dur = n_samples / sample_rate

    def arrivals(rate):
        if rate <= 0:
            return np.zeros(0, dtype=np.int64)
        k = rng.poisson(rate * dur)
        t = np.sort(rng.integers(0, max(1, n_samples - FRAME_SAMPLES_LONG - 2), size=k))
        return t.astype(np.int64)

    starts17 = arrivals(df17_per_s)
    startsdc = arrivals(decoy_per_s)
    golden = [bytes.fromhex(h) for h in GOLDEN_FRAMES]

    payload = np.zeros((starts17.size + startsdc.size, 14), dtype=np.uint8)
    nbits = np.zeros(payload.shape[0], dtype=np.int32)
    cpr_odd = {}
    for k in range(starts17.size):
        if include_golden and k % 64 == 0:
            frame = golden[(k // 64) % len(golden)]
        else:
            icao = int(icaos[rng.integers(0, n_icao)])
            r = rng.random()
            me = bytearray(rng.integers(0, 256, size=7, dtype=np.uint8).tobytes())
            if r < 0.08:
                tc = int(rng.integers(1, 5))
            elif r < 0.6:
                tc = int(rng.integers(9, 19))
                odd = cpr_odd.get(icao, 0)
                cpr_odd[icao] = odd ^ 1
                me[2] = (me[2] & ~0x04) | (odd << 2)
            else:
                tc = 19
            me[0] = (tc << 3) | (me[0] & 7)
            frame = df17_frame(icao, bytes(me))
        payload[k] = np.frombuffer(frame, dtype=np.uint8)
        nbits[k] = 112
    for k in range(startsdc.size):
        icao = int(icaos[rng.integers(0, n_icao)])
        df = int(rng.choice([4, 5, 11, 20, 21]))
        row = starts17.size + k
        if df == 11:
            body = bytes([(11 << 3) | 5, (icao >> 16) & 0xFF, (icao >> 8) & 0xFF, icao & 0xFF])
            frame = body + crc24(body).to_bytes(3, "big")
        elif df in (4, 5):
            body = bytes([(df << 3) | int(rng.integers(0, 8))]) + rng.integers(0, 256, size=3, dtype=np.uint8).tobytes()
            frame = _overlaid(df, body, icao)
        else:
            body = bytes([(df << 3) | int(rng.integers(0, 8))]) + rng.integers(0, 256, size=10, dtype=np.uint8).tobytes()
            frame = _overlaid(df, body, icao)
        payload[row, : len(frame)] = np.frombuffer(frame, dtype=np.uint8)
        nbits[row] = len(frame) * 8

    start = np.concatenate([starts17, startsdc])
    f = start.size
    snr = rng.uniform(snr_db[0], snr_db[1], size=f)
    amp = sigma * 10.0 ** (snr / 20.0) * amp_scale
    phase = rng.uniform(0.0, 2.0 * np.pi, size=f)
    table = FrameTable(
        start, nbits, payload,
        np.rint(amp * np.cos(phase)).astype(np.int32),
        np.rint(amp * np.sin(phase)).astype(np.int32),
        (rng.random(f) < smear_fraction).astype(np.uint8),
    )
    return table.take(np.argsort(table.start, kind="stable"))


def single_frames(frames, starts, amp_i, amp_q=0, smear=0) -> FrameTable:
    """Table with the given 14- or 7-byte frames at the given sample offsets."""
    f = len(frames)
    payload = np.zeros((f, 14), dtype=np.uint8)
    nbits = np.zeros(f, dtype=np.int32)
    for k, fr in enumerate(frames):
        fr = bytes.fromhex(fr) if isinstance(fr, str) else bytes(fr)
        payload[k, : len(fr)] = np.frombuffer(fr, dtype=np.uint8)
        nbits[k] = len(fr) * 8
    return FrameTable(np.asarray(starts), nbits, payload, np.broadcast_to(amp_i, (f,)).copy(),
                      np.broadcast_to(amp_q, (f,)).copy(), np.broadcast_to(smear, (f,)).copy())


# --------------------------------------------------------------------------- #
# Device twin (csrc/airgpu_synth.cu)                                          #
# --------------------------------------------------------------------------- #

class DeviceSynth:
    """Frame table resident on one GPU; renders captures straight into HBM."""

    def __init__(self, table: FrameTable, device: int = 0):
        import ctypes as C

        from . import native

        self._native = native
        self._lib = native.lib()
        self.device = device
        h = C.c_void_p()
        t = table
        native.check_synth(self._lib.airgpu_synth_table_create(
            device, t.start.ctypes.data, t.nbits.ctypes.data, t.payload.ctypes.data, t.amp_i.ctypes.data,
            t.amp_q.ctypes.data, t.smear.ctypes.data, len(t), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.airgpu_synth_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_into(self, out, seed: int, j0: int, n: int, fmt: int = FMT_U8, sigma: float = 2.0,
                    period: int = 0, stream: int = 0):
        """out: CUDA torch tensor with room for n samples (2n uint8 or 2n int16)."""
        self._native.check_synth(self._lib.airgpu_synth_render(
            self._h, seed, j0, n, fmt, noise_gain(sigma), period, out.data_ptr(), stream or None))
        return out

    def render(self, seed: int, j0: int, n: int, fmt: int = FMT_U8, sigma: float = 2.0, period: int = 0):
        import torch

        dt = torch.uint8 if fmt == FMT_U8 else torch.int16
        out = torch.empty(2 * n, dtype=dt, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(out.device).cuda_stream
        self.render_into(out, seed, j0, n, fmt, sigma, period, stream)
        return out

"""GPU: the parts of the C ABI beyond plain decode -- ring overflow handling, workspace reservation, CUDA-graph
capture, the multi-destination (peer) form of the ordering kernels, and the one-thread multi-GPU group
(SURVEY 8(b) `airgpu_decode_sharded`).  Everything is compared with the CPU oracle or with a plain decode."""
import ctypes as C

import numpy as np
import pytest
import torch

from air_rs_b200 import native, synth
from air_rs_b200.decoder import AdsbDecoder, DecoderGroup
from air_rs_b200.native import FMT_CS16, FMT_U8, FRAME_DTYPE
from oracle import oracle_c

from common import capture_cs16, capture_u8, describe_diff, frames_equal

pytestmark = pytest.mark.gpu


def test_ring_never_drops_frames_of_a_constant_buffer():
    """the default ring capacity (8192) is below the per-buffer worst case (n - 240 frames for a constant
    20 000-sample chunk, all of which the reference sends).  The device side of the ring now always has room for the
    worst case, airgpu_collect reports the count and keeps the ticket, and the mirror collects again with room."""
    with AdsbDecoder(fmt=FMT_CS16, max_buffer_samples=20_000) as dec:            # default max_frames
        assert dec.max_frames == 8192
        t1 = dec.submit(np.zeros(2 * 20_000, dtype=np.int16), base_offset=7)
        _, noisy = capture_cs16(seed=5, n=20_000, df17=5000.0)
        t2 = dec.submit(noisy, base_offset=20_007)
        # raw ABI: too small an array -> OVERFLOW with the count, ticket still collectable
        small = np.zeros(10, dtype=FRAME_DTYPE)
        n = C.c_size_t(0)
        rc = dec._lib.airgpu_collect(dec._h, t1, small.ctypes.data, small.size, C.byref(n))
        assert rc == native.ERR_OVERFLOW and n.value == 19_760
        got = dec.collect(t1)
        want, _ = oracle_c.decode_literal(np.zeros(2 * 20_000, dtype=np.int16), 0, 7)
        assert len(got) == 19_760 and frames_equal(got, want), describe_diff(got, want)
        got2 = dec.collect(t2)
        want2, _ = oracle_c.decode_fast(noisy, 0, 20_007)
        assert frames_equal(got2, want2), describe_diff(got2, want2)


def test_ring_head_and_tail_records():
    """A buffer with more frames than the fixed head (256) that travels with the count: the rest is fetched on demand."""
    n = 100_000
    frames = [bytes.fromhex(synth.GOLDEN_FRAMES[k % len(synth.GOLDEN_FRAMES)]) for k in range(390)]
    iq = synth.render(synth.single_frames(frames, [100 + 250 * k for k in range(390)], amp_i=40), 6, 0, n, FMT_U8, 1.0)
    with AdsbDecoder(fmt=FMT_U8, max_buffer_samples=n, max_frames=4096) as dec:
        got = dec.collect(dec.submit(iq))
        want, gp = oracle_c.decode_fast(iq)
        assert len(want) > 300 and frames_equal(got, want), describe_diff(got, want)
        assert dec.stats()["gate_passes"] == gp


def test_reserve_then_no_growth_and_graph_replay():
    """airgpu_reserve sizes the workspace; a captured airgpu_decode_device replays on new data in the same buffers."""
    n = 600_000
    caps = [capture_u8(seed=20 + k, n=n, df17=4000.0)[1] for k in range(3)]
    with AdsbDecoder(fmt=FMT_U8) as dec:
        cap = 1 << 14
        dec.reserve(n, 0, cap)
        d_iq = torch.zeros(2 * n, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros((cap, 24), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            dec.graph_begin(s.cuda_stream)
            dec.decode_device(d_iq.data_ptr(), n, d_out.data_ptr(), cap, 0, 1000, d_cnt.data_ptr(), s.cuda_stream)
            g = dec.graph_end(s.cuda_stream)
            for iq in caps:
                d_iq.copy_(torch.from_numpy(iq), non_blocking=False)
                s.wait_stream(torch.cuda.default_stream())
                g.launch(s.cuda_stream)
                s.synchronize()
                got = AdsbDecoder.frames_from_tensor(d_out, int(d_cnt.item()))
                want, _ = oracle_c.decode_fast(iq, 0, 1000)
                assert frames_equal(got, want), describe_diff(got, want)
            g.close()
        # a capture that would need a larger workspace fails instead of allocating
        big = torch.zeros(2 * 8 * n, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(s):
            dec.graph_begin(s.cuda_stream)
            with pytest.raises(native.AirgpuError):
                dec.decode_device(big.data_ptr(), 8 * n, d_out.data_ptr(), cap, 0, 0, d_cnt.data_ptr(), s.cuda_stream)
            dec.graph_end(s.cuda_stream).close()


@pytest.mark.parametrize("fmt,maker", [(FMT_U8, capture_u8), (FMT_CS16, capture_cs16)])
def test_decode_device_peers_local_destinations(fmt, maker):
    """The fused-exchange form of the ordering kernels with every destination in local memory: each destination gets
    the same records and the same count as a plain decode (1, 3 and 8 destinations; the multi-GPU test drives peer
    and multicast addresses)."""
    n = 500_000
    _, iq = maker(seed=31, n=n, df17=4000.0)
    want, gp = oracle_c.decode_fast(iq, 0, 55)
    d_iq = torch.from_numpy(iq).cuda()
    cap = 1 << 13
    with AdsbDecoder(fmt=fmt) as dec:
        s = torch.cuda.Stream()
        flags = torch.zeros(8, dtype=torch.int64, device="cuda")
        for epoch, k in enumerate((1, 3, 8), start=1):
            slab = torch.zeros((k, cap + 1, 24), dtype=torch.uint8, device="cuda")
            outs = [slab[j, 1:].data_ptr() for j in range(k)]
            counts = [slab[j].data_ptr() for j in range(k)]
            with torch.cuda.stream(s):
                dec.decode_device_peers(d_iq.data_ptr(), n, outs, counts, cap, 0, 55, s.cuda_stream)
                dec.peer_barrier([flags.data_ptr()], 0, epoch, s.cuda_stream)   # one rank: publishes its epoch and returns
                s.synchronize()
            assert int(flags[0].item()) == epoch
            assert dec.sync_count() == len(want) and dec.stats()["gate_passes"] == gp
            host = slab.cpu().numpy()
            for j in range(k):
                assert int(host[j, 0, :8].view(np.uint64)[0]) == len(want)
                got = host[j, 1:1 + len(want)].reshape(-1).view(FRAME_DTYPE)
                assert frames_equal(got, want), describe_diff(got, want)


@pytest.mark.parametrize("fmt,maker", [(FMT_U8, capture_u8), (FMT_CS16, capture_cs16)])
def test_group_decode_equals_single_decode(fmt, maker):
    """airgpu_group_decode over (0,), (0, 0), (0, 0, 0) and every visible GPU == airgpu_decode on one == the oracle;
    a capture shorter than one shard, an empty one, the overflow report, airgpu_decode_sharded."""
    n = 3_000_000
    _, iq = maker(seed=41, n=n, df17=3000.0)
    want, _ = oracle_c.decode_fast(iq, 0, 9, threads=4)
    ndev = native.lib().airgpu_device_count()
    for devices in ([0], [0, 0], [0, 0, 0], list(range(ndev)), list(range(ndev)) * 2):
        with DecoderGroup(devices, fmt=fmt) as grp:
            got = grp.decode(iq, base_offset=9)
            assert frames_equal(got, want), (devices, describe_diff(got, want))
            st = grp.stats()
            assert sum(x["n_frames"] for x in st) == len(want) and len(st) == len(devices)
            small = iq[: 2 * 5_000]
            assert frames_equal(grp.decode(small), oracle_c.decode_fast(small)[0])
            assert len(grp.decode(iq[:0])) == 0
            with pytest.raises(native.AirgpuError) as ei:
                grp.decode(iq, max_frames=10)
            assert ei.value.code == native.ERR_OVERFLOW
    # the one-call form
    out = np.zeros(len(want) + 10, dtype=FRAME_DTYPE)
    got_n = C.c_size_t(0)
    devs = (C.c_int * 2)(0, 0)
    native.check(native.lib().airgpu_decode_sharded(devs, 2, fmt, iq.ctypes.data, n, 9, out.ctypes.data, out.size, C.byref(got_n)))
    assert frames_equal(out[: got_n.value], want)


def test_playback_samples_matches_the_reference_loop():
    """adsb.rs:77: `while i < data.len() - 20000` -- the last chunk is never sent, complete or not."""
    L = native.lib()
    for n, want in ((0, 0), (19_999, 0), (20_000, 0), (20_001, 20_000), (40_000, 20_000), (40_001, 40_000), (100_000, 80_000)):
        assert L.airgpu_playback_samples(n, 20_000) == want
        sent, i = 0, 0
        while n >= 20_000 and i < n - 20_000:          # the reference loop (usize arithmetic: skipped when it would underflow)
            sent += 20_000
            i += 20_000
        assert sent == want


def test_set_timing_off_keeps_results():
    _, iq = capture_u8(seed=51, n=300_000, df17=3000.0)
    with AdsbDecoder(fmt=FMT_U8) as dec:
        dec.set_timing(False)
        got = dec.decode(iq)
        assert dec.stats()["kernel_ms"] == 0.0
        dec.set_timing(True)
        assert frames_equal(got, dec.decode(iq)) and dec.stats()["kernel_ms"] > 0.0
        assert frames_equal(got, oracle_c.decode_fast(iq)[0])

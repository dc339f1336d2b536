"""N3: the `.c16` loader / writer and the 20 000-sample playback chunking (src/utils.rs:7-43, src/adsb.rs:75-89).
CPU tests for the host logic; one GPU test replays a file through the decode thread on the C ABI."""
import queue

import numpy as np
import pytest

from air_rs_b200 import ingest

from common import capture_cs16


def test_c16_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(5)
    iq = rng.integers(-32768, 32767, size=2 * 12_345, dtype=np.int16)
    path = tmp_path / "cap.c16"
    ingest.save_data(iq, str(path))
    raw = path.read_bytes()
    assert len(raw) == 4 * 12_345
    # I then Q, little-endian i16 (utils.rs:12-15)
    assert int.from_bytes(raw[0:2], "little", signed=True) == int(iq[0])
    assert int.from_bytes(raw[2:4], "little", signed=True) == int(iq[1])
    back = ingest.load_data(str(path))
    assert back.dtype == np.int16 and np.array_equal(back, iq)


def test_load_rejects_truncated_file(tmp_path):
    path = tmp_path / "bad.c16"
    path.write_bytes(b"\x01\x02\x03\x04\x05\x06")
    with pytest.raises(ValueError, match="not divisible by 4"):      # utils.rs:28-30
        ingest.load_data(str(path))
    with pytest.raises(TypeError):
        ingest.save_data(np.zeros(4, dtype=np.uint8), str(tmp_path / "x.c16"))


@pytest.mark.parametrize("n,chunks", [(0, 0), (19_999, 0), (20_000, 0), (20_001, 1), (40_000, 1), (60_000, 2),
                                      (60_001, 3), (250_000, 12)])
def test_playback_chunking_drops_the_tail(n, chunks):
    """`while i < data.len() - 20000` (adsb.rs:77): the last chunk is dropped even when complete."""
    iq = np.arange(2 * n, dtype=np.int64).astype(np.int16)
    got = list(ingest.playback_chunks(iq))
    assert len(got) == chunks
    for k, buf in enumerate(got):
        assert buf.size == 40_000 and np.array_equal(buf, iq[40_000 * k: 40_000 * (k + 1)])
        assert buf.base is None or not np.shares_memory(buf, iq)      # owned, like `.to_vec()`
    tx = queue.Queue()
    assert ingest.playback_thread(tx, iq) == chunks
    items = [tx.get_nowait() for _ in range(chunks + 1)]
    assert not isinstance(items[-1], np.ndarray)                       # the channel is closed after the last chunk


@pytest.mark.gpu
def test_replay_file_through_decode_thread(tmp_path):
    from air_rs_b200.decoder import AdsbDecoder
    from air_rs_b200.native import FMT_CS16
    from oracle import oracle_c

    _, iq = capture_cs16(seed=88, n=310_000, df17=3000.0)
    path = tmp_path / "replay.c16"
    ingest.save_data(iq, str(path))
    with AdsbDecoder(fmt=FMT_CS16, max_buffer_samples=1 << 18, max_frames=1 << 14) as dec:
        pkts = ingest.launch_adsb_playback(str(path), decoder=dec)
    sent = 15 * 20_000                                                  # 310 000 samples -> 15 chunks, tail dropped
    want, _ = oracle_c.decode_fast(iq[: 2 * sent], 20_000, 0, threads=2)
    assert len(want) > 50
    assert [p.packet for p in pkts] == [bytes(r["bytes"]) for r in want]

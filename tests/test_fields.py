"""N1: frame field decode.  CPU: the oracle restatement of AdsbPacket::new against the
reference's unit tests (msgs.rs:225-321, aircraft.rs:184-199) and against the Python mirror.
GPU: fields_kernel through the C ABI, bit-exact vs the oracle."""
import numpy as np
import pytest

from air_rs_b200 import synth
from air_rs_b200.packet import AdsbPacket, AircraftID, AircraftPosition
from oracle import oracle_c


def _pkt(me_hex: str, head: str = "8d40621d") -> bytes:
    return bytes.fromhex(head + me_hex + "000000")


def test_oracle_fields_reference_kats():
    f = oracle_c.packet_fields(_pkt("202CC371C32CE0"))                    # msgs.rs:229-243
    assert f["kind"] == 1 and f["callsign"] == b"KLM1023_" and f["msg_type"] == 4
    for me, alt in (("58C382D690C8AC", 38000), ("58C282D690C8AC", 155000), ("580102D690C8AC", -1000),
                    ("580112D690C8AC", -975)):                            # msgs.rs:245-275
        assert oracle_c.packet_fields(_pkt(me))["altitude"] == alt
    even = oracle_c.packet_fields(_pkt("58C382D690C8AC"))                 # msgs.rs:277-321
    odd = oracle_c.packet_fields(_pkt("58c386435cc412"))
    for p in (even, odd):
        assert (p["kind"], p["msg_type"], p["surveillance_status"], p["nic_supplement"], p["cpr_time"]) == (2, 11, 0, 0, 0)
    assert (even["cpr_odd"], odd["cpr_odd"]) == (0, 1)
    assert (even["cpr_latitude"], even["cpr_longitude"]) == (93000, 51372)
    assert (odd["cpr_latitude"], odd["cpr_longitude"]) == (74158, 50194)
    g = oracle_c.packet_fields(bytes.fromhex("8d7c6b3020293532d70820fc8090"))   # aircraft.rs:184-191
    assert g["callsign"] == b"JST250__" and g["icao"] == 0x7C6B30 and g["downlink_format"] == 17
    assert g["capability"] == (0x8D & 5)                                  # packet.rs:27 masks with 5
    assert oracle_c.packet_fields(bytes.fromhex("8d7c6b30581304f388bb4455896f"))["altitude"] == 2600


def _random_frames(n, seed=3):
    rng = np.random.default_rng(seed)
    fr = np.zeros(n, dtype=oracle_c.FRAME_DTYPE)
    fr["bytes"] = rng.integers(0, 256, size=(n, 14), dtype=np.uint8)
    fr["bytes"][: 32 * 8, 4] = np.repeat(np.arange(32, dtype=np.uint8), 8) << 3 | rng.integers(0, 8, 256, dtype=np.uint8)
    fr["offset"] = np.arange(n)
    return fr


def test_oracle_fields_equal_python_mirror():
    fr = _random_frames(4000)
    got = oracle_c.frames_fields(fr)
    for r, f in zip(fr, got):
        p = AdsbPacket(bytes(r["bytes"]))
        assert (p.icao, p.downlink_format, p.capability, p.msg_type) == (f["icao"], f["downlink_format"], f["capability"], f["msg_type"])
        if isinstance(p.msg, AircraftID):
            assert f["kind"] == 1 and p.msg.callsign.encode() == f["callsign"]
        elif isinstance(p.msg, AircraftPosition):
            m = p.msg
            assert f["kind"] == 2
            assert (m.altitude, m.cpr_latitude, m.cpr_longitude, m.surveillance_status, m.nic_supplement, m.cpr_time,
                    int(m.cpr_odd)) == (f["altitude"], f["cpr_latitude"], f["cpr_longitude"], f["surveillance_status"],
                                        f["nic_supplement"], f["cpr_time"], f["cpr_odd"])
        else:
            assert f["kind"] == 0 and f["altitude"] == 0 and f["callsign"] == b""


@pytest.mark.gpu
def test_device_fields_match_oracle():
    from air_rs_b200.decoder import AdsbDecoder
    from air_rs_b200.native import FMT_U8

    with AdsbDecoder(fmt=FMT_U8) as dec:
        # every type code, random payloads
        fr = _random_frames(100_000, seed=5)
        got = dec.decode_fields(fr)
        want = oracle_c.frames_fields(fr)
        assert got.tobytes() == want.tobytes()
        # the frames of a decoded capture (incl. the reference's golden frames)
        tab = synth.make_traffic(9, 2_400_000, df17_per_s=3000, decoy_per_s=1000, snr_db=(10, 30))
        frames = dec.decode(synth.render(tab, 9, 0, 2_400_000, synth.FMT_U8, 2.0))
        got = dec.decode_fields(frames)
        assert got.tobytes() == oracle_c.frames_fields(frames).tobytes()
        assert (got["kind"] == 1).sum() > 20 and (got["kind"] == 2).sum() > 200
        assert b"JST250__" in set(got["callsign"].tolist())
        assert len(dec.decode_fields(frames[:0])) == 0


@pytest.mark.gpu
def test_device_fields_device_pointers():
    import torch

    from air_rs_b200.decoder import AdsbDecoder
    from air_rs_b200.native import FMT_U8

    fr = _random_frames(50_000, seed=6)
    with AdsbDecoder(fmt=FMT_U8) as dec:
        d_in = torch.from_numpy(fr.view(np.uint8).reshape(-1, 24)).cuda()
        d_out = torch.empty((fr.size, 32), dtype=torch.uint8, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            dec.decode_fields_device(d_in.data_ptr(), fr.size, d_out.data_ptr(), s.cuda_stream)
            s.synchronize()
        assert d_out.cpu().numpy().tobytes() == oracle_c.frames_fields(fr).tobytes()

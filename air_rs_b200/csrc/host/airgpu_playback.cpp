// airgpu_playback -- `air_rs adsb -m stream -p <file.c16>` with the decode thread on a B200.
//
// Mirrors launch_adsb (reference src/adsb.rs:126-173): a playback thread, the decode thread
// and a display thread joined by two unbounded channels.  Display is the reference's
// "stream" mode reduced to one line per packet so the output can be diffed against the oracle.
//
// Build: air_rs_b200/build.py::build_host() (g++ -O2 -std=c++17 -pthread, linked against libairgpu.so).
#include <cstdlib>
#include <cstring>
#include <thread>

#include "adsb_track.hpp"

using namespace adsb_host;

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <capture.c16> [chunk_samples=20000] [device=0] [stream|json]\n", argv[0]);
        return 2;
    }
    const size_t chunk = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 20000;
    const int device = argc > 3 ? std::atoi(argv[3]) : 0;
    const bool json_mode = argc > 4 && std::strcmp(argv[4], "json") == 0;   // what the web thread broadcasts

    airgpu_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.device = device;
    cfg.format = AIRGPU_FMT_CS16;                 // what the reference's channel carries
    cfg.ring_slots = 4;
    cfg.max_buffer_samples = chunk;
    cfg.max_frames = 8192;
    airgpu_ctx *ctx = nullptr;
    if (airgpu_create(&cfg, &ctx) != AIRGPU_OK) {  // the reference expect()s here (adsb.rs:36-44)
        std::fprintf(stderr, "Couldn't create decode stage: %s\n", airgpu_last_error());
        return 1;
    }

    IqBuffer data;
    try {
        data = load_data(argv[1]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::printf("Loaded %zu samples from playback file\n", data.size() / 2);

    Channel<IqBuffer> raw;
    Channel<AdsbPacket> msgs;
    uint64_t sent = 0;
    std::thread stream_thread([&] { playback_thread(raw, data, chunk); });
    std::thread process_thread([&] { sent = process_sdr_data_thread(raw, msgs, ctx); });
    std::thread display_thread([&] {
        AircraftMap aircrafts;
        double now = 0.0;     // replay clock: one tick per packet, so the 10 s pairing window is deterministic
        while (auto p = msgs.recv()) {
            if (json_mode) {  // web.rs:117-126: handle_aircraft_update -> get_summary -> serde_json::to_string
                now += 0.001;
                const Aircraft a = handle_aircraft_update(*p, aircrafts, now);
                std::printf("Broadcasting aircraft summary: %s\n", summary_json(a).c_str());
                continue;
            }
            std::printf("== %s == DF %u CA %u ICAO %06X TC %u", p->hex().c_str(), p->downlink_format, p->capability,
                        p->icao, p->msg_type);
            if (p->kind == AdsbPacket::Kind::AircraftID) std::printf(" callsign %s", p->callsign.c_str());
            if (p->kind == AdsbPacket::Kind::AircraftPosition)
                std::printf(" alt %d cpr %s %u %u", p->altitude, p->cpr_odd ? "odd" : "even", p->cpr_latitude,
                            p->cpr_longitude);
            std::printf("\n");
        }
    });
    stream_thread.join();
    process_thread.join();
    display_thread.join();
    std::printf("packets: %llu\n", (unsigned long long)sent);
    airgpu_destroy(ctx);
    return 0;
}

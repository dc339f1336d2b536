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

#include "adsb_host.hpp"

using namespace adsb_host;

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <capture.c16> [chunk_samples=20000] [device=0]\n", argv[0]);
        return 2;
    }
    const size_t chunk = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 20000;
    const int device = argc > 3 ? std::atoi(argv[3]) : 0;

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
        while (auto p = msgs.recv()) {
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

// adsb_host.hpp -- C++ mirror of the host side of the reference's ADS-B pipeline,
// everything that sits around the decode thread (jaxsonpd/air_rs):
//   Channel<T>          std::sync::mpsc::channel (unbounded)         src/adsb.rs:131,146
//   load_data           .c16 loader (LE i16 I,Q)                      src/utils.rs:23-43
//   playback_thread     20 000-sample chunks, tail dropped            src/adsb.rs:75-89
//   AdsbPacket          AdsbPacket::new and its Display               src/adsb/packet.rs:10-102
//   process_sdr_data_thread   the decode thread, on the airgpu C ABI  src/adsb.rs:92-122
// The reference is Rust; no Rust toolchain exists in this image, so this C++ harness is
// the runnable stand-in.  INTEGRATION.md shows the equivalent Rust FFI.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <fstream>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/airgpu.h"

namespace adsb_host {

// Interleaved I,Q samples of one buffer: what Vec<Complex<i16>>::as_ptr() points at.
using IqBuffer = std::vector<int16_t>;

template <typename T>
class Channel {
public:
    // Sender::send; returns false when the receiver has been dropped (adsb.rs:65-68)
    bool send(T v)
    {
        std::lock_guard<std::mutex> lk(m_);
        if (rx_dropped_) return false;
        q_.push_back(std::move(v));
        cv_.notify_one();
        return true;
    }
    // Receiver::recv; nullopt once every sender is dropped and the queue is empty (adsb.rs:95)
    std::optional<T> recv()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty() || tx_dropped_; });
        if (q_.empty()) return std::nullopt;
        T v = std::move(q_.front());
        q_.pop_front();
        return v;
    }
    void drop_sender()
    {
        std::lock_guard<std::mutex> lk(m_);
        tx_dropped_ = true;
        cv_.notify_all();
    }
    void drop_receiver()
    {
        std::lock_guard<std::mutex> lk(m_);
        rx_dropped_ = true;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool tx_dropped_ = false, rx_dropped_ = false;
};

// utils.rs:23-43: whole file, little-endian i16 I then Q; length must be a multiple of 4.
inline IqBuffer load_data(const std::string &path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("Couldn't load playback data file: " + path);
    const std::streamsize n = f.tellg();
    if (n % 4 != 0) throw std::runtime_error("Invalid file length (not divisible by 4)");
    IqBuffer data(static_cast<size_t>(n / 2));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(data.data()), n);   // x86-64 / aarch64 hosts are little-endian
    return data;
}

// adsb.rs:75-89: `while i < data.len() - 20000` -- the final (possibly full) chunk is dropped.
inline void playback_thread(Channel<IqBuffer> &tx, const IqBuffer &data, size_t chunk_samples = 20000)
{
    const size_t n = data.size() / 2;
    size_t i = 0;
    while (n >= chunk_samples && i < n - chunk_samples) {
        IqBuffer buf(data.begin() + 2 * i, data.begin() + 2 * (i + chunk_samples));
        i += chunk_samples;
        if (!tx.send(std::move(buf))) {
            std::puts("Raw sdr receiver is dropped");
            return;
        }
        // the reference sleeps 5 ms per chunk here to imitate a 4 MS/s device (adsb.rs:84);
        // the harness replays as fast as the decode stage accepts
    }
    tx.drop_sender();
}

struct AdsbPacket {   // packet.rs:10-18
    std::vector<uint8_t> packet;
    uint8_t downlink_format = 0, capability = 0, msg_type = 0;
    uint32_t icao = 0;
    enum class Kind { AircraftID, AircraftPosition, Unknown } kind = Kind::Unknown;
    std::string callsign;                      // AircraftID   (msgs.rs:171-201)
    int32_t altitude = 0;                      // AircraftPosition (msgs.rs:69-102)
    uint8_t surveillance_status = 0, nic_supplement = 0, cpr_time = 0;
    bool cpr_odd = false;
    uint32_t cpr_latitude = 0, cpr_longitude = 0;

    explicit AdsbPacket(std::vector<uint8_t> p) : packet(std::move(p))   // packet.rs:25-49
    {
        static const char *kChars = "#ABCDEFGHIJKLMNOPQRSTUVWXYZ#####_###############0123456789######";
        downlink_format = packet[0] >> 3;
        capability = packet[0] & 5;            // sic (packet.rs:27)
        icao = (uint32_t(packet[1]) << 16) | (uint32_t(packet[2]) << 8) | packet[3];
        msg_type = packet[4] >> 3;
        const uint8_t *me = &packet[4];
        if (msg_type >= 1 && msg_type <= 4) {
            kind = Kind::AircraftID;
            uint64_t acc = 0;
            for (int k = 1; k < 7; ++k) acc = (acc << 8) | me[k];
            for (int k = 0; k < 8; ++k) callsign.push_back(kChars[(acc >> (42 - 6 * k)) & 0x3F]);
        } else if (msg_type >= 9 && msg_type <= 18) {
            kind = Kind::AircraftPosition;
            const bool alt25 = (me[1] & 1) == 1;
            altitude = (int32_t((me[1] & 0xFE) >> 1) << 4) | ((me[2] & 0xF0) >> 4);
            altitude = altitude * (alt25 ? 25 : 100) - 1000;
            surveillance_status = (me[0] & 6) >> 1;
            nic_supplement = me[0] & 1;
            cpr_time = (me[2] & 8) >> 3;
            cpr_odd = ((me[2] & 4) >> 2) == 1;
            cpr_latitude = (uint32_t(me[2] & 3) << 15) | (uint32_t(me[3]) << 7) | ((me[4] & 0xFE) >> 1);
            cpr_longitude = (uint32_t(me[4] & 1) << 16) | (uint32_t(me[5]) << 8) | me[6];
        }
    }

    std::string hex() const
    {
        static const char *d = "0123456789abcdef";
        std::string s;
        for (uint8_t b : packet) {
            s.push_back(d[b >> 4]);
            s.push_back(d[b & 15]);
        }
        return s;
    }
};

// The decode thread on the GPU stage.  Same contract as adsb.rs:92-122: one buffer per
// message, frames sent in ascending offset order, buffers in arrival order, tx dropped on
// return.  `depth` buffers are in flight in the pinned ring while the next one is received.
inline uint64_t process_sdr_data_thread(Channel<IqBuffer> &rx, Channel<AdsbPacket> &tx, airgpu_ctx *ctx,
                                        size_t depth = 2, size_t max_frames = 8192)
{
    std::deque<uint64_t> pending;
    std::vector<airgpu_frame> out(max_frames);
    uint64_t sent = 0, base = 0;
    bool alive = true;
    auto drain = [&](size_t keep) {
        while (alive && pending.size() > keep) {
            size_t n = 0;
            int rc = airgpu_collect(ctx, pending.front(), out.data(), out.size(), &n);
            if (rc == AIRGPU_ERR_OVERFLOW && n > out.size()) {
                // more frames than the array holds (a constant buffer yields one at EVERY offset): the ticket is still
                // collectable -- the reference sends every packet (adsb.rs:98-111), so nothing is dropped here either
                out.resize(n);
                rc = airgpu_collect(ctx, pending.front(), out.data(), out.size(), &n);
            }
            pending.pop_front();
            if (rc != AIRGPU_OK) throw std::runtime_error(std::string("airgpu_collect: ") + airgpu_last_error());
            for (size_t k = 0; k < n; ++k) {
                AdsbPacket pkt(std::vector<uint8_t>(out[k].bytes, out[k].bytes + 14));   // adsb.rs:107
                if (!tx.send(std::move(pkt))) {
                    std::puts("Adsb msg receiver is dropped");                          // adsb.rs:108-111
                    alive = false;
                    return;
                }
                ++sent;
            }
        }
    };
    while (alive) {
        std::optional<IqBuffer> buf = rx.recv();                                       // adsb.rs:95
        if (!buf) break;
        uint64_t ticket = 0;
        const size_t n = buf->size() / 2;
        if (airgpu_submit(ctx, buf->data(), n, base, &ticket) != AIRGPU_OK)
            throw std::runtime_error(std::string("airgpu_submit: ") + airgpu_last_error());
        base += n;
        pending.push_back(ticket);
        drain(depth - 1);
    }
    drain(0);
    tx.drop_sender();                                                                   // adsb.rs:121
    return sent;
}

}  // namespace adsb_host

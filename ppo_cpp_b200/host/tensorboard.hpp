// TensorboardWriter with the reference's interface (ppo2/tensorboard.hpp:13-52) and no TensorFlow: the event file is
// written directly.  Format (TensorFlow's EventsWriter / RecordWriter, which TensorBoard reads):
//   file   <prefix>.out.tfevents.<unix time>.<hostname>, first record = Event{wall_time, file_version "brain.Event:2"}
//   record uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)      (little endian)
//   data   Event protobuf: 1 wall_time (double), 2 step (int64), 3 file_version (string),
//          5 summary { 1 value { 1 tag (string), 2 simple_value (float) } }
//   masked crc = rotr15(crc32c) + 0xa282ead8, crc32c = CRC-32/Castagnoli (polynomial 0x1EDC6F41, reflected 0x82F63B78)
#ifndef PPO_CPP_TENSORBOARD_HPP
#define PPO_CPP_TENSORBOARD_HPP

#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <string>

class TensorboardWriter {
public:
    TensorboardWriter(const std::string& tensorboard_log_path, const std::string& tb_log_name, bool new_tb_log = true)
        : save_path{tensorboard_log_path + tb_log_name} {
        (void)new_tb_log;
        std::cout << "tb" << std::endl;
        std::cout << save_path << std::endl;
        char host[256] = "localhost";
        gethostname(host, sizeof(host) - 1);
        const double now = static_cast<double>(time(nullptr));
        file_name = save_path + ".out.tfevents." + std::to_string(static_cast<long long>(now)) + "." + host;
        out.open(file_name, std::ios::binary | std::ios::trunc);
        if (!out.is_open()) {
            std::cout << "Unable to open the event file " << file_name << std::endl;
            return;
        }
        std::string ev;
        put_double(ev, 1, now);
        put_bytes(ev, 3, "brain.Event:2");
        write_record(ev);
    }

    void write_scalar(double wall_time, int64_t step, const std::string& tag, float simple_value) {
        std::string value;
        put_bytes(value, 1, tag);
        put_float(value, 2, simple_value);
        std::string summary;
        put_bytes(summary, 1, value);
        std::string ev;
        put_double(ev, 1, wall_time);
        put_varint_field(ev, 2, static_cast<uint64_t>(step));
        put_bytes(ev, 5, summary);
        write_record(ev);
    }

    void write_scalar(int64_t step, const std::string& tag, float simple_value) {
        write_scalar(static_cast<double>(time(nullptr)), step, tag, simple_value);
    }

    // an already serialised Summary message (the reference feeds the graph's merged summaries through this)
    void write_summary(int64_t step, const std::string& encoded_summary) {
        std::string ev;
        put_double(ev, 1, static_cast<double>(time(nullptr)));
        put_varint_field(ev, 2, static_cast<uint64_t>(step));
        put_bytes(ev, 5, encoded_summary);
        write_record(ev);
    }

    void flush() { out.flush(); }
    const std::string& path() const { return file_name; }

    static uint32_t crc32c(const void* data, size_t n) {
        static uint32_t table[256];
        static bool ready = false;
        if (!ready) {
            for (uint32_t i = 0; i < 256; ++i) {
                uint32_t c = i;
                for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
                table[i] = c;
            }
            ready = true;
        }
        uint32_t c = 0xFFFFFFFFu;
        const unsigned char* p = static_cast<const unsigned char*>(data);
        for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    }
    static uint32_t masked_crc32c(const void* data, size_t n) {
        const uint32_t c = crc32c(data, n);
        return ((c >> 15) | (c << 17)) + 0xa282ead8u;
    }

private:
    static void put_varint(std::string& s, uint64_t v) {
        while (v >= 0x80u) {
            s.push_back(static_cast<char>((v & 0x7Fu) | 0x80u));
            v >>= 7;
        }
        s.push_back(static_cast<char>(v));
    }
    static void put_varint_field(std::string& s, int field, uint64_t v) {
        put_varint(s, static_cast<uint64_t>(field) << 3);  // wire type 0
        put_varint(s, v);
    }
    static void put_double(std::string& s, int field, double v) {
        put_varint(s, (static_cast<uint64_t>(field) << 3) | 1u);  // wire type 1: 64-bit
        char b[8];
        std::memcpy(b, &v, 8);
        s.append(b, 8);
    }
    static void put_float(std::string& s, int field, float v) {
        put_varint(s, (static_cast<uint64_t>(field) << 3) | 5u);  // wire type 5: 32-bit
        char b[4];
        std::memcpy(b, &v, 4);
        s.append(b, 4);
    }
    static void put_bytes(std::string& s, int field, const std::string& v) {
        put_varint(s, (static_cast<uint64_t>(field) << 3) | 2u);  // wire type 2: length-delimited
        put_varint(s, v.size());
        s.append(v);
    }
    void write_record(const std::string& data) {
        if (!out.is_open()) return;
        const uint64_t len = data.size();
        char hdr[12];
        std::memcpy(hdr, &len, 8);
        const uint32_t lc = masked_crc32c(hdr, 8);
        std::memcpy(hdr + 8, &lc, 4);
        out.write(hdr, 12);
        out.write(data.data(), static_cast<std::streamsize>(data.size()));
        const uint32_t dc = masked_crc32c(data.data(), data.size());
        out.write(reinterpret_cast<const char*>(&dc), 4);
        out.flush();
    }

    std::string save_path;
    std::string file_name;
    std::ofstream out;
};

#endif  // PPO_CPP_TENSORBOARD_HPP

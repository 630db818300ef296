// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for Xilinx's hls_stream.h so the
// reference designs under /root/reference compile with plain g++ (C-simulation semantics).
// Written from the documented hls::stream<T> behaviour (unbounded FIFO in csim); no Xilinx code.
#ifndef ORACLE_SHIM_HLS_STREAM_H
#define ORACLE_SHIM_HLS_STREAM_H
#include <cstdio>
#include <cstdlib>
#include <deque>
namespace hls {
template <typename T>
class stream {
    std::deque<T> q_;
public:
    stream() {}
    explicit stream(const char*) {}
    stream(const stream&) = delete;
    stream& operator=(const stream&) = delete;
    void write(const T& v) { q_.push_back(v); }
    T read() {
        if (q_.empty()) {
            // Vivado csim prints a warning and returns a default value; a drained FIFO here means
            // the sequential csim schedule is broken, so fail loudly instead of inventing data.
            std::fprintf(stderr, "hls::stream shim: read from empty FIFO\n");
            std::abort();
        }
        T v = q_.front();
        q_.pop_front();
        return v;
    }
    void read(T& v) { v = read(); }
    bool empty() const { return q_.empty(); }
    bool full() const { return false; }
    size_t size() const { return q_.size(); }
    stream& operator<<(const T& v) { write(v); return *this; }
    stream& operator>>(T& v) { v = read(); return *this; }
};
}  // namespace hls
#endif

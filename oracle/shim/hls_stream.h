// oracle/shim/hls_stream.h -- TEST INFRASTRUCTURE ONLY (oracle build).
//
// Stand-in for the Vivado-HLS "hls_stream.h" the reference includes
// (conv_nonsquare_top.cpp:43, bnn-library.h:47) but does not ship: an unbounded
// FIFO with read()/write()/empty()/size() and the (const char*) constructor,
// which is everything the C-simulation of the reference uses.
#ifndef FCB_ORACLE_SHIM_HLS_STREAM_H
#define FCB_ORACLE_SHIM_HLS_STREAM_H

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace hls {

template <typename T> class stream {
  std::vector<T> m_buf;
  size_t m_head;
  const char* m_name;

 public:
  stream() : m_head(0), m_name("stream") {}
  explicit stream(const char* name) : m_head(0), m_name(name) {}
  stream(const stream&) = delete;
  stream& operator=(const stream&) = delete;

  bool empty() const { return m_head == m_buf.size(); }
  size_t size() const { return m_buf.size() - m_head; }

  void write(const T& v) { m_buf.push_back(v); }
  T read() {
    if (m_head == m_buf.size()) {
      // C-sim would warn and return garbage; an oracle must not continue.
      std::fprintf(stderr, "hls::stream '%s': read on empty stream\n", m_name);
      std::abort();
    }
    T v = m_buf[m_head++];
    if (m_head == m_buf.size()) {
      m_buf.clear();
      m_head = 0;
    } else if (m_head >= (1u << 16) && m_head * 2 >= m_buf.size()) {
      m_buf.erase(m_buf.begin(), m_buf.begin() + m_head);
      m_head = 0;
    }
    return v;
  }
  void read(T& v) { v = read(); }
  bool read_nb(T& v) {
    if (empty()) return false;
    v = read();
    return true;
  }
  void operator>>(T& v) { v = read(); }
  void operator<<(const T& v) { write(v); }
  void reserve(size_t n) { m_buf.reserve(n); }
};

}  // namespace hls

#endif

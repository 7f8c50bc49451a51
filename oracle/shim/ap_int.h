// oracle/shim/ap_int.h -- TEST INFRASTRUCTURE ONLY (oracle build).
//
// A from-scratch stand-in for the Xilinx Vivado-HLS 2020.1 "ap_int.h" header,
// which the reference sources include but do not ship (reference README:5
// points at E:\Xilinx\Vivado\2020.1\include).  It provides exactly the subset
// of arbitrary-precision integer behaviour the reference relies on
// (SURVEY.md Appendix B):
//   * ap_uint<W> / ap_int<W>, 1 <= W <= AP_INT_MAX_W, `static const int width`
//   * two's-complement wrap on every store (AP_WRAP), sign-/zero-extension on
//     widening, implicit construction from built-in integers (needed by the
//     aggregate initialisers of memdata_nonsquare.h:4-15)
//   * range select x(hi,lo) as l-value and r-value (also on const objects,
//     interpret.hpp:209-212), bit select x[i]
//   * the operators used by mac.hpp:169, streamtools.h:485-515,
//     conv_nonsquare_top.cpp:272-275, maxpool.h:81-170, activations.hpp:57-99
//   * ap_fixed / ap_q_mode / ap_o_mode declared (interpret.hpp:183-189; never
//     instantiated)
// Nothing here is derived from Xilinx code; semantics were restated from the
// call sites above.  W <= 64 values live in the smallest native container so
// that sizeof(ap_int<8>) == 1 (keeps the testbench's stack arrays small).
#ifndef FCB_ORACLE_SHIM_AP_INT_H
#define FCB_ORACLE_SHIM_AP_INT_H

#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <type_traits>

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };
template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP, int N = 0> struct ap_fixed;

template <int W, bool S, bool Wide = (W > 64)> struct ap_int_base;
template <int W> struct ap_int;
template <int W> struct ap_uint;
template <int W, bool S> struct ap_range_ref;
template <int W, bool S> struct ap_bit_ref;

namespace apshim {

template <int W, bool S> struct store {
  typedef typename std::conditional<
      (W <= 8), typename std::conditional<S, int8_t, uint8_t>::type,
      typename std::conditional<
          (W <= 16), typename std::conditional<S, int16_t, uint16_t>::type,
          typename std::conditional<
              (W <= 32), typename std::conditional<S, int32_t, uint32_t>::type,
              typename std::conditional<S, int64_t, uint64_t>::type>::type>::type>::type type;
};

// wrap a 64-bit pattern to W bits, then sign- or zero-extend back to 64 bits
template <int W, bool S> constexpr uint64_t wrap64(uint64_t x) {
  return (W >= 64) ? x
         : S      ? (uint64_t)((int64_t)(x << (64 - (W >= 64 ? 0 : W))) >> (64 - (W >= 64 ? 0 : W)))
                  : (x & ((~0ull) >> (64 - (W >= 64 ? 0 : W))));
}

inline uint64_t mask_n(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

}  // namespace apshim

// ---------------------------------------------------------------------------
// W <= 64
// ---------------------------------------------------------------------------
template <int W, bool S> struct ap_int_base<W, S, false> {
  static_assert(W >= 1, "width");
  static const int width = W;
  typedef typename apshim::store<W, S>::type store_t;
  typedef typename std::conditional<(S || W < 64), long long, unsigned long long>::type conv_t;
  store_t V;

  static constexpr store_t norm(uint64_t x) { return (store_t)apshim::wrap64<W, S>(x); }

  constexpr ap_int_base() : V(0) {}
  constexpr ap_int_base(bool v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(char v) : V(norm((uint64_t)(long long)v)) {}
  constexpr ap_int_base(signed char v) : V(norm((uint64_t)(long long)v)) {}
  constexpr ap_int_base(unsigned char v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(short v) : V(norm((uint64_t)(long long)v)) {}
  constexpr ap_int_base(unsigned short v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(int v) : V(norm((uint64_t)(long long)v)) {}
  constexpr ap_int_base(unsigned v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(long v) : V(norm((uint64_t)(long long)v)) {}
  constexpr ap_int_base(unsigned long v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(long long v) : V(norm((uint64_t)v)) {}
  constexpr ap_int_base(unsigned long long v) : V(norm((uint64_t)v)) {}

  template <int W2, bool S2>
  ap_int_base(const ap_int_base<W2, S2, false>& o) : V(norm(o.bits64())) {}
  template <int W2, bool S2>
  ap_int_base(const ap_int_base<W2, S2, true>& o) : V(norm(o.get_bits(0, 64))) {}
  template <int W2, bool S2> ap_int_base(const ap_range_ref<W2, S2>& r) : V(norm(r.get64())) {}
  template <int W2, bool S2> ap_int_base(const ap_bit_ref<W2, S2>& r) : V(norm((uint64_t)(bool)r)) {}

  // value as a sign-/zero-extended 64-bit pattern
  constexpr uint64_t bits64() const { return (uint64_t)(conv_t)V; }
  constexpr operator conv_t() const { return (conv_t)V; }

  // raw bit access (n in 1..64, lo+n <= W is the caller's business)
  uint64_t get_bits(int lo, int n) const {
    uint64_t u = (uint64_t)V & apshim::mask_n(W);
    return (u >> lo) & apshim::mask_n(n);
  }
  void set_bits(int lo, int n, uint64_t v) {
    uint64_t m = apshim::mask_n(n) << lo;
    uint64_t u = ((uint64_t)V & ~m) | ((v << lo) & m);
    V = norm(u);
  }

  ap_range_ref<W, S> operator()(int hi, int lo) { return ap_range_ref<W, S>(this, hi, lo); }
  ap_range_ref<W, S> operator()(int hi, int lo) const {
    return ap_range_ref<W, S>(const_cast<ap_int_base*>(this), hi, lo);
  }
  ap_range_ref<W, S> range(int hi, int lo) { return ap_range_ref<W, S>(this, hi, lo); }
  ap_range_ref<W, S> range(int hi, int lo) const {
    return ap_range_ref<W, S>(const_cast<ap_int_base*>(this), hi, lo);
  }
  ap_bit_ref<W, S> operator[](int i) { return ap_bit_ref<W, S>(this, i); }
  bool operator[](int i) const { return (((uint64_t)V) >> i) & 1u; }

#define FCB_SHIM_CASSIGN(OP)                                  \
  ap_int_base& operator OP##=(long long o) {                  \
    V = norm((uint64_t)((long long)(conv_t)V OP o));          \
    return *this;                                             \
  }
  FCB_SHIM_CASSIGN(+)
  FCB_SHIM_CASSIGN(-)
  FCB_SHIM_CASSIGN(*)
  FCB_SHIM_CASSIGN(|)
  FCB_SHIM_CASSIGN(&)
  FCB_SHIM_CASSIGN(^)
#undef FCB_SHIM_CASSIGN
  ap_int_base& operator>>=(int s) {
    V = norm((uint64_t)(s >= 64 ? (conv_t)(((conv_t)V) < 0 ? -1 : 0) : (conv_t)(((conv_t)V) >> s)));
    return *this;
  }
  ap_int_base& operator<<=(int s) {
    V = norm(s >= 64 ? 0ull : (((uint64_t)(conv_t)V) << s));
    return *this;
  }
  ap_int_base& operator++() { V = norm((uint64_t)((conv_t)V + 1)); return *this; }
  ap_int_base& operator--() { V = norm((uint64_t)((conv_t)V - 1)); return *this; }

  long long to_int64() const { return (long long)(conv_t)V; }
  unsigned long long to_uint64() const { return (unsigned long long)(conv_t)V; }
  int to_int() const { return (int)(conv_t)V; }
  unsigned to_uint() const { return (unsigned)(conv_t)V; }
  int length() const { return W; }
};

// ---------------------------------------------------------------------------
// W > 64 : little-endian array of 64-bit limbs, bits >= W kept zero
// ---------------------------------------------------------------------------
template <int W, bool S> struct ap_int_base<W, S, true> {
  static const int width = W;
  static const int NL = (W + 63) / 64;
  uint64_t L[NL];

  void trim() {
    if (W % 64) L[NL - 1] &= apshim::mask_n(W % 64);
  }
  void fill_from64(uint64_t lo, bool neg) {
    L[0] = lo;
    for (int i = 1; i < NL; i++) L[i] = neg ? ~0ull : 0ull;
    trim();
  }

  ap_int_base() { for (int i = 0; i < NL; i++) L[i] = 0; }
  ap_int_base(bool v) { fill_from64((uint64_t)v, false); }
  ap_int_base(int v) { fill_from64((uint64_t)(long long)v, v < 0); }
  ap_int_base(unsigned v) { fill_from64((uint64_t)v, false); }
  ap_int_base(long v) { fill_from64((uint64_t)(long long)v, v < 0); }
  ap_int_base(unsigned long v) { fill_from64((uint64_t)v, false); }
  ap_int_base(long long v) { fill_from64((uint64_t)v, v < 0); }
  ap_int_base(unsigned long long v) { fill_from64((uint64_t)v, false); }

  template <int W2, bool S2> ap_int_base(const ap_int_base<W2, S2, false>& o) {
    fill_from64(o.bits64(), S2 && ((long long)o.bits64() < 0));
  }
  template <int W2, bool S2> ap_int_base(const ap_int_base<W2, S2, true>& o) {
    const int n2 = ap_int_base<W2, S2, true>::NL;
    bool neg = S2 && ((o.L[n2 - 1] >> ((W2 - 1) % 64)) & 1u);
    for (int i = 0; i < NL; i++) {
      uint64_t v = (i < n2) ? o.L[i] : (neg ? ~0ull : 0ull);
      if (neg && i == n2 - 1 && (W2 % 64)) v |= ~apshim::mask_n(W2 % 64);
      L[i] = v;
    }
    trim();
  }
  template <int W2, bool S2> ap_int_base(const ap_range_ref<W2, S2>& r) {
    for (int i = 0; i < NL; i++) L[i] = 0;
    int n = r.hi - r.lo + 1;
    if (n > W) n = W;
    for (int b = 0; b < n; b += 64) {
      int c = (n - b) < 64 ? (n - b) : 64;
      set_bits(b, c, r.p->get_bits(r.lo + b, c));
    }
  }

  uint64_t get_bits(int lo, int n) const {
    int li = lo >> 6, sh = lo & 63;
    if (li >= NL) return 0;
    uint64_t v = L[li] >> sh;
    if (sh && (li + 1) < NL && (sh + n) > 64) v |= L[li + 1] << (64 - sh);
    return v & apshim::mask_n(n);
  }
  void set_bits(int lo, int n, uint64_t v) {
    int li = lo >> 6, sh = lo & 63;
    if (li >= NL) return;
    v &= apshim::mask_n(n);
    uint64_t m0 = apshim::mask_n(n) << sh;
    L[li] = (L[li] & ~m0) | (v << sh);
    if (sh && (sh + n) > 64 && (li + 1) < NL) {
      int n1 = sh + n - 64;
      uint64_t m1 = apshim::mask_n(n1);
      L[li + 1] = (L[li + 1] & ~m1) | ((v >> (64 - sh)) & m1);
    }
    trim();
  }

  ap_range_ref<W, S> operator()(int hi, int lo) { return ap_range_ref<W, S>(this, hi, lo); }
  ap_range_ref<W, S> operator()(int hi, int lo) const {
    return ap_range_ref<W, S>(const_cast<ap_int_base*>(this), hi, lo);
  }
  ap_range_ref<W, S> range(int hi, int lo) { return ap_range_ref<W, S>(this, hi, lo); }
  ap_range_ref<W, S> range(int hi, int lo) const {
    return ap_range_ref<W, S>(const_cast<ap_int_base*>(this), hi, lo);
  }
  ap_bit_ref<W, S> operator[](int i) { return ap_bit_ref<W, S>(this, i); }
  bool operator[](int i) const { return (L[i >> 6] >> (i & 63)) & 1u; }

  ap_int_base& operator>>=(int s) {
    if (s <= 0) return *this;
    int ls = s >> 6, bs = s & 63;
    for (int i = 0; i < NL; i++) {
      uint64_t lo = (i + ls) < NL ? L[i + ls] : 0ull;
      uint64_t hi = (i + ls + 1) < NL ? L[i + ls + 1] : 0ull;
      L[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
    return *this;
  }
  ap_int_base& operator<<=(int s) {
    if (s <= 0) return *this;
    int ls = s >> 6, bs = s & 63;
    for (int i = NL - 1; i >= 0; i--) {
      uint64_t hi = (i - ls) >= 0 ? L[i - ls] : 0ull;
      uint64_t lo = (i - ls - 1) >= 0 ? L[i - ls - 1] : 0ull;
      L[i] = bs ? ((hi << bs) | (lo >> (64 - bs))) : hi;
    }
    trim();
    return *this;
  }
  ap_int_base& operator|=(const ap_int_base& o) { for (int i = 0; i < NL; i++) L[i] |= o.L[i]; return *this; }
  ap_int_base& operator&=(const ap_int_base& o) { for (int i = 0; i < NL; i++) L[i] &= o.L[i]; return *this; }
  ap_int_base& operator^=(const ap_int_base& o) { for (int i = 0; i < NL; i++) L[i] ^= o.L[i]; return *this; }
  bool operator==(const ap_int_base& o) const {
    for (int i = 0; i < NL; i++) if (L[i] != o.L[i]) return false;
    return true;
  }
  bool operator!=(const ap_int_base& o) const { return !(*this == o); }
  explicit operator bool() const {
    for (int i = 0; i < NL; i++) if (L[i]) return true;
    return false;
  }
  unsigned long long to_uint64() const { return L[0]; }
  long long to_int64() const { return (long long)L[0]; }
  int length() const { return W; }
  std::string to_hex() const {
    static const char* d = "0123456789abcdef";
    std::string s;
    bool started = false;
    for (int nib = (W + 3) / 4 - 1; nib >= 0; nib--) {
      unsigned v = (unsigned)get_bits(nib * 4, 4);
      if (v || started || nib == 0) { s.push_back(d[v]); started = true; }
    }
    return s;
  }
};

// ---------------------------------------------------------------------------
// user-facing types
// ---------------------------------------------------------------------------
template <int W> struct ap_uint : ap_int_base<W, false> {
  typedef ap_int_base<W, false> base;
  using base::base;
  constexpr ap_uint() : base() {}
  ap_uint(const base& b) : base(b) {}
};
template <int W> struct ap_int : ap_int_base<W, true> {
  typedef ap_int_base<W, true> base;
  using base::base;
  constexpr ap_int() : base() {}
  ap_int(const base& b) : base(b) {}
};

// wide-only free operators (narrow types go through their built-in conversion)
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator>>(const ap_uint<W>& a, int s) { ap_uint<W> r(a); r >>= s; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator>>(const ap_uint<W>& a, unsigned s) { ap_uint<W> r(a); r >>= (int)s; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator<<(const ap_uint<W>& a, int s) { ap_uint<W> r(a); r <<= s; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator<<(const ap_uint<W>& a, unsigned s) { ap_uint<W> r(a); r <<= (int)s; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator|(const ap_uint<W>& a, const ap_uint<W>& b) { ap_uint<W> r(a); r |= b; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator&(const ap_uint<W>& a, const ap_uint<W>& b) { ap_uint<W> r(a); r &= b; return r; }
template <int W> typename std::enable_if<(W > 64), ap_uint<W> >::type operator^(const ap_uint<W>& a, const ap_uint<W>& b) { ap_uint<W> r(a); r ^= b; return r; }
template <int W, bool S>
typename std::enable_if<(W > 64), std::ostream&>::type operator<<(std::ostream& os, const ap_int_base<W, S, true>& v) {
  return os << "0x" << v.to_hex();
}

// ---------------------------------------------------------------------------
// range / bit references (unsigned views, as in the vendor library)
// ---------------------------------------------------------------------------
template <int W, bool S> struct ap_range_ref {
  ap_int_base<W, S>* p;
  int hi, lo;
  ap_range_ref(ap_int_base<W, S>* p_, int hi_, int lo_) : p(p_), hi(hi_), lo(lo_) {}
  ap_range_ref(const ap_range_ref& o) : p(o.p), hi(o.hi), lo(o.lo) {}
  int length() const { return hi - lo + 1; }

  uint64_t get64() const {
    int n = hi - lo + 1;
    return p->get_bits(lo, n > 64 ? 64 : n);
  }
  operator unsigned long long() const { return get64(); }

  void assign64(uint64_t v) {
    int n = hi - lo + 1;
    p->set_bits(lo, n > 64 ? 64 : n, v);
    for (int b = 64; b < n; b += 64) p->set_bits(lo + b, (n - b) < 64 ? (n - b) : 64, 0);
  }
  // sign-extending fill for signed sources narrower than the range
  void assign64s(uint64_t v, bool neg) {
    int n = hi - lo + 1;
    p->set_bits(lo, n > 64 ? 64 : n, v);
    for (int b = 64; b < n; b += 64) p->set_bits(lo + b, (n - b) < 64 ? (n - b) : 64, neg ? ~0ull : 0ull);
  }
  ap_range_ref& operator=(unsigned long long v) { assign64(v); return *this; }
  ap_range_ref& operator=(long long v) { assign64s((uint64_t)v, v < 0); return *this; }
  ap_range_ref& operator=(int v) { assign64s((uint64_t)(long long)v, v < 0); return *this; }
  ap_range_ref& operator=(unsigned v) { assign64(v); return *this; }
  ap_range_ref& operator=(long v) { assign64s((uint64_t)(long long)v, v < 0); return *this; }
  ap_range_ref& operator=(unsigned long v) { assign64(v); return *this; }
  template <int W2, bool S2> ap_range_ref& operator=(const ap_int_base<W2, S2, false>& v) {
    assign64s(v.bits64(), S2 && ((long long)v.bits64() < 0));
    return *this;
  }
  template <int W2, bool S2> ap_range_ref& operator=(const ap_int_base<W2, S2, true>& v) {
    int n = hi - lo + 1;
    for (int b = 0; b < n; b += 64) {
      int c = (n - b) < 64 ? (n - b) : 64;
      p->set_bits(lo + b, c, (b < W2) ? v.get_bits(b, c) : 0ull);
    }
    return *this;
  }
  template <int W2, bool S2> ap_range_ref& operator=(const ap_range_ref<W2, S2>& r) {
    int n = hi - lo + 1, n2 = r.hi - r.lo + 1;
    for (int b = 0; b < n; b += 64) {
      int c = (n - b) < 64 ? (n - b) : 64;
      uint64_t v = 0;
      if (b < n2) {
        int c2 = (n2 - b) < c ? (n2 - b) : c;
        v = r.p->get_bits(r.lo + b, c2);
      }
      p->set_bits(lo + b, c, v);
    }
    return *this;
  }
  ap_range_ref& operator=(const ap_range_ref& r) { return this->template operator=<W, S>(r); }
};

template <int W, bool S> struct ap_bit_ref {
  ap_int_base<W, S>* p;
  int i;
  ap_bit_ref(ap_int_base<W, S>* p_, int i_) : p(p_), i(i_) {}
  operator bool() const { return p->get_bits(i, 1) != 0; }
  ap_bit_ref& operator=(bool v) { p->set_bits(i, 1, v ? 1u : 0u); return *this; }
  ap_bit_ref& operator=(int v) { p->set_bits(i, 1, (v & 1) ? 1u : 0u); return *this; }
  ap_bit_ref& operator=(unsigned long long v) { p->set_bits(i, 1, v & 1u); return *this; }
  ap_bit_ref& operator=(const ap_bit_ref& o) { p->set_bits(i, 1, (bool)o ? 1u : 0u); return *this; }
  template <int W2, bool S2> ap_bit_ref& operator=(const ap_bit_ref<W2, S2>& o) {
    p->set_bits(i, 1, (bool)o ? 1u : 0u);
    return *this;
  }
};

#endif  // FCB_ORACLE_SHIM_AP_INT_H

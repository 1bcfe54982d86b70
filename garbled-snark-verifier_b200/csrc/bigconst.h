// bigconst.h -- tiny fixed-width unsigned integers for the off-circuit constants that the
// gadgets bake into the topology (moduli, Montgomery constants, p/2, p/3 ...).  The reference
// gets these from num-bigint / ark-ff at run time (src/gadgets/bn254/fq.rs:57-77,
// fp254impl.rs:23-77); here they are re-derived and self-checked in tests.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace gsv {

struct U256 {
  std::array<uint64_t, 4> l{{0, 0, 0, 0}};
  U256() {}
  explicit U256(uint64_t v) { l[0] = v; }
  static U256 from_dec(const std::string& s) {
    U256 r;
    for (char ch : s) {
      if (ch < '0' || ch > '9') throw std::invalid_argument("bad decimal");
      unsigned __int128 carry = (unsigned)(ch - '0');
      for (int i = 0; i < 4; i++) {
        unsigned __int128 v = (unsigned __int128)r.l[i] * 10 + carry;
        r.l[i] = (uint64_t)v;
        carry = v >> 64;
      }
    }
    return r;
  }
  static U256 from_hex(const std::string& s) {
    U256 r;
    for (char ch : s) {
      unsigned d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
      for (int i = 3; i > 0; i--) r.l[i] = (r.l[i] << 4) | (r.l[i - 1] >> 60);
      r.l[0] = (r.l[0] << 4) | d;
    }
    return r;
  }
  std::string to_hex() const {
    static const char* H = "0123456789abcdef";
    std::string s;
    for (int i = 3; i >= 0; i--)
      for (int j = 60; j >= 0; j -= 4) s.push_back(H[(l[i] >> j) & 15]);
    return s;
  }
  bool bit(unsigned i) const { return i < 256 && ((l[i / 64] >> (i % 64)) & 1); }
  bool is_zero() const { return !(l[0] | l[1] | l[2] | l[3]); }
  unsigned bits() const {
    for (int i = 255; i >= 0; i--)
      if (bit((unsigned)i)) return (unsigned)i + 1;
    return 0;
  }
  bool operator==(const U256& o) const { return l == o.l; }
  bool operator!=(const U256& o) const { return !(l == o.l); }
  bool operator<(const U256& o) const {
    for (int i = 3; i >= 0; i--)
      if (l[i] != o.l[i]) return l[i] < o.l[i];
    return false;
  }
  // LSB-first bit vector of `len` bits (bits_from_biguint_with_len, bigint/mod.rs:33-47)
  std::vector<bool> bits_le(size_t len) const {
    if (bits() > len) throw std::overflow_error("constant does not fit");
    std::vector<bool> v(len);
    for (size_t i = 0; i < len; i++) v[i] = bit((unsigned)i);
    return v;
  }
};

inline U256 add(const U256& a, const U256& b, bool* carry_out = nullptr) {
  U256 r;
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (unsigned __int128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (carry_out) *carry_out = (bool)c;
  return r;
}
inline U256 sub(const U256& a, const U256& b) {  // wrapping
  U256 r;
  unsigned __int128 br = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 d = (unsigned __int128)a.l[i] - b.l[i] - br;
    r.l[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  return r;
}
inline U256 shl1(const U256& a) {
  U256 r;
  for (int i = 3; i > 0; i--) r.l[i] = (a.l[i] << 1) | (a.l[i - 1] >> 63);
  r.l[0] = a.l[0] << 1;
  return r;
}
inline U256 shr1(const U256& a) {
  U256 r;
  for (int i = 0; i < 3; i++) r.l[i] = (a.l[i] >> 1) | (a.l[i + 1] << 63);
  r.l[3] = a.l[3] >> 1;
  return r;
}
// modular helpers; m < 2^255 assumed (BN254 moduli are 254-bit)
inline U256 addmod(const U256& a, const U256& b, const U256& m) {
  U256 r = add(a, b);
  if (!(r < m)) r = sub(r, m);
  return r;
}
inline U256 submod(const U256& a, const U256& b, const U256& m) {
  return (a < b) ? sub(add(a, m), b) : sub(a, b);
}
inline U256 mulmod(const U256& a, const U256& b, const U256& m) {
  U256 r, x = a;
  for (unsigned i = 0; i < 256; i++) {
    if (b.bit(i)) r = addmod(r, x, m);
    x = addmod(x, x, m);
  }
  return r;
}
inline U256 powmod(const U256& a, const U256& e, const U256& m) {
  U256 r(1), x = a;
  unsigned n = e.bits();
  for (unsigned i = 0; i < n; i++) {
    if (e.bit(i)) r = mulmod(r, x, m);
    x = mulmod(x, x, m);
  }
  return r;
}
inline U256 invmod(const U256& a, const U256& p) { return powmod(a, sub(p, U256(2)), p); }  // p prime

}  // namespace gsv

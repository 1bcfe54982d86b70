// host_chain.cpp -- AES-NI fold of interleaved ciphertext streams (see host_chain.h).
#include "host_chain.h"

#include <immintrin.h>
#include <wmmintrin.h>

#include <cstdlib>

namespace gsv {
namespace {

struct RoundKeys {
  __m128i k[11];
};

template <int RC>
inline __m128i expand_step(__m128i key) {
  __m128i t = _mm_aeskeygenassist_si128(key, RC);
  t = _mm_shuffle_epi32(t, 0xFF);
  key = _mm_xor_si128(key, _mm_slli_si128(key, 4));
  key = _mm_xor_si128(key, _mm_slli_si128(key, 4));
  key = _mm_xor_si128(key, _mm_slli_si128(key, 4));
  return _mm_xor_si128(key, t);
}

// AES-128 schedule of the fixed key 0x42 x 16 (src/hashers/aes_ni.rs:10-13)
RoundKeys make_keys() {
  RoundKeys r;
  r.k[0] = _mm_set1_epi8(0x42);
  r.k[1] = expand_step<0x01>(r.k[0]);
  r.k[2] = expand_step<0x02>(r.k[1]);
  r.k[3] = expand_step<0x04>(r.k[2]);
  r.k[4] = expand_step<0x08>(r.k[3]);
  r.k[5] = expand_step<0x10>(r.k[4]);
  r.k[6] = expand_step<0x20>(r.k[5]);
  r.k[7] = expand_step<0x40>(r.k[6]);
  r.k[8] = expand_step<0x80>(r.k[7]);
  r.k[9] = expand_step<0x1B>(r.k[8]);
  r.k[10] = expand_step<0x36>(r.k[9]);
  return r;
}

// The chain's critical path is one AES per ciphertext.  aesenclast(s, rk) = SubBytes(ShiftRows(s)) ^ rk, so
// the last round of step p and the whitening of step p + 1 (h ^ ct ^ k0) fold into ONE instruction with the
// round key k10 ^ k0 ^ ct[p + 1], which does not depend on the chain: 10 dependent AES instructions per
// ciphertext and nothing else.  `t` below is the state after round 9 of the pending step.
template <int W>
void fold_w(uint8_t* h, const uint8_t* base, size_t pos_bytes, size_t inst_bytes, size_t n_pos, const RoundKeys& rk) {
  if (n_pos == 0) return;
  __m128i t[W];
  const uint8_t* src[W];
  const __m128i k10_0 = _mm_xor_si128(rk.k[10], rk.k[0]);
  for (int i = 0; i < W; i++) {
    src[i] = base + i * inst_bytes;
    t[i] = _mm_xor_si128(_mm_xor_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(h) + i),
                                       _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[i]))), rk.k[0]);
    src[i] += pos_bytes;
  }
  for (size_t p = 1;; p++) {
    for (int r = 1; r < 10; r++)
      for (int i = 0; i < W; i++) t[i] = _mm_aesenc_si128(t[i], rk.k[r]);
    if (p == n_pos) break;
    for (int i = 0; i < W; i++) {
      const __m128i key = _mm_xor_si128(k10_0, _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[i])));
      t[i] = _mm_aesenclast_si128(t[i], key);
      src[i] += pos_bytes;
    }
  }
  for (int i = 0; i < W; i++)
    _mm_storeu_si128(reinterpret_cast<__m128i*>(h) + i, _mm_aesenclast_si128(t[i], rk.k[10]));
}

// VAES (AVX-512): four chains per 512-bit register, Z registers interleaved -> 4 * Z chains per thread
// at one vaesenc per cycle instead of 8 chains of 128-bit aesenc.
// the 16-byte blocks of four instance streams side by side in one 512-bit register
__attribute__((target("avx512f,avx512vl,vaes"))) static inline __m512i load4(const uint8_t* const* s) {
  __m512i c = _mm512_castsi128_si512(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s[0])));
  c = _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(s[1])), 1);
  c = _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(s[2])), 2);
  return _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(s[3])), 3);
}

template <int Z>
__attribute__((target("avx512f,avx512vl,vaes"))) void fold_vaes(uint8_t* h, const uint8_t* base, size_t pos_bytes,
                                                                size_t inst_bytes, size_t n_pos, const RoundKeys& rk) {
  if (n_pos == 0) return;
  __m512i k[11];
  for (int r = 0; r < 11; r++) k[r] = _mm512_broadcast_i32x4(rk.k[r]);
  const __m512i k10_0 = _mm512_xor_si512(k[10], k[0]);
  __m512i t[Z];
  const uint8_t* src[Z][4];
  for (int z = 0; z < Z; z++) {
    for (int j = 0; j < 4; j++) src[z][j] = base + (size_t)(4 * z + j) * inst_bytes;
    t[z] = _mm512_xor_si512(_mm512_xor_si512(_mm512_loadu_si512(h + 64 * z), load4(src[z])), k[0]);
    for (int j = 0; j < 4; j++) src[z][j] += pos_bytes;
  }
  for (size_t p = 1;; p++) {
    for (int r = 1; r < 10; r++)
      for (int z = 0; z < Z; z++) t[z] = _mm512_aesenc_epi128(t[z], k[r]);
    if (p == n_pos) break;
    for (int z = 0; z < Z; z++) {
      t[z] = _mm512_aesenclast_epi128(t[z], _mm512_xor_si512(k10_0, load4(src[z])));  // see fold_w
      for (int j = 0; j < 4; j++) src[z][j] += pos_bytes;
    }
  }
  for (int z = 0; z < Z; z++) _mm512_storeu_si512(h + 64 * z, _mm512_aesenclast_epi128(t[z], k[10]));
}

// quad layout: Z quads side by side, each step of a quad is one 64-byte row
template <int Z>
__attribute__((target("avx512f,avx512vl,vaes"))) void fold_quads_vaes(uint8_t* h, const uint8_t* base, size_t quad_bytes,
                                                                      size_t n_pos, const RoundKeys& rk) {
  if (n_pos == 0) return;
  __m512i k[11];
  for (int r = 0; r < 11; r++) k[r] = _mm512_broadcast_i32x4(rk.k[r]);
  const __m512i k10_0 = _mm512_xor_si512(k[10], k[0]);
  __m512i t[Z];
  const uint8_t* src[Z];
  for (int z = 0; z < Z; z++) {
    src[z] = base + (size_t)z * quad_bytes;
    t[z] = _mm512_xor_si512(_mm512_xor_si512(_mm512_loadu_si512(h + 64 * z), _mm512_loadu_si512(src[z])), k[0]);
    src[z] += 64;
  }
  for (size_t p = 1;; p++) {
    for (int r = 1; r < 10; r++)
      for (int z = 0; z < Z; z++) t[z] = _mm512_aesenc_epi128(t[z], k[r]);
    if (p == n_pos) break;
    for (int z = 0; z < Z; z++) {
      t[z] = _mm512_aesenclast_epi128(t[z], _mm512_xor_si512(k10_0, _mm512_loadu_si512(src[z])));  // see fold_w
      src[z] += 64;
    }
  }
  for (int z = 0; z < Z; z++) _mm512_storeu_si512(h + 64 * z, _mm512_aesenclast_epi128(t[z], k[10]));
}

bool have_vaes() {
  static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
                        __builtin_cpu_supports("vaes") && !getenv("GSV_HOST_CHAIN_NO_VAES");
  return v;
}

}  // namespace

bool host_chain_available() { return __builtin_cpu_supports("aes") && __builtin_cpu_supports("sse4.1"); }

void host_chain_fold(uint8_t* h, const uint8_t* base, size_t pos_stride, size_t inst_stride, size_t n_pos,
                     uint32_t n_inst) {
  static const RoundKeys rk = make_keys();
  const size_t pb = pos_stride * 16, ib = inst_stride * 16;
  uint32_t i = 0;
  if (have_vaes()) {
    for (; i + 16 <= n_inst; i += 16) fold_vaes<4>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
    if (i + 8 <= n_inst) {
      fold_vaes<2>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
      i += 8;
    }
    if (i + 4 <= n_inst) {
      fold_vaes<1>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
      i += 4;
    }
  }
  for (; i + 8 <= n_inst; i += 8) fold_w<8>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
  if (i + 4 <= n_inst) {
    fold_w<4>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
    i += 4;
  }
  if (i + 2 <= n_inst) {
    fold_w<2>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
    i += 2;
  }
  if (i < n_inst) fold_w<1>(h + 16 * i, base + ib * i, pb, ib, n_pos, rk);
}


// W chains from W unrelated buffers (see fold_w for the merged last round)
template <int W>
void fold_ptrs(uint8_t* h, const uint8_t* const* src0, size_t n_pos, const RoundKeys& rk) {
  if (n_pos == 0) return;
  __m128i t[W];
  const uint8_t* src[W];
  const __m128i k10_0 = _mm_xor_si128(rk.k[10], rk.k[0]);
  for (int i = 0; i < W; i++) {
    src[i] = src0[i];
    t[i] = _mm_xor_si128(_mm_xor_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(h) + i),
                                       _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[i]))), rk.k[0]);
    src[i] += 16;
  }
  for (size_t p = 1;; p++) {
    for (int r = 1; r < 10; r++)
      for (int i = 0; i < W; i++) t[i] = _mm_aesenc_si128(t[i], rk.k[r]);
    if (p == n_pos) break;
    for (int i = 0; i < W; i++) {
      t[i] = _mm_aesenclast_si128(t[i], _mm_xor_si128(k10_0, _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[i]))));
      src[i] += 16;
    }
  }
  for (int i = 0; i < W; i++)
    _mm_storeu_si128(reinterpret_cast<__m128i*>(h) + i, _mm_aesenclast_si128(t[i], rk.k[10]));
}

void host_chain_fold_streams(uint8_t* h, const uint8_t* const* streams, uint64_t first, size_t n_pos, uint32_t n_inst) {
  static const RoundKeys rk = make_keys();
  const uint8_t* p[4];
  uint32_t i = 0;
  for (; i + 4 <= n_inst; i += 4) {
    for (int j = 0; j < 4; j++) p[j] = streams[i + j] + first * 16;
    fold_ptrs<4>(h + 16 * i, p, n_pos, rk);
  }
  for (; i < n_inst; i++) {
    p[0] = streams[i] + first * 16;
    fold_ptrs<1>(h + 16 * i, p, n_pos, rk);
  }
}

void host_chain_fold_quads(uint8_t* h, const uint8_t* base, size_t quad_bytes, size_t n_pos, uint32_t n_quads) {
  static const RoundKeys rk = make_keys();
  uint32_t q = 0;
  if (have_vaes()) {
    for (; q + 4 <= n_quads; q += 4) fold_quads_vaes<4>(h + 64 * q, base + quad_bytes * q, quad_bytes, n_pos, rk);
    if (q + 3 <= n_quads) {
      fold_quads_vaes<3>(h + 64 * q, base + quad_bytes * q, quad_bytes, n_pos, rk);
      q += 3;
    }
    if (q + 2 <= n_quads) {
      fold_quads_vaes<2>(h + 64 * q, base + quad_bytes * q, quad_bytes, n_pos, rk);
      q += 2;
    }
    if (q < n_quads) {
      fold_quads_vaes<1>(h + 64 * q, base + quad_bytes * q, quad_bytes, n_pos, rk);
      q++;
    }
    return;
  }
  // AES-NI only: one quad (4 interleaved chains) at a time; rows are 64 bytes apart, chains 16
  for (; q < n_quads; q++) fold_w<4>(h + 64 * q, base + quad_bytes * q, 64, 16, n_pos, rk);
}

}  // namespace gsv

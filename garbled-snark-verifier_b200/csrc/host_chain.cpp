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

template <int W>
void fold_w(uint8_t* h, const uint8_t* base, size_t pos_bytes, size_t inst_bytes, size_t n_pos, const RoundKeys& rk) {
  __m128i s[W];
  const uint8_t* src[W];
  for (int i = 0; i < W; i++) {
    s[i] = _mm_loadu_si128(reinterpret_cast<const __m128i*>(h) + i);
    src[i] = base + i * inst_bytes;
  }
  for (size_t p = 0; p < n_pos; p++) {
    for (int i = 0; i < W; i++) {
      s[i] = _mm_xor_si128(_mm_xor_si128(s[i], _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[i]))), rk.k[0]);
      src[i] += pos_bytes;
    }
    for (int r = 1; r < 10; r++)
      for (int i = 0; i < W; i++) s[i] = _mm_aesenc_si128(s[i], rk.k[r]);
    for (int i = 0; i < W; i++) s[i] = _mm_aesenclast_si128(s[i], rk.k[10]);
  }
  for (int i = 0; i < W; i++) _mm_storeu_si128(reinterpret_cast<__m128i*>(h) + i, s[i]);
}

// VAES (AVX-512): four chains per 512-bit register, Z registers interleaved -> 4 * Z chains per thread
// at one vaesenc per cycle instead of 8 chains of 128-bit aesenc.
template <int Z>
__attribute__((target("avx512f,avx512vl,vaes"))) void fold_vaes(uint8_t* h, const uint8_t* base, size_t pos_bytes,
                                                                size_t inst_bytes, size_t n_pos, const RoundKeys& rk) {
  __m512i k[11];
  for (int r = 0; r < 11; r++) k[r] = _mm512_broadcast_i32x4(rk.k[r]);
  __m512i s[Z];
  const uint8_t* src[Z][4];
  for (int z = 0; z < Z; z++) {
    s[z] = _mm512_loadu_si512(h + 64 * z);
    for (int j = 0; j < 4; j++) src[z][j] = base + (size_t)(4 * z + j) * inst_bytes;
  }
  for (size_t p = 0; p < n_pos; p++) {
    for (int z = 0; z < Z; z++) {
      __m512i c = _mm512_castsi128_si512(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src[z][0])));
      c = _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[z][1])), 1);
      c = _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[z][2])), 2);
      c = _mm512_inserti32x4(c, _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[z][3])), 3);
      s[z] = _mm512_xor_si512(_mm512_xor_si512(s[z], c), k[0]);
      for (int j = 0; j < 4; j++) src[z][j] += pos_bytes;
    }
    for (int r = 1; r < 10; r++)
      for (int z = 0; z < Z; z++) s[z] = _mm512_aesenc_epi128(s[z], k[r]);
    for (int z = 0; z < Z; z++) s[z] = _mm512_aesenclast_epi128(s[z], k[10]);
  }
  for (int z = 0; z < Z; z++) _mm512_storeu_si512(h + 64 * z, s[z]);
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

}  // namespace gsv

// host_chain.cpp -- AES-NI fold of interleaved ciphertext streams (see host_chain.h).
#include "host_chain.h"

#include <immintrin.h>
#include <wmmintrin.h>

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

}  // namespace

bool host_chain_available() { return __builtin_cpu_supports("aes") && __builtin_cpu_supports("sse4.1"); }

void host_chain_fold(uint8_t* h, const uint8_t* base, size_t pos_stride, size_t inst_stride, size_t n_pos,
                     uint32_t n_inst) {
  static const RoundKeys rk = make_keys();
  const size_t pb = pos_stride * 16, ib = inst_stride * 16;
  uint32_t i = 0;
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

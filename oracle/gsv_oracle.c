/*
 * gsv_oracle.c -- CPU ORACLE (test infrastructure, see gsv_oracle.h).
 *
 * Restates, function by function, the reference's garbling hot path:
 *   src/core/s.rs, src/core/delta.rs, src/core/gate_type.rs, src/hashers/mod.rs,
 *   src/hashers/aes_ni.rs, src/circuit/modes/garble_mode.rs,
 *   src/circuit/modes/garble_mode/halfgates_garbling.rs, src/circuit/modes/evaluate_mode.rs,
 *   src/circuit/modes/execute_mode.rs, src/ciphertext_hasher.rs, src/cut_and_choose/mod.rs.
 * Never used by the product path.
 */
#include "gsv_oracle.h"

#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) || defined(__i386__)
#include <cpuid.h>
#include <immintrin.h>
#define GSVO_X86 1
#endif

/* ------------------------------------------------------------------ AES-128 (FIPS-197) */
static uint8_t SBOX[256];
static uint32_t TE0[256], TE1[256], TE2[256], TE3[256];
static uint8_t FIXED_RK[11][16]; /* round keys of K = 0x42 * 16 */
static int g_init_done = 0;
static int g_have_aesni = 0;

static uint8_t gmul(uint8_t a, uint8_t b) {
  uint8_t p = 0;
  for (int i = 0; i < 8; i++) {
    if (b & 1) p ^= a;
    uint8_t hi = a & 0x80;
    a = (uint8_t)(a << 1);
    if (hi) a ^= 0x1b;
    b >>= 1;
  }
  return p;
}

static void key_expand(const uint8_t key[16], uint8_t rk[11][16]) {
  static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
  memcpy(rk[0], key, 16);
  for (int r = 1; r <= 10; r++) {
    const uint8_t* p = rk[r - 1];
    uint8_t t[4] = {SBOX[p[13]], SBOX[p[14]], SBOX[p[15]], SBOX[p[12]]};
    t[0] ^= rcon[r - 1];
    for (int i = 0; i < 4; i++) rk[r][i] = p[i] ^ t[i];
    for (int i = 4; i < 16; i++) rk[r][i] = p[i] ^ rk[r][i - 4];
  }
}

static void oracle_init(void) {
  if (g_init_done) return;
  /* S-box from the GF(2^8) inverse + affine map (FIPS-197 5.1.1). */
  for (int x = 0; x < 256; x++) {
    uint8_t inv = 0;
    if (x) {
      for (int y = 1; y < 256; y++)
        if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
    }
    uint8_t s = inv, r = inv;
    for (int i = 0; i < 4; i++) {
      r = (uint8_t)((r << 1) | (r >> 7));
      s ^= r;
    }
    SBOX[x] = s ^ 0x63;
  }
  for (int x = 0; x < 256; x++) {
    uint8_t s = SBOX[x], s2 = gmul(s, 2), s3 = gmul(s, 3);
    /* little-endian column word: byte0 = 2s, byte1 = s, byte2 = s, byte3 = 3s */
    TE0[x] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
    TE1[x] = (TE0[x] << 8) | (TE0[x] >> 24);
    TE2[x] = (TE0[x] << 16) | (TE0[x] >> 16);
    TE3[x] = (TE0[x] << 24) | (TE0[x] >> 8);
  }
  uint8_t k[16];
  memset(k, 0x42, 16);
  key_expand(k, FIXED_RK);
#ifdef GSVO_X86
  unsigned a, b, c, d;
  if (__get_cpuid(1, &a, &b, &c, &d)) g_have_aesni = (c >> 25) & 1;
#endif
  g_init_done = 1;
}

static inline uint32_t ld32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void st32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

static void aes_encrypt_rk(const uint8_t rk[11][16], const uint8_t in[16], uint8_t out[16]) {
  uint32_t s0 = ld32(in) ^ ld32(rk[0]), s1 = ld32(in + 4) ^ ld32(rk[0] + 4);
  uint32_t s2 = ld32(in + 8) ^ ld32(rk[0] + 8), s3 = ld32(in + 12) ^ ld32(rk[0] + 12);
  for (int r = 1; r < 10; r++) {
    uint32_t t0 = TE0[s0 & 255] ^ TE1[(s1 >> 8) & 255] ^ TE2[(s2 >> 16) & 255] ^ TE3[s3 >> 24] ^ ld32(rk[r]);
    uint32_t t1 = TE0[s1 & 255] ^ TE1[(s2 >> 8) & 255] ^ TE2[(s3 >> 16) & 255] ^ TE3[s0 >> 24] ^ ld32(rk[r] + 4);
    uint32_t t2 = TE0[s2 & 255] ^ TE1[(s3 >> 8) & 255] ^ TE2[(s0 >> 16) & 255] ^ TE3[s1 >> 24] ^ ld32(rk[r] + 8);
    uint32_t t3 = TE0[s3 & 255] ^ TE1[(s0 >> 8) & 255] ^ TE2[(s1 >> 16) & 255] ^ TE3[s2 >> 24] ^ ld32(rk[r] + 12);
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
  }
  uint32_t t0 = (uint32_t)SBOX[s0 & 255] | ((uint32_t)SBOX[(s1 >> 8) & 255] << 8) |
                ((uint32_t)SBOX[(s2 >> 16) & 255] << 16) | ((uint32_t)SBOX[s3 >> 24] << 24);
  uint32_t t1 = (uint32_t)SBOX[s1 & 255] | ((uint32_t)SBOX[(s2 >> 8) & 255] << 8) |
                ((uint32_t)SBOX[(s3 >> 16) & 255] << 16) | ((uint32_t)SBOX[s0 >> 24] << 24);
  uint32_t t2 = (uint32_t)SBOX[s2 & 255] | ((uint32_t)SBOX[(s3 >> 8) & 255] << 8) |
                ((uint32_t)SBOX[(s0 >> 16) & 255] << 16) | ((uint32_t)SBOX[s1 >> 24] << 24);
  uint32_t t3 = (uint32_t)SBOX[s3 & 255] | ((uint32_t)SBOX[(s0 >> 8) & 255] << 8) |
                ((uint32_t)SBOX[(s1 >> 16) & 255] << 16) | ((uint32_t)SBOX[s2 >> 24] << 24);
  st32(out, t0 ^ ld32(rk[10]));
  st32(out + 4, t1 ^ ld32(rk[10] + 4));
  st32(out + 8, t2 ^ ld32(rk[10] + 8));
  st32(out + 12, t3 ^ ld32(rk[10] + 12));
}

void gsvo_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
  oracle_init();
  uint8_t rk[11][16];
  key_expand(key, rk);
  aes_encrypt_rk(rk, in, out);
}

void gsvo_aes_fixed_portable(const uint8_t in[16], uint8_t out[16]) {
  oracle_init();
  aes_encrypt_rk(FIXED_RK, in, out);
}

#ifdef GSVO_X86
__attribute__((target("aes,sse2"))) static inline __m128i aesni_fixed(__m128i s) {
  const __m128i* rk = (const __m128i*)FIXED_RK;
  s = _mm_xor_si128(s, _mm_loadu_si128(rk));
  for (int r = 1; r < 10; r++) s = _mm_aesenc_si128(s, _mm_loadu_si128(rk + r));
  return _mm_aesenclast_si128(s, _mm_loadu_si128(rk + 10));
}
__attribute__((target("aes,sse2"))) static void aesni_fixed_bytes(const uint8_t in[16], uint8_t out[16]) {
  _mm_storeu_si128((__m128i*)out, aesni_fixed(_mm_loadu_si128((const __m128i*)in)));
}
/* two independent blocks, pipelined like Aes128::encrypt2_blocks (aes_ni.rs:~120-160) */
__attribute__((target("aes,sse2"))) static void aesni_fixed2(const uint8_t in0[16], const uint8_t in1[16],
                                                            uint8_t out0[16], uint8_t out1[16]) {
  const __m128i* rk = (const __m128i*)FIXED_RK;
  __m128i k = _mm_loadu_si128(rk);
  __m128i s0 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)in0), k);
  __m128i s1 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)in1), k);
  for (int r = 1; r < 10; r++) {
    k = _mm_loadu_si128(rk + r);
    s0 = _mm_aesenc_si128(s0, k);
    s1 = _mm_aesenc_si128(s1, k);
  }
  k = _mm_loadu_si128(rk + 10);
  _mm_storeu_si128((__m128i*)out0, _mm_aesenclast_si128(s0, k));
  _mm_storeu_si128((__m128i*)out1, _mm_aesenclast_si128(s1, k));
}
#endif

int gsvo_have_aesni(void) {
  oracle_init();
  return g_have_aesni;
}

int gsvo_aes_fixed_aesni(const uint8_t in[16], uint8_t out[16]) {
  oracle_init();
#ifdef GSVO_X86
  if (g_have_aesni) {
    aesni_fixed_bytes(in, out);
    return 0;
  }
#endif
  (void)in; (void)out;
  return -1;
}

void gsvo_aes_fixed(const uint8_t in[16], uint8_t out[16]) {
  oracle_init();
#ifdef GSVO_X86
  if (g_have_aesni) {
    aesni_fixed_bytes(in, out);
    return;
  }
#endif
  aes_encrypt_rk(FIXED_RK, in, out);
}

static void aes_fixed2(const uint8_t in0[16], const uint8_t in1[16], uint8_t out0[16], uint8_t out1[16]) {
#ifdef GSVO_X86
  if (g_have_aesni) {
    aesni_fixed2(in0, in1, out0, out1);
    return;
  }
#endif
  aes_encrypt_rk(FIXED_RK, in0, out0);
  aes_encrypt_rk(FIXED_RK, in1, out1);
}

/* ------------------------------------------------------------------ gate hashers */
void gsvo_tweak(uint64_t gid, uint8_t out[16]) {
  /* src/hashers/mod.rs:56-64: t0 = gid ^ 0x1234..., t1 = gid * 0xDEAD... (wrapping);
   * mod.rs:88-95: bytes = t0.to_le_bytes() || t1.to_le_bytes(). */
  uint64_t t0 = gid ^ 0x123456789ABCDEF0ull;
  uint64_t t1 = gid * 0xDEADBEEFCAFEBABEull;
  for (int i = 0; i < 8; i++) {
    out[i] = (uint8_t)(t0 >> (8 * i));
    out[8 + i] = (uint8_t)(t1 >> (8 * i));
  }
}

void gsvo_hash_aes(const uint8_t x[16], uint64_t gid, uint8_t out[16]) {
  uint8_t tw[16], in[16];
  gsvo_tweak(gid, tw);
  for (int i = 0; i < 16; i++) in[i] = x[i] ^ tw[i];
  gsvo_aes_fixed(in, out);
}

static void hash_aes2(const uint8_t x0[16], const uint8_t x1[16], uint64_t gid, uint8_t o0[16], uint8_t o1[16]) {
  uint8_t tw[16], i0[16], i1[16];
  gsvo_tweak(gid, tw);
  for (int i = 0; i < 16; i++) {
    i0[i] = x0[i] ^ tw[i];
    i1[i] = x1[i] ^ tw[i];
  }
  oracle_init();
  aes_fixed2(i0, i1, o0, o1);
}

/* ------------------------------------------------------------------ BLAKE3 (single block) */
static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
#define B3_G(a, b, c, d, mx, my)        \
  do {                                  \
    v[a] = v[a] + v[b] + (mx);          \
    v[d] = rotr32(v[d] ^ v[a], 16);     \
    v[c] = v[c] + v[d];                 \
    v[b] = rotr32(v[b] ^ v[c], 12);     \
    v[a] = v[a] + v[b] + (my);          \
    v[d] = rotr32(v[d] ^ v[a], 8);      \
    v[c] = v[c] + v[d];                 \
    v[b] = rotr32(v[b] ^ v[c], 7);      \
  } while (0)

static void blake3_compress_root(const uint32_t block[16], uint32_t block_len, uint32_t out[8]) {
  /* flags = CHUNK_START | CHUNK_END | ROOT = 1 | 2 | 8, counter 0, default IV as chaining value */
  uint32_t v[16], m[16], t[16];
  for (int i = 0; i < 8; i++) v[i] = B3_IV[i];
  for (int i = 0; i < 4; i++) v[8 + i] = B3_IV[i];
  v[12] = 0; v[13] = 0; v[14] = block_len; v[15] = 1u | 2u | 8u;
  memcpy(m, block, sizeof(m));
  for (int r = 0; r < 7; r++) {
    B3_G(0, 4, 8, 12, m[0], m[1]);
    B3_G(1, 5, 9, 13, m[2], m[3]);
    B3_G(2, 6, 10, 14, m[4], m[5]);
    B3_G(3, 7, 11, 15, m[6], m[7]);
    B3_G(0, 5, 10, 15, m[8], m[9]);
    B3_G(1, 6, 11, 12, m[10], m[11]);
    B3_G(2, 7, 8, 13, m[12], m[13]);
    B3_G(3, 4, 9, 14, m[14], m[15]);
    for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
    memcpy(m, t, sizeof(m));
  }
  for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}

int gsvo_blake3_small(const uint8_t* msg, size_t len, uint8_t out[32]) {
  if (len > 64) return -1;
  uint8_t blk[64];
  memset(blk, 0, 64);
  if (len) memcpy(blk, msg, len);
  uint32_t w[16], h[8];
  for (int i = 0; i < 16; i++) w[i] = ld32(blk + 4 * i);
  blake3_compress_root(w, (uint32_t)len, h);
  for (int i = 0; i < 8; i++) st32(out + 4 * i, h[i]);
  return 0;
}

void gsvo_hash_blake3(const uint8_t x[16], uint64_t gid, uint8_t out[16]) {
  /* hasher.update(label.to_bytes()); hasher.update(gate_id.to_le_bytes()) (usize = 8 bytes) */
  uint8_t msg[24], h[32];
  memcpy(msg, x, 16);
  for (int i = 0; i < 8; i++) msg[16 + i] = (uint8_t)(gid >> (8 * i));
  gsvo_blake3_small(msg, 24, h);
  memcpy(out, h, 16);
}

/* ------------------------------------------------------------------ half-gates */
static const uint8_t ALPHA[8][3] = {
    /* gate_type.rs:20-37 alphas_const: (alpha_a, alpha_b, alpha_c) */
    {0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};

static inline void xor16(uint8_t* d, const uint8_t* a, const uint8_t* b) {
  for (int i = 0; i < 16; i++) d[i] = a[i] ^ b[i];
}

int gsvo_garble_gate(int hasher, int gate_type, const uint8_t a0[16], const uint8_t b0[16],
                     const uint8_t delta[16], uint64_t gid, uint8_t c0_out[16], uint8_t ct_out[16]) {
  uint8_t t[16];
  switch (gate_type) {
    case GSVO_XOR:
      xor16(c0_out, a0, b0);
      return 0;
    case GSVO_XNOR:
      xor16(t, a0, b0);
      xor16(c0_out, t, delta);
      return 0;
    case GSVO_NOT:
      xor16(c0_out, a0, delta);
      return 0;
    default:
      break;
  }
  const uint8_t* al = ALPHA[gate_type];
  uint8_t sel[16], oth[16], h0[16], h1[16], bsel[16];
  if (al[0]) {
    xor16(sel, a0, delta);
    memcpy(oth, a0, 16);
  } else {
    memcpy(sel, a0, 16);
    xor16(oth, a0, delta);
  }
  if (hasher == GSVO_HASH_AES) {
    hash_aes2(sel, oth, gid, h0, h1);
  } else {
    gsvo_hash_blake3(sel, gid, h0);
    gsvo_hash_blake3(oth, gid, h1);
  }
  if (al[1]) xor16(bsel, b0, delta); else memcpy(bsel, b0, 16);
  xor16(t, h0, h1);
  xor16(ct_out, t, bsel);
  if (al[2]) xor16(c0_out, h0, delta); else memcpy(c0_out, h0, 16);
  return 1;
}

void gsvo_degarble_gate(int hasher, int gate_type, const uint8_t* ct, const uint8_t a_act[16],
                        int a_val, const uint8_t b_act[16], uint64_t gid, uint8_t c_out[16]) {
  switch (gate_type) {
    case GSVO_XOR:
    case GSVO_XNOR:
      xor16(c_out, a_act, b_act);
      return;
    case GSVO_NOT:
      memcpy(c_out, a_act, 16);
      return;
    default:
      break;
  }
  uint8_t h[16], t[16];
  if (hasher == GSVO_HASH_AES) gsvo_hash_aes(a_act, gid, h); else gsvo_hash_blake3(a_act, gid, h);
  if ((a_val != 0) != (ALPHA[gate_type][0] != 0)) {
    xor16(t, ct, h);
    xor16(c_out, t, b_act);
  } else {
    memcpy(c_out, h, 16);
  }
}

int gsvo_gate_eval(int gate_type, int a, int b) {
  a = !!a; b = !!b;
  switch (gate_type) {
    case GSVO_AND: return a & b;
    case GSVO_NAND: return !(a & b);
    case GSVO_NIMP: return a & (!b);
    case GSVO_IMP: return (!a) | b;
    case GSVO_NCIMP: return (!a) & b;
    case GSVO_CIMP: return (!b) | a;
    case GSVO_NOR: return !(a | b);
    case GSVO_OR: return a | b;
    case GSVO_XOR: return a ^ b;
    case GSVO_XNOR: return !(a ^ b);
    case GSVO_NOT: return !a;
  }
  return 0;
}

void gsvo_chain_update(uint8_t h[16], const uint8_t ct[16]) {
  uint8_t t[16];
  xor16(t, h, ct);
  gsvo_aes_fixed(t, h);
}

void gsvo_commit_label(const uint8_t label[16], uint8_t out[16]) { gsvo_aes_fixed(label, out); }

/* ------------------------------------------------------------------ RNG */
void gsvo_seed_key(uint64_t seed, uint8_t key_out[32]) {
  /* rand_core 0.6.4 SeedableRng::seed_from_u64: PCG32 (XSH RR) expands the u64 into the seed */
  const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
  uint64_t state = seed;
  for (int i = 0; i < 8; i++) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    st32(key_out + 4 * i, x);
  }
}

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define CC_QR(a, b, c, d)   \
  do {                      \
    a += b; d ^= a; d = rotl32(d, 16); \
    c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  \
    c += d; b ^= c; b = rotl32(b, 7);  \
  } while (0)

void gsvo_chacha20_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; i++) s[4 + i] = key[i];
  s[12] = (uint32_t)counter;
  s[13] = (uint32_t)(counter >> 32);
  s[14] = 0; /* stream id 0 (rand_chacha get_stream default) */
  s[15] = 0;
  uint32_t x[16];
  memcpy(x, s, sizeof(x));
  for (int i = 0; i < 10; i++) {
    CC_QR(x[0], x[4], x[8], x[12]);
    CC_QR(x[1], x[5], x[9], x[13]);
    CC_QR(x[2], x[6], x[10], x[14]);
    CC_QR(x[3], x[7], x[11], x[15]);
    CC_QR(x[0], x[5], x[10], x[15]);
    CC_QR(x[1], x[6], x[11], x[12]);
    CC_QR(x[2], x[7], x[8], x[13]);
    CC_QR(x[3], x[4], x[9], x[14]);
  }
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

void gsvo_rng_init(gsvo_rng* r, uint64_t seed) {
  uint8_t key[32];
  gsvo_seed_key(seed, key);
  for (int i = 0; i < 8; i++) r->key[i] = ld32(key + 4 * i);
  r->block = 0;
  r->pos = 16;
}

uint32_t gsvo_rng_u32(gsvo_rng* r) {
  if (r->pos >= 16) {
    gsvo_chacha20_block(r->key, r->block++, r->buf);
    r->pos = 0;
  }
  return r->buf[r->pos++];
}

uint64_t gsvo_rng_u64(gsvo_rng* r) {
  /* rand_core BlockRng::next_u64: two consecutive words, low word first */
  uint64_t lo = gsvo_rng_u32(r);
  uint64_t hi = gsvo_rng_u32(r);
  return lo | (hi << 32);
}

void gsvo_rng_label(gsvo_rng* r, uint8_t out[16]) {
  /* rand 0.8.5 Standard for u128: x = next_u64() (low half), y = next_u64() (high half) */
  uint64_t x = gsvo_rng_u64(r);
  uint64_t y = gsvo_rng_u64(r);
  for (int i = 0; i < 8; i++) {
    out[i] = (uint8_t)(y >> (56 - 8 * i));
    out[8 + i] = (uint8_t)(x >> (56 - 8 * i));
  }
}

/* ------------------------------------------------------------------ streams */
int gsvo_garble_stream(int hasher, uint64_t seed, const gsvo_stream* s, uint8_t* input_label0_out,
                       uint8_t* output_label0_out, uint8_t* ct_out, uint64_t ct_capacity,
                       gsvo_garble_summary* sum) {
  oracle_init();
  uint8_t* lab = (uint8_t*)malloc((size_t)s->n_wires * 16);
  uint8_t* set = (uint8_t*)calloc(s->n_wires, 1);
  if (!lab || !set) { free(lab); free(set); return -3; }
  gsvo_rng rng;
  gsvo_rng_init(&rng, seed);
  uint8_t delta[16];
  /* garble_mode.rs:81-85: delta, then false wire, then true wire */
  gsvo_rng_label(&rng, delta);
  gsvo_rng_label(&rng, lab + 16 * GSVO_WIRE_FALSE);
  gsvo_rng_label(&rng, lab + 16 * GSVO_WIRE_TRUE);
  set[0] = set[1] = 1;
  /* EncodeInput::encode: issue_garbled_wire() per input bit (garble_mode.rs:116-118) */
  for (uint32_t i = 0; i < s->n_inputs; i++) {
    gsvo_rng_label(&rng, lab + 16 * (size_t)(2 + i));
    set[2 + i] = 1;
  }
  if (input_label0_out) memcpy(input_label0_out, lab + 32, (size_t)s->n_inputs * 16);

  uint8_t h[16];
  memset(h, 0, 16);
  uint64_t n_ct = 0;
  int rc = 0;
  for (uint64_t g = 0; g < s->n_gates; g++) {
    uint32_t a = s->a[g], b = s->b[g], c = s->c[g];
    if (a >= s->n_wires || b >= s->n_wires || !set[a] || !set[b]) { rc = -1; break; }
    /* gate index advances for every gate, dead or not (garble_mode.rs:192) */
    uint64_t gid = g;
    if (c == GSVO_WIRE_DEAD) continue;
    if (c >= s->n_wires || c < 2) { rc = -1; break; }
    uint8_t c0[16], ct[16];
    int has_ct = gsvo_garble_gate(hasher, s->type[g], lab + 16 * (size_t)a, lab + 16 * (size_t)b, delta, gid, c0, ct);
    if (has_ct) {
      gsvo_chain_update(h, ct);
      if (ct_out && n_ct < ct_capacity) memcpy(ct_out + 16 * n_ct, ct, 16);
      n_ct++;
    }
    memcpy(lab + 16 * (size_t)c, c0, 16);
    set[c] = 1;
  }
  if (rc == 0 && output_label0_out) {
    for (uint32_t i = 0; i < s->n_outputs; i++) {
      uint32_t w = s->outputs[i];
      if (w >= s->n_wires || !set[w]) { rc = -1; break; }
      memcpy(output_label0_out + 16 * (size_t)i, lab + 16 * (size_t)w, 16);
    }
  }
  if (sum) {
    memcpy(sum->delta, delta, 16);
    memcpy(sum->false_label0, lab, 16);
    memcpy(sum->true_label0, lab + 16, 16);
    memcpy(sum->ct_commit, h, 16);
    sum->n_ct = n_ct;
    sum->n_gates = s->n_gates;
  }
  free(lab);
  free(set);
  return rc;
}

int gsvo_evaluate_stream(int hasher, const gsvo_stream* s, const uint8_t true_label[16],
                         const uint8_t false_label[16], const uint8_t* input_active,
                         const uint8_t* input_bits, const uint8_t* cts, uint64_t n_ct,
                         uint8_t* output_active_out, uint8_t* output_bits_out,
                         uint8_t ct_commit_out[16], uint64_t* n_ct_used) {
  oracle_init();
  uint8_t* lab = (uint8_t*)malloc((size_t)s->n_wires * 16);
  uint8_t* val = (uint8_t*)calloc(s->n_wires, 1);
  uint8_t* set = (uint8_t*)calloc(s->n_wires, 1);
  if (!lab || !val || !set) { free(lab); free(val); free(set); return -3; }
  /* evaluate_mode.rs:97-115: FALSE -> (false_wire, false), TRUE -> (true_wire, true) */
  memcpy(lab, false_label, 16); val[0] = 0;
  memcpy(lab + 16, true_label, 16); val[1] = 1;
  set[0] = set[1] = 1;
  for (uint32_t i = 0; i < s->n_inputs; i++) {
    memcpy(lab + 16 * (size_t)(2 + i), input_active + 16 * (size_t)i, 16);
    val[2 + i] = input_bits[i] ? 1 : 0;
    set[2 + i] = 1;
  }
  uint8_t h[16];
  memset(h, 0, 16);
  uint64_t used = 0;
  int rc = 0;
  for (uint64_t g = 0; g < s->n_gates; g++) {
    uint32_t a = s->a[g], b = s->b[g], c = s->c[g];
    if (a >= s->n_wires || b >= s->n_wires || !set[a] || !set[b]) { rc = -1; break; }
    uint64_t gid = g; /* evaluate_mode.rs:128 */
    if (c == GSVO_WIRE_DEAD) continue;
    if (c >= s->n_wires || c < 2) { rc = -1; break; }
    int t = s->type[g];
    const uint8_t* ct = NULL;
    if (t < GSVO_XOR) {
      /* ciphertext pulled unconditionally for every live non-free gate (halfgates_garbling.rs:57) */
      if (used >= n_ct) { rc = -2; break; }
      ct = cts + 16 * used;
      gsvo_chain_update(h, ct);
      used++;
    }
    uint8_t out[16];
    gsvo_degarble_gate(hasher, t, ct, lab + 16 * (size_t)a, val[a], lab + 16 * (size_t)b, gid, out);
    memcpy(lab + 16 * (size_t)c, out, 16);
    val[c] = (uint8_t)gsvo_gate_eval(t, val[a], val[b]);
    set[c] = 1;
  }
  if (rc == 0) {
    for (uint32_t i = 0; i < s->n_outputs; i++) {
      uint32_t w = s->outputs[i];
      if (w >= s->n_wires || !set[w]) { rc = -1; break; }
      if (output_active_out) memcpy(output_active_out + 16 * (size_t)i, lab + 16 * (size_t)w, 16);
      if (output_bits_out) output_bits_out[i] = val[w];
    }
  }
  if (ct_commit_out) memcpy(ct_commit_out, h, 16);
  if (n_ct_used) *n_ct_used = used;
  free(lab); free(val); free(set);
  return rc;
}

int gsvo_execute_stream(const gsvo_stream* s, const uint8_t* input_bits, uint8_t* output_bits_out) {
  uint8_t* val = (uint8_t*)calloc(s->n_wires, 1);
  if (!val) return -3;
  val[1] = 1;
  for (uint32_t i = 0; i < s->n_inputs; i++) val[2 + i] = input_bits[i] ? 1 : 0;
  for (uint64_t g = 0; g < s->n_gates; g++) {
    uint32_t c = s->c[g];
    if (c == GSVO_WIRE_DEAD) continue;
    val[c] = (uint8_t)gsvo_gate_eval(s->type[g], val[s->a[g]], val[s->b[g]]);
  }
  for (uint32_t i = 0; i < s->n_outputs; i++) output_bits_out[i] = val[s->outputs[i]];
  free(val);
  return 0;
}

int gsvo_compact_stream(const gsvo_stream* s, uint32_t* a2, uint32_t* b2, uint32_t* c2,
                        uint32_t* outputs2, uint32_t* n_slots) {
  const uint64_t NEVER = ~0ull;
  uint64_t* last = (uint64_t*)malloc((size_t)s->n_wires * sizeof(uint64_t));
  uint32_t* slot = (uint32_t*)malloc((size_t)s->n_wires * sizeof(uint32_t));
  uint32_t* free_list = (uint32_t*)malloc((size_t)s->n_wires * sizeof(uint32_t));
  if (!last || !slot || !free_list) { free(last); free(slot); free(free_list); return -3; }
  for (uint32_t w = 0; w < s->n_wires; w++) { last[w] = 0; slot[w] = GSVO_WIRE_DEAD; }
  for (uint64_t g = 0; g < s->n_gates; g++) {
    last[s->a[g]] = g + 1;
    last[s->b[g]] = g + 1;
  }
  for (uint32_t i = 0; i < s->n_outputs; i++) last[s->outputs[i]] = NEVER;
  last[0] = last[1] = NEVER;
  uint32_t next = 2 + s->n_inputs, n_free = 0;
  for (uint32_t w = 0; w < next; w++) slot[w] = w;
  for (uint64_t g = 0; g < s->n_gates; g++) {
    uint32_t a = s->a[g], b = s->b[g], c = s->c[g];
    a2[g] = slot[a];
    b2[g] = slot[b];
    if (last[a] == g + 1) free_list[n_free++] = slot[a];
    if (b != a && last[b] == g + 1) free_list[n_free++] = slot[b];
    if (c == GSVO_WIRE_DEAD) { c2[g] = GSVO_WIRE_DEAD; continue; }
    uint32_t sl = n_free ? free_list[--n_free] : next++;
    slot[c] = sl;
    c2[g] = sl;
    if (last[c] == 0) free_list[n_free++] = sl; /* written, never read, not an output */
  }
  for (uint32_t i = 0; i < s->n_outputs; i++) outputs2[i] = slot[s->outputs[i]];
  *n_slots = next;
  free(last); free(slot); free(free_list);
  return 0;
}


/* ---- garbling over the template DAG (see gsv_oracle.h) ------------------------------------- */
typedef struct {
  const gsvo_templates* t;
  int hasher;
  uint8_t delta[16];
  uint8_t h[16];
  uint64_t gid, n_ct;
  uint64_t max_gates; /* stop after this many gates (bounded CPU-baseline samples); 0 = whole circuit */
  uint8_t* arena;   /* label frames, 16 bytes per local wire */
  size_t cap;       /* in labels */
  int rc;
} walk_state;

static int walk_reserve(walk_state* w, size_t n) {
  if (n <= w->cap) return 0;
  size_t nc = w->cap * 2 > n ? w->cap * 2 : n;
  uint8_t* na = (uint8_t*)realloc(w->arena, nc * 16);
  if (!na) return -3;
  w->arena = na;
  w->cap = nc;
  return 0;
}

/* runs template ti whose frame starts at label index `base`; inputs already sit at base+2.. */
static void walk_run(walk_state* w, uint32_t ti, size_t base) {
  const uint32_t* T = w->t->tmpl + 12 * (size_t)ti;
  const uint32_t n_wires = T[1];
  const uint32_t* gates = w->t->gates + 4 * (size_t)T[2];
  const uint32_t* calls = w->t->calls + 3 * (size_t)T[4];
  const uint32_t* items = w->t->items + T[6];
  const uint32_t* cw = w->t->call_wires + T[8];
  for (uint32_t k = 0; k < T[7] && w->rc == 0; k++) {
    const uint32_t it = items[k];
    if (w->max_gates && w->gid >= w->max_gates) { w->rc = 1; return; }
    if (!(it & 0x80000000u)) {
      const uint32_t* g = gates + 4 * (size_t)it;
      const uint64_t gid = w->gid++;            /* every gate advances the index (garble_mode.rs:192) */
      if (g[2] == GSVO_WIRE_DEAD) continue;
      uint8_t c0[16], ct[16];
      if (gsvo_garble_gate(w->hasher, (int)g[3], w->arena + 16 * (base + g[0]), w->arena + 16 * (base + g[1]),
                           w->delta, gid, c0, ct)) {
        gsvo_chain_update(w->h, ct);
        w->n_ct++;
      }
      memcpy(w->arena + 16 * (base + g[2]), c0, 16);
    } else {
      const uint32_t* c = calls + 3 * (size_t)(it & 0x7FFFFFFFu);
      const uint32_t* C = w->t->tmpl + 12 * (size_t)c[0];
      const size_t cb = base + n_wires;
      if (walk_reserve(w, cb + C[1]) != 0) { w->rc = -3; return; }
      memcpy(w->arena + 16 * cb, w->arena + 16 * base, 32);  /* the two constant labels */
      for (uint32_t i = 0; i < C[0]; i++) {
        const uint32_t src = cw[c[1] + i];
        if (src == GSVO_WIRE_DEAD) memset(w->arena + 16 * (cb + 2 + i), 0, 16);
        else memcpy(w->arena + 16 * (cb + 2 + i), w->arena + 16 * (base + src), 16);
      }
      walk_run(w, c[0], cb);
      const uint32_t* couts = w->t->outs + C[10];
      for (uint32_t j = 0; j < C[11]; j++) {
        const uint32_t p = cw[c[2] + j], o = couts[j];
        if (p == GSVO_WIRE_DEAD || p < 2 || o == GSVO_WIRE_DEAD) continue;
        if (o >= 2 + C[0]) memcpy(w->arena + 16 * (base + p), w->arena + 16 * (cb + o), 16);  /* pass-throughs alias */
      }
    }
  }
}

int gsvo_garble_templates_prefix(int hasher, uint64_t seed, const gsvo_templates* t, uint64_t max_gates,
                                 uint8_t* input_label0_out, uint8_t* output_label0_out, gsvo_garble_summary* sum);
int gsvo_garble_templates(int hasher, uint64_t seed, const gsvo_templates* t, uint8_t* input_label0_out,
                          uint8_t* output_label0_out, gsvo_garble_summary* sum) {
  return gsvo_garble_templates_prefix(hasher, seed, t, 0, input_label0_out, output_label0_out, sum);
}

/* max_gates > 0: garble only the first max_gates gates of the emission order (returns 1, outputs not
 * written; the summary holds the gate / ciphertext counts and the chain hash so far). */
int gsvo_garble_templates_prefix(int hasher, uint64_t seed, const gsvo_templates* t, uint64_t max_gates,
                                 uint8_t* input_label0_out, uint8_t* output_label0_out, gsvo_garble_summary* sum) {
  oracle_init();
  if (!t || t->root >= t->n_templates) return -1;
  const uint32_t* R = t->tmpl + 12 * (size_t)t->root;
  walk_state w;
  memset(&w, 0, sizeof(w));
  w.t = t;
  w.hasher = hasher;
  w.max_gates = max_gates;
  if (walk_reserve(&w, (size_t)R[1] + (1u << 16)) != 0) return -3;
  gsvo_rng rng;
  gsvo_rng_init(&rng, seed);
  gsvo_rng_label(&rng, w.delta);               /* garble_mode.rs:81-85: delta, false, true, then inputs */
  gsvo_rng_label(&rng, w.arena);
  gsvo_rng_label(&rng, w.arena + 16);
  for (uint32_t i = 0; i < R[0]; i++) gsvo_rng_label(&rng, w.arena + 16 * (size_t)(2 + i));
  if (input_label0_out) memcpy(input_label0_out, w.arena + 32, (size_t)R[0] * 16);
  uint8_t consts[32];
  memcpy(consts, w.arena, 32);
  walk_run(&w, t->root, 0);
  if (w.rc == 0 && output_label0_out) {
    const uint32_t* outs = t->outs + R[10];
    for (uint32_t j = 0; j < R[11]; j++) {
      if (outs[j] == GSVO_WIRE_DEAD) memset(output_label0_out + 16 * (size_t)j, 0, 16);
      else memcpy(output_label0_out + 16 * (size_t)j, w.arena + 16 * (size_t)outs[j], 16);
    }
  }
  if (sum) {
    memcpy(sum->delta, w.delta, 16);
    memcpy(sum->false_label0, consts, 16);
    memcpy(sum->true_label0, consts + 16, 16);
    memcpy(sum->ct_commit, w.h, 16);
    sum->n_ct = w.n_ct;
    sum->n_gates = w.gid;
  }
  free(w.arena);
  return w.rc;
}

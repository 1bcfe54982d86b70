// device_hash.cuh -- sm_100a device primitives of the garbling hot path:
//   * fixed-key AES-128 (K = 0x42*16, src/hashers/aes_ni.rs:165) as T-table rounds out of
//     shared memory, one- and two-block (interleaved) forms;
//   * the gate-id tweak (src/hashers/mod.rs:56-64,88-95);
//   * single-block BLAKE3 of label || gid_le (src/hashers/mod.rs:35-51);
//   * half-gates garble / degarble (src/circuit/modes/garble_mode/halfgates_garbling.rs:5-69).
// Integer ALU + LDS work only: no tensor cores (this is not a contraction).
//
// A label is a uint4 whose memory bytes are S::to_bytes() (big-endian u128); AES column j is
// the little-endian word of bytes 4j..4j+3, i.e. exactly .x/.y/.z/.w.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gsvdev {

__constant__ uint32_t c_te0[256];  // Te0[x] = (2s, s, s, 3s) little-endian
__constant__ uint32_t c_rk[44];    // expanded round keys of the fixed key, LE words

__device__ __forceinline__ uint4 xor4(uint4 a, uint4 b) {
  return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
}
__device__ __forceinline__ uint4 and4(uint4 a, uint32_t m) {
  return make_uint4(a.x & m, a.y & m, a.z & m, a.w & m);
}

// Shared-memory table block, 64 KB, bank-conflict free by construction:
//   row r (256 bytes, r = 0..255): bytes [0,128)   = Te0[r] replicated for the 32 lanes,
//                                  bytes [128,256) = Te2[r] = rotl(Te0[r], 16) replicated.
// Lane l only ever touches word l (or 32 + l) of a row, i.e. bank l: every LDS is a single
// wavefront whatever the 32 indices are (the plain 4 KB layout measured 57 % conflict replays,
// profiles/r01_first_engine.md).  The byte address (r << 8) | (tbl << 7) | (l << 2) is formed by
// ONE PRMT from the state word and a per-lane constant, so a lookup costs PRMT + LDS.
// Te1 / Te3 are rotl8 of Te0 / Te2; rotation is linear, so one PRMT rotates the XOR of both.
constexpr int AES_TABLE_BYTES = 65536;

__device__ __forceinline__ void load_tables(uint32_t* te, int tid, int nthreads) {
  for (int i = tid; i < 256 * 64; i += nthreads) {
    const uint32_t v = c_te0[i >> 6];
    te[i] = (i & 32) ? ((v << 16) | (v >> 16)) : v;
  }
}
// per-lane PRMT operands: byte 0 = lane*4 (+128 for the Te2 half), other bytes zero
__device__ __forceinline__ uint32_t lane_sel0() { return (threadIdx.x & 31u) << 2; }

#define GSV_LK(S, J, LB) \
  (*reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(te) + __byte_perm((S), (LB), 0x5504 | ((J) << 4))))
#define GSV_ROT8(V) __byte_perm((V), 0, 0x2103)

// one column: Te0[b0(A)] ^ Te1[b1(B)] ^ Te2[b2(C)] ^ Te3[b3(D)] ^ rk
#define GSV_AES_COL(A, B, C, D, RK) \
  (GSV_LK(A, 0, lb0) ^ GSV_LK(C, 2, lb2) ^ (RK) ^ GSV_ROT8(GSV_LK(B, 1, lb0) ^ GSV_LK(D, 3, lb2)))

#define GSV_AES_ROUND(T, S, R)                          \
  T.x = GSV_AES_COL(S.x, S.y, S.z, S.w, c_rk[4 * (R) + 0]); \
  T.y = GSV_AES_COL(S.y, S.z, S.w, S.x, c_rk[4 * (R) + 1]); \
  T.z = GSV_AES_COL(S.z, S.w, S.x, S.y, c_rk[4 * (R) + 2]); \
  T.w = GSV_AES_COL(S.w, S.x, S.y, S.z, c_rk[4 * (R) + 3]);

// last round: SubBytes + ShiftRows only.  S-box bytes sit in Te2 byte 0 / byte 3 and Te0
// byte 1 / byte 2, so three PRMTs assemble the column.
#define GSV_AES_LASTCOL(A, B, C, D, RK)                                                     \
  (__byte_perm(__byte_perm(GSV_LK(A, 0, lb2), GSV_LK(B, 1, lb0), 0x7650),                    \
               __byte_perm(GSV_LK(C, 2, lb0), GSV_LK(D, 3, lb2), 0x7210), 0x7610) ^ (RK))
#define GSV_AES_LAST(T, S)                              \
  T.x = GSV_AES_LASTCOL(S.x, S.y, S.z, S.w, c_rk[40]); \
  T.y = GSV_AES_LASTCOL(S.y, S.z, S.w, S.x, c_rk[41]); \
  T.z = GSV_AES_LASTCOL(S.z, S.w, S.x, S.y, c_rk[42]); \
  T.w = GSV_AES_LASTCOL(S.w, S.x, S.y, S.z, c_rk[43]);

__device__ __forceinline__ uint4 aes_fixed(const uint32_t* __restrict__ te, uint4 in) {
  const uint32_t lb0 = lane_sel0(), lb2 = lb0 | 128u;
  uint4 s = make_uint4(in.x ^ c_rk[0], in.y ^ c_rk[1], in.z ^ c_rk[2], in.w ^ c_rk[3]);
  uint4 t;
#pragma unroll
  for (int r = 1; r < 9; r += 2) {
    GSV_AES_ROUND(t, s, r)
    GSV_AES_ROUND(s, t, r + 1)
  }
  GSV_AES_ROUND(t, s, 9)
  GSV_AES_LAST(s, t)
  return s;
}

// two independent blocks, rounds interleaved for ILP (the GPU analogue of encrypt2_blocks,
// src/hashers/aes_ni.rs:~120-160)
__device__ __forceinline__ void aes_fixed2(const uint32_t* __restrict__ te, uint4& a, uint4& b) {
  const uint32_t lb0 = lane_sel0(), lb2 = lb0 | 128u;
  uint4 s = make_uint4(a.x ^ c_rk[0], a.y ^ c_rk[1], a.z ^ c_rk[2], a.w ^ c_rk[3]);
  uint4 u = make_uint4(b.x ^ c_rk[0], b.y ^ c_rk[1], b.z ^ c_rk[2], b.w ^ c_rk[3]);
  uint4 t, v;
#pragma unroll
  for (int r = 1; r < 9; r += 2) {
    GSV_AES_ROUND(t, s, r)
    GSV_AES_ROUND(v, u, r)
    GSV_AES_ROUND(s, t, r + 1)
    GSV_AES_ROUND(u, v, r + 1)
  }
  GSV_AES_ROUND(t, s, 9)
  GSV_AES_ROUND(v, u, 9)
  GSV_AES_LAST(s, t)
  GSV_AES_LAST(u, v)
  a = s;
  b = u;
}

// tweak(gid) as a label-shaped mask: bytes = LE64(gid ^ C0) || LE64(gid * C1)
__device__ __forceinline__ uint4 tweak(uint64_t gid) {
  uint64_t t0 = gid ^ 0x123456789ABCDEF0ull;
  uint64_t t1 = gid * 0xDEADBEEFCAFEBABEull;
  return make_uint4((uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)t1, (uint32_t)(t1 >> 32));
}

// ---- BLAKE3, one 24-byte block: label(16) || gid_le(8); flags CHUNK_START|CHUNK_END|ROOT
__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
#define GSV_B3_G(a, b, c, d, mx, my) \
  a = a + b + (mx);                  \
  d = rotr(d ^ a, 16);               \
  c = c + d;                         \
  b = rotr(b ^ c, 12);               \
  a = a + b + (my);                  \
  d = rotr(d ^ a, 8);                \
  c = c + d;                         \
  b = rotr(b ^ c, 7);

__device__ __forceinline__ uint4 blake3_label(uint4 x, uint64_t gid) {
  // message words; only m0..m5 are non-zero, the schedule below is the standard permutation
  // applied r times to (0..15), written out so the zero words fold away at compile time.
  const uint32_t m0 = x.x, m1 = x.y, m2 = x.z, m3 = x.w, m4 = (uint32_t)gid, m5 = (uint32_t)(gid >> 32);
  const uint32_t Z = 0u;
  uint32_t v0 = 0x6A09E667u, v1 = 0xBB67AE85u, v2 = 0x3C6EF372u, v3 = 0xA54FF53Au;
  uint32_t v4 = 0x510E527Fu, v5 = 0x9B05688Cu, v6 = 0x1F83D9ABu, v7 = 0x5BE0CD19u;
  uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
  uint32_t v12 = 0u, v13 = 0u, v14 = 24u, v15 = 11u;
#define GSV_B3_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  GSV_B3_G(v0, v4, v8, v12, s0, s1)                                                          \
  GSV_B3_G(v1, v5, v9, v13, s2, s3)                                                          \
  GSV_B3_G(v2, v6, v10, v14, s4, s5)                                                         \
  GSV_B3_G(v3, v7, v11, v15, s6, s7)                                                         \
  GSV_B3_G(v0, v5, v10, v15, s8, s9)                                                         \
  GSV_B3_G(v1, v6, v11, v12, s10, s11)                                                       \
  GSV_B3_G(v2, v7, v8, v13, s12, s13)                                                        \
  GSV_B3_G(v3, v4, v9, v14, s14, s15)
  // round r uses m[SCHED[r][i]]; words 6..15 are zero
  // r0: 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15
  GSV_B3_ROUND(m0, m1, m2, m3, m4, m5, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z)
  // r1: 2 6 3 10 7 0 4 13 1 11 12 5 9 14 15 8
  GSV_B3_ROUND(m2, Z, m3, Z, Z, m0, m4, Z, m1, Z, Z, m5, Z, Z, Z, Z)
  // r2: 3 4 10 12 13 2 7 14 6 5 9 0 11 15 8 1
  GSV_B3_ROUND(m3, m4, Z, Z, Z, m2, Z, Z, Z, m5, Z, m0, Z, Z, Z, m1)
  // r3: 10 7 12 9 14 3 13 15 4 0 11 2 5 8 1 6
  GSV_B3_ROUND(Z, Z, Z, Z, Z, m3, Z, Z, m4, m0, Z, m2, m5, Z, m1, Z)
  // r4: 12 13 9 11 15 10 14 8 7 2 5 3 0 1 6 4
  GSV_B3_ROUND(Z, Z, Z, Z, Z, Z, Z, Z, Z, m2, m5, m3, m0, m1, Z, m4)
  // r5: 9 14 11 5 8 12 15 1 13 3 0 10 2 6 4 7
  GSV_B3_ROUND(Z, Z, Z, m5, Z, Z, Z, m1, Z, m3, m0, Z, m2, Z, m4, Z)
  // r6: 11 15 5 0 1 9 8 6 14 10 2 12 3 4 7 13
  GSV_B3_ROUND(Z, Z, m5, m0, m1, Z, Z, Z, Z, Z, m2, Z, m3, m4, Z, Z)
#undef GSV_B3_ROUND
  return make_uint4(v0 ^ v8, v1 ^ v9, v2 ^ v10, v3 ^ v11);
}

enum { HASH_AES = 0, HASH_BLAKE3 = 1 };

template <int HASH>
__device__ __forceinline__ uint4 hash1(const uint32_t* __restrict__ te, uint4 x, uint64_t gid) {
  if (HASH == HASH_AES) return aes_fixed(te, xor4(x, tweak(gid)));
  return blake3_label(x, gid);
}
template <int HASH>
__device__ __forceinline__ void hash2(const uint32_t* __restrict__ te, uint4& x0, uint4& x1, uint64_t gid) {
  if (HASH == HASH_AES) {
    uint4 tw = tweak(gid);
    x0 = xor4(x0, tw);
    x1 = xor4(x1, tw);
    aes_fixed2(te, x0, x1);
  } else {
    x0 = blake3_label(x0, gid);
    x1 = blake3_label(x1, gid);
  }
}

// garble_gate for the 8 AND-family types: (alpha_a, alpha_b, alpha_c) = bits 2,1,0 of the type
// discriminant (gate_type.rs:20-37).  Returns c0; writes the ciphertext.
template <int HASH>
__device__ __forceinline__ uint4 garble_nonfree(const uint32_t* __restrict__ te, uint32_t type, uint4 a0,
                                                uint4 b0, uint4 delta, uint64_t gid, uint4& ct) {
  const uint32_t ma = 0u - ((type >> 2) & 1u), mb = 0u - ((type >> 1) & 1u), mc = 0u - (type & 1u);
  uint4 h0 = xor4(a0, and4(delta, ma));  // selected
  uint4 h1 = xor4(h0, delta);            // other
  hash2<HASH>(te, h0, h1, gid);
  ct = xor4(xor4(h0, h1), xor4(b0, and4(delta, mb)));
  return xor4(h0, and4(delta, mc));
}

// degarble_gate for AND-family gates; plaintext value via ((a^aa)&(b^ab))^ac
template <int HASH>
__device__ __forceinline__ uint4 degarble_nonfree(const uint32_t* __restrict__ te, uint32_t type, uint4 ct,
                                                  uint4 a_act, uint32_t a_val, uint4 b_act, uint64_t gid) {
  uint4 h = hash1<HASH>(te, a_act, gid);
  const uint32_t m = 0u - ((a_val ^ (type >> 2)) & 1u);  // a_value != alpha_a
  return xor4(h, and4(xor4(ct, b_act), m));
}
__device__ __forceinline__ uint32_t gate_value(uint32_t type, uint32_t a, uint32_t b) {
  if (type < 8) return (((a ^ (type >> 2)) & (b ^ (type >> 1))) ^ type) & 1u;
  if (type == 8) return (a ^ b) & 1u;
  if (type == 9) return (a ^ b ^ 1u) & 1u;
  return (a ^ 1u) & 1u;
}

}  // namespace gsvdev

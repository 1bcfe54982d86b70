/*
 * gsv_oracle.h -- CPU ORACLE for the garbling hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's per-gate garbling algorithm
 * (BitVM/garbled-snark-verifier v0.4.0).  It is used ONLY by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs
 * as the checker and the CPU baseline.  The product path (libgsv_cuda.so) never
 * links, loads or calls anything in this directory.
 *
 * Parity status: the reference ships no golden vectors for this path and cannot be
 * compiled here (no Rust toolchain).  The primitives are pinned against FIPS-197,
 * the official BLAKE3 vectors (via the `blake3` python binding), RFC 7539 ChaCha20
 * and OpenSSL AES (python `cryptography`) in tests/test_oracle_primitives.py, and
 * against the survey-time known-answer table.  The RNG
 * expansion (rand 0.8.5 / rand_core 0.6.4 / rand_chacha 0.3.1, third-party crates
 * absent from /root/reference) is restated from their published algorithms:
 * "parity unpinned" for seed -> label derivation.
 *
 * Every label crosses this API as 16 bytes in S::to_bytes() order, i.e. the
 * big-endian bytes of the reference's u128 (src/core/s.rs:25-32).
 */
#ifndef GSV_ORACLE_H
#define GSV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Gate types, same discriminants as src/core/gate_type.rs:1-15. */
enum {
  GSVO_AND = 0, GSVO_NAND = 1, GSVO_NIMP = 2, GSVO_IMP = 3, GSVO_NCIMP = 4, GSVO_CIMP = 5,
  GSVO_NOR = 6, GSVO_OR = 7, GSVO_XOR = 8, GSVO_XNOR = 9, GSVO_NOT = 10
};
enum { GSVO_HASH_AES = 0, GSVO_HASH_BLAKE3 = 1 };

#define GSVO_WIRE_FALSE 0u
#define GSVO_WIRE_TRUE 1u
#define GSVO_WIRE_DEAD 0xFFFFFFFFu /* WireId::UNREACHABLE, src/core/wire.rs:8 */

/* ---- primitives ------------------------------------------------------------------ */
/* FIPS-197 AES-128 (generic key; used for the KATs). */
void gsvo_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]);
/* AES-128 under the reference's fixed key 0x42*16 (src/hashers/aes_ni.rs:165). */
void gsvo_aes_fixed(const uint8_t in[16], uint8_t out[16]);
/* Same, forced through the portable table code / the AES-NI code (returns -1 if absent). */
void gsvo_aes_fixed_portable(const uint8_t in[16], uint8_t out[16]);
int gsvo_aes_fixed_aesni(const uint8_t in[16], uint8_t out[16]);
int gsvo_have_aesni(void);
/* tweak(gid), src/hashers/mod.rs:56-64,88-95. */
void gsvo_tweak(uint64_t gid, uint8_t out[16]);
/* AesNiHasher::hash_with_gate<1>, src/hashers/mod.rs:79-86. */
void gsvo_hash_aes(const uint8_t x[16], uint64_t gid, uint8_t out[16]);
/* Blake3Hasher::hash_with_gate<1>, src/hashers/mod.rs:35-51. */
void gsvo_hash_blake3(const uint8_t x[16], uint64_t gid, uint8_t out[16]);
/* BLAKE3 of a message of at most 64 bytes (single block, single chunk). */
int gsvo_blake3_small(const uint8_t* msg, size_t len, uint8_t out[32]);

/* garble_gate, src/circuit/modes/garble_mode/halfgates_garbling.rs:5-38.
 * Returns 1 when a ciphertext was produced (non-free gate), else 0. */
int gsvo_garble_gate(int hasher, int gate_type, const uint8_t a0[16], const uint8_t b0[16],
                     const uint8_t delta[16], uint64_t gid, uint8_t c0_out[16], uint8_t ct_out[16]);
/* degarble_gate, halfgates_garbling.rs:41-69.  `ct` may be NULL for free gates. */
void gsvo_degarble_gate(int hasher, int gate_type, const uint8_t* ct, const uint8_t a_act[16],
                        int a_val, const uint8_t b_act[16], uint64_t gid, uint8_t c_out[16]);
/* GateType::f, src/core/gate_type.rs:38-60. */
int gsvo_gate_eval(int gate_type, int a, int b);

/* AESAccumulatingHash::update, src/ciphertext_hasher.rs:23-29. */
void gsvo_chain_update(uint8_t h[16], const uint8_t ct[16]);
/* commit_label, src/cut_and_choose/mod.rs:41-48. */
void gsvo_commit_label(const uint8_t label[16], uint8_t out[16]);

/* ---- RNG (rand_core 0.6.4 seed_from_u64 + rand_chacha 0.3.1 ChaCha20Rng) ------------ */
typedef struct {
  uint32_t key[8];
  uint64_t block;     /* next block counter */
  uint32_t buf[16];   /* current block */
  int pos;            /* next unread word in buf, 16 = empty */
} gsvo_rng;
void gsvo_seed_key(uint64_t seed, uint8_t key_out[32]);
void gsvo_rng_init(gsvo_rng* r, uint64_t seed);
uint32_t gsvo_rng_u32(gsvo_rng* r);
uint64_t gsvo_rng_u64(gsvo_rng* r);
/* rng.gen::<u128>() returned as S::to_bytes() (big-endian), src/core/s.rs:57-59. */
void gsvo_rng_label(gsvo_rng* r, uint8_t out[16]);
/* Raw ChaCha20 block (RFC 7539 layout with a 64-bit counter and zero stream id). */
void gsvo_chacha20_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]);

/* ---- whole-stream garbling / evaluation ------------------------------------------- */
/*
 * A flat gate stream in emission order (what StreamingMode::add_gate hands to the
 * mode, src/circuit/streaming_mode.rs:134-148).  Wire ids are SSA: 0 = FALSE,
 * 1 = TRUE, 2..2+n_inputs-1 = inputs in EncodeInput::encode order, then one id per
 * issued wire.  c == GSVO_WIRE_DEAD marks an UNREACHABLE output: the gate consumes
 * its gate index but is not garbled (garble_mode.rs:192-197).
 */
typedef struct {
  uint64_t n_gates;
  const uint8_t* type;
  const uint32_t* a;
  const uint32_t* b;
  const uint32_t* c;
  uint32_t n_wires;  /* upper bound on ids (exclusive) */
  uint32_t n_inputs;
  uint32_t n_outputs;
  const uint32_t* outputs;
} gsvo_stream;

typedef struct {
  uint8_t delta[16];
  uint8_t false_label0[16];
  uint8_t true_label0[16];
  uint8_t ct_commit[16];
  uint64_t n_ct;
  uint64_t n_gates;
} gsvo_garble_summary;

/*
 * GarbleMode::new + EncodeInput + evaluate_gate loop (garble_mode.rs:80-222).
 *   input_label0_out : n_inputs*16 bytes or NULL
 *   output_label0_out: n_outputs*16 bytes or NULL
 *   ct_out           : n_ct*16 bytes (emission order, gc_{i}.bin format) or NULL
 *   gid_base         : first gate index (0 for a whole circuit)
 * Returns 0, or -1 on a malformed stream (read of an unset wire).
 */
int gsvo_garble_stream(int hasher, uint64_t seed, const gsvo_stream* s, uint8_t* input_label0_out,
                       uint8_t* output_label0_out, uint8_t* ct_out, uint64_t ct_capacity,
                       gsvo_garble_summary* sum);

/*
 * The same garbling run over a circuit given as its template DAG (the product's
 * gsv_program_export_templates format) instead of a flat stream: a depth-first walk in emission
 * order with one label frame per active component, so circuits far too large to flatten (the
 * 11 G-gate Groth16 verifier) garble in a few MB.  Same RNG draw order, gate index (every gate,
 * dead ones included), ciphertext order and chain commitment as gsvo_garble_stream.
 */
typedef struct {
  uint32_t n_templates, root;
  const uint32_t* tmpl;       /* 12 words per template */
  const uint32_t* gates;      /* 4 words per gate */
  const uint32_t* calls;      /* 3 words per call */
  const uint32_t* items;
  const uint32_t* call_wires;
  const uint32_t* outs;
} gsvo_templates;
int gsvo_garble_templates(int hasher, uint64_t seed, const gsvo_templates* t, uint8_t* input_label0_out,
                          uint8_t* output_label0_out, gsvo_garble_summary* sum);
/* max_gates > 0: only the first max_gates gates of the emission order (bounded CPU-baseline samples);
 * returns 1, outputs not written, the summary holds the counts and the chain hash so far. */
int gsvo_garble_templates_prefix(int hasher, uint64_t seed, const gsvo_templates* t, uint64_t max_gates,
                                 uint8_t* input_label0_out, uint8_t* output_label0_out, gsvo_garble_summary* sum);

/*
 * EvaluateMode (evaluate_mode.rs:70-158).  input_active: n_inputs*16, input_bits: n_inputs.
 * cts: the garbler's stream.  Outputs: active label + bit per output wire, chain hash of
 * the consumed ciphertexts (what FileSource computes, ciphertext_source.rs:35-106).
 * Returns 0, -1 malformed, -2 "Ciphertext source exhausted".
 */
int gsvo_evaluate_stream(int hasher, const gsvo_stream* s, const uint8_t true_label[16],
                         const uint8_t false_label[16], const uint8_t* input_active,
                         const uint8_t* input_bits, const uint8_t* cts, uint64_t n_ct,
                         uint8_t* output_active_out, uint8_t* output_bits_out,
                         uint8_t ct_commit_out[16], uint64_t* n_ct_used);

/*
 * Renames the SSA wire ids of a stream to recycled slot ids (a wire's slot is freed after its
 * last read, outputs stay pinned) -- the static equivalent of the reference's credits slab
 * (src/storage.rs:119-198), which keeps the live set cache resident.  The result is again a
 * valid stream (n_wires = *n_slots).  Arrays a2/b2/c2 have n_gates entries, outputs2 n_outputs.
 */
int gsvo_compact_stream(const gsvo_stream* s, uint32_t* a2, uint32_t* b2, uint32_t* c2,
                        uint32_t* outputs2, uint32_t* n_slots);

/* ExecuteMode: plain boolean evaluation of the stream (execute_mode.rs). */
int gsvo_execute_stream(const gsvo_stream* s, const uint8_t* input_bits, uint8_t* output_bits_out);

#ifdef __cplusplus
}
#endif
#endif

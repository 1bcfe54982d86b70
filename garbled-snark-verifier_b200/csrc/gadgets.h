// gadgets.h -- C++ restatement of the reference's circuit GENERATOR for the hot-path workloads
// (src/gadgets/basic.rs, src/gadgets/bigint/*, src/gadgets/bn254/{fp254impl,fq,fq2,fq6,fq12}.rs).
// The reference's toolchain (Rust) is absent from this image, so the gate stream -- which is
// pure program order of `add_gate` calls -- is reproduced here statement by statement.  Gate
// ORDER and wire ISSUE order are part of the contract: gid (the hash tweak) is the running
// gate index and credits are popped in issue order.
#pragma once
#include <string>

#include "bigconst.h"
#include "bn254_host.h"
#include "circuit.h"

namespace gsv {

using BigInt = Wires;  // BigIntWires: LSB-first bit wires (bigint/mod.rs:50-53)

// ---- src/gadgets/basic.rs ------------------------------------------------------------------
struct Pair2 { Wire first, second; };
Pair2 half_adder(Builder& c, Wire a, Wire b);
Pair2 full_adder(Builder& c, Wire a, Wire b, Wire cin);
Pair2 half_subtracter(Builder& c, Wire a, Wire b);
Pair2 full_subtracter(Builder& c, Wire a, Wire b, Wire cin);
Wire selector(Builder& c, Wire a, Wire b, Wire s);

// ---- src/gadgets/bigint ---------------------------------------------------------------------
BigInt bn_constant(size_t len, const U256& u);  // BigIntWires::new_constant
BigInt bn_add(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_add_without_carry(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_add_constant(Builder& c, const BigInt& a, const U256& k);
BigInt bn_add_constant_without_carry(Builder& c, const BigInt& a, const U256& k);
BigInt bn_sub(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_sub_without_borrow(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_half(const BigInt& a);
BigInt bn_self_or_zero(Builder& c, const BigInt& a, Wire s);
Wire bn_equal_constant(Builder& c, const BigInt& a, const U256& k);
Wire bn_equal_zero(Builder& c, const BigInt& a);
Wire bn_greater_than(Builder& c, const BigInt& a, const BigInt& b);
Wire bn_less_than_constant(Builder& c, const BigInt& a, const U256& k);
BigInt bn_select(Builder& c, const BigInt& a, const BigInt& b, Wire s);
BigInt bn_mul_naive(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_mul_karatsuba(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_mul(Builder& c, const BigInt& a, const BigInt& b);
BigInt bn_mul_by_constant(Builder& c, const BigInt& a, const U256& k);
BigInt bn_mul_by_constant_modulo_power_two(Builder& c, const BigInt& a, const U256& k, size_t power);

// ---- src/gadgets/bn254/fp254impl.rs + fq.rs (Fq only) -----------------------------------------
struct FqConsts {
  U256 p, r, m_inv, r_inv, not_mod, half_mod, third, two_third;
  static const FqConsts& get();
};
constexpr size_t FQ_BITS = 254;
using Fq = BigInt;
Fq fq_add(Builder& c, const Fq& a, const Fq& b);
Fq fq_add_constant(Builder& c, const Fq& a, const U256& k);  // k: standard-form integer
Fq fq_sub(Builder& c, const Fq& a, const Fq& b);
Fq fq_neg(Builder& c, const Fq& a);
Fq fq_double(Builder& c, const Fq& a);
Fq fq_half(Builder& c, const Fq& a);
Fq fq_triple(Builder& c, const Fq& a);
Fq fq_div6(Builder& c, const Fq& a);
Fq fq_mul_montgomery(Builder& c, const Fq& a, const Fq& b);
Fq fq_square_montgomery(Builder& c, const Fq& a);
Fq fq_montgomery_reduce(Builder& c, const BigInt& x);

// ---- fq2.rs / fq6.rs / fq12.rs ------------------------------------------------------------------
struct Fq2 { Fq c0, c1; };
struct Fq6 { Fq2 c0, c1, c2; };
struct Fq12 { Fq6 c0, c1; };
Wires to_wires(const Fq2& a);
Wires to_wires(const Fq6& a);
Wires to_wires(const Fq12& a);
Fq2 fq2_from_wires(const Wire* w);
Fq6 fq6_from_wires(const Wire* w);
Fq12 fq12_from_wires(const Wire* w);

Fq2 fq2_add(Builder& c, const Fq2& a, const Fq2& b);
Fq2 fq2_sub(Builder& c, const Fq2& a, const Fq2& b);
Fq2 fq2_double(Builder& c, const Fq2& a);
Fq2 fq2_triple(Builder& c, const Fq2& a);
Fq2 fq2_div6(Builder& c, const Fq2& a);
Fq2 fq2_mul_montgomery(Builder& c, const Fq2& a, const Fq2& b);
Fq2 fq2_mul_by_nonresidue(Builder& c, const Fq2& a);

Fq6 fq6_add(Builder& c, const Fq6& a, const Fq6& b);
Fq6 fq6_sub(Builder& c, const Fq6& a, const Fq6& b);
Fq6 fq6_double(Builder& c, const Fq6& a);
Fq6 fq6_div6(Builder& c, const Fq6& a);
Fq6 fq6_mul_montgomery(Builder& c, const Fq6& a, const Fq6& b);
Fq6 fq6_mul_by_nonresidue(Builder& c, const Fq6& a);

Fq12 fq12_mul_montgomery(Builder& c, const Fq12& a, const Fq12& b);

// ---- gadgets_bn254.cpp: pairing / Groth16 verifier (src/gadgets/bn254/*, src/gadgets/groth16.rs) -----
struct G1P { Fq x, y, z; };   // G1Projective wires (Montgomery Jacobian coordinates)
struct G2P { Fq2 x, y, z; };  // G2Projective wires
U256 mont254(const U256& x);  // x * 2^254 mod p
Fq fq_constant(const U256& v);
Wires to_wires(const G1P& p);
Wires to_wires(const G2P& p);
G1P g1_from_wires(const Wire* w);
G2P g2_from_wires(const Wire* w);
Fq fq_inverse(Builder& c, const Fq& a);
Fq fq_inverse_montgomery(Builder& c, const Fq& a);
Fq fq_mul_by_constant_montgomery(Builder& c, const Fq& a, const U256& b);
Fq fq_exp_by_constant_montgomery(Builder& c, const Fq& a, const U256& exp);
Fq fq_sqrt_montgomery(Builder& c, const Fq& a);
Wire fq_is_qnr_montgomery(Builder& c, const Fq& x);
G1P g1_add_montgomery(Builder& c, const G1P& p, const G1P& q);
Wire groth16_verify(Builder& c, const std::vector<Wires>& publics, const G1P& a, const G2P& b, const G1P& cc,
                    const host::VerifyingKey& vk);
Wire groth16_verify_compressed(Builder& c, const Wires& in, size_t n_public, const host::VerifyingKey& vk);
uint32_t build_groth16_verify_compressed(Builder& b, const host::VerifyingKey& vk, size_t n_public);
// groth16_verify on uncompressed points (src/gadgets/groth16.rs:57-110; inputs per src/garbled_groth16.rs:141-176)
uint32_t build_groth16_verify(Builder& b, const host::VerifyingKey& vk, size_t n_public);
// dispatch over the circuit names documented at gsv_program_build (include/gsv_cuda.h)
uint32_t build_named_circuit(Builder& b, const std::string& name);
uint32_t build_fq_inverse(Builder& b);
uint32_t build_fq_sqrt(Builder& b);
uint32_t build_fq2_sqrt(Builder& b);
uint32_t build_g1_add(Builder& b);
uint32_t build_g1_msm1(Builder& b, const host::G1Affine& base);
uint32_t build_fq12_square(Builder& b);
uint32_t build_fq12_cyclotomic_square(Builder& b);
uint32_t build_fq12_inverse(Builder& b);
uint32_t build_fq12_frobenius(Builder& b, size_t i);
uint32_t build_g2_double_step(Builder& b);
uint32_t build_g2_add_step(Builder& b);
uint32_t build_g2_mul_by_char(Builder& b);
uint32_t build_ell(Builder& b);
uint32_t build_ell_const(Builder& b);
uint32_t build_g1_to_affine(Builder& b);
uint32_t build_decompress_g1(Builder& b);
uint32_t build_final_exponentiation(Builder& b);
uint32_t build_miller_loop_groth16(Builder& b, const host::G2Affine& q1, const host::G2Affine& q2);

// ---- named workload circuits (root closures) ---------------------------------------------------
// Returns the root template index.  Input order = the reference's EncodeInput order.
uint32_t build_fq12_mul(Builder& b);   // tests/fq12_mul_e2e.rs:168-174 (6096 inputs, 3048 outputs)
uint32_t build_fq_mul(Builder& b);     // Fq::mul_montgomery(a, b) (508 inputs, 254 outputs)
uint32_t build_fq_add(Builder& b);     // Fq::add(a, b)
uint32_t build_fq2_mul(Builder& b);    // Fq2::mul_montgomery
uint32_t build_fq6_mul(Builder& b);    // Fq6::mul_montgomery
uint32_t build_bn_mul(Builder& b, size_t n_bits);  // bigint::mul on n-bit operands
// every gate type once + one dead gate (tests/streaming_evaluate.rs:67-134 shape)
uint32_t build_gate_zoo(Builder& b);
// Fq "((a^2) * b) + a" (tests/streaming_evaluate.rs Fq case)
uint32_t build_fq_expr(Builder& b);

}  // namespace gsv

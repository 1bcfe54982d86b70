// bn254_host.h -- off-circuit BN254 arithmetic for the constants the verifier topology bakes in.
//
// The reference takes these from arkworks 0.5 at run time (ark-bn254 / ark-ec / ark-ff, absent from
// /root/reference): MSM window tables in arkworks' un-normalised Jacobian coordinates
// (src/gadgets/bn254/g1.rs:318-356), constant-Q line coefficients (pairing.rs:30-132),
// alpha_beta = FE(ML(alpha, -beta))^-1 (groth16.rs:98-105), Frobenius / twist coefficients.
// Every value here is a canonical field element of a mathematically determined quantity, so any
// correct implementation reproduces arkworks' bits; the point formulas follow arkworks' (EFD
// add-2007-bl / dbl-2009-l, and the homogeneous-projective G2 steps spelled out in pairing.rs).
// Self-checked in tests (bilinearity, curve membership, Frobenius^12 = id).
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "bigconst.h"

namespace gsv {
namespace host {

// ---- Fp: BN254 base field, internal Montgomery form w.r.t. 2^256 (CIOS multiplication)
struct FpCtx {
  U256 p, r2;  // modulus, (2^256)^2 mod p
  uint64_t inv;  // -p^-1 mod 2^64
  static const FpCtx& get();
};

struct Fp {
  U256 m;  // Montgomery representation
  Fp() {}
  static Fp from_u256(const U256& v);
  static Fp from_u64(uint64_t v) { return from_u256(U256(v)); }
  static Fp from_dec(const std::string& s) { return from_u256(U256::from_dec(s)); }
  U256 to_u256() const;
  bool is_zero() const { return m.is_zero(); }
  bool operator==(const Fp& o) const { return m == o.m; }
  bool operator!=(const Fp& o) const { return !(m == o.m); }
};
Fp operator+(const Fp& a, const Fp& b);
Fp operator-(const Fp& a, const Fp& b);
Fp operator-(const Fp& a);
Fp operator*(const Fp& a, const Fp& b);
Fp fp_pow(const Fp& a, const U256& e);
Fp fp_inv(const Fp& a);
inline Fp fp_dbl(const Fp& a) { return a + a; }

// ---- tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - xi), xi = 9+u, Fp12 = Fp6[w]/(w^2 - v)
struct Fp2 {
  Fp c0, c1;
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
};
Fp2 operator+(const Fp2& a, const Fp2& b);
Fp2 operator-(const Fp2& a, const Fp2& b);
Fp2 operator-(const Fp2& a);
Fp2 operator*(const Fp2& a, const Fp2& b);
Fp2 fp2_scale(const Fp2& a, const Fp& k);
Fp2 fp2_conj(const Fp2& a);
Fp2 fp2_inv(const Fp2& a);
Fp2 fp2_pow(const Fp2& a, const U256& e);
Fp2 fp2_mul_xi(const Fp2& a);
inline Fp2 fp2_sq(const Fp2& a) { return a * a; }
inline Fp2 fp2_dbl(const Fp2& a) { return a + a; }
Fp2 fp2_one();
Fp2 fp2_xi();

struct Fp6 {
  Fp2 c0, c1, c2;
  bool operator==(const Fp6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
};
Fp6 operator+(const Fp6& a, const Fp6& b);
Fp6 operator-(const Fp6& a, const Fp6& b);
Fp6 operator-(const Fp6& a);
Fp6 operator*(const Fp6& a, const Fp6& b);
Fp6 fp6_mul_by_v(const Fp6& a);
Fp6 fp6_inv(const Fp6& a);
Fp6 fp6_frobenius(const Fp6& a, unsigned i);

struct Fp12 {
  Fp6 c0, c1;
  bool operator==(const Fp12& o) const { return c0 == o.c0 && c1 == o.c1; }
};
Fp12 operator*(const Fp12& a, const Fp12& b);
Fp12 fp12_one();
Fp12 fp12_inv(const Fp12& a);
Fp12 fp12_conj(const Fp12& a);
Fp12 fp12_frobenius(const Fp12& a, unsigned i);
inline Fp12 fp12_sq(const Fp12& a) { return a * a; }
// f * (c0 + c3 w^3 + c4 w^4) in arkworks' 034 indexing: Fp12 = (c0, 0, 0) + (c3, c4, 0) w
Fp12 fp12_mul_by_034(const Fp12& f, const Fp2& c0, const Fp2& c3, const Fp2& c4);

// ---- curve / pairing parameters (ark-bn254)
struct Params {
  U256 p, r;                       // base / scalar field moduli
  uint64_t x;                      // BN parameter, positive for BN254
  std::vector<int8_t> ate_loop;    // signed digits of 6x+2, little-endian (ark Config::ATE_LOOP_COUNT)
  std::vector<int8_t> x_naf;       // find_naf(X), little-endian
  Fp g1_b;                         // 3
  Fp2 g2_b;                        // 3 / xi
  Fp2 twist_mul_by_q_x, twist_mul_by_q_y;
  Fp2 frob_fp2_c1[2];
  Fp2 frob_fp6_c1[6], frob_fp6_c2[6];
  Fp2 frob_fp12_c1[12];
  static const Params& get();
};

// ---- G1: arkworks short-Weierstrass Jacobian (x, y, z), zero = (1, 1, 0)
struct G1Affine { Fp x, y; bool inf = false; };
struct G1Jac { Fp x, y, z; };
G1Jac g1_zero();
G1Jac g1_from_affine(const G1Affine& a);
G1Affine g1_to_affine(const G1Jac& p);
G1Jac g1_double(const G1Jac& p);                 // dbl-2009-l (a = 0)
G1Jac g1_add(const G1Jac& p, const G1Jac& q);    // `+=` of arkworks: add-2007-bl with the zero / equal cases
G1Jac g1_mul(const G1Jac& p, const U256& k);
G1Affine g1_generator();
bool g1_on_curve(const G1Affine& a);

// ---- G2 (affine over Fp2) and the line-coefficient precomputation of pairing.rs:30-132
struct G2Affine { Fp2 x, y; bool inf = false; };
struct G2Proj { Fp2 x, y, z; };  // homogeneous projective, as used by arkworks' G2Prepared
G2Affine g2_generator();
G2Affine g2_neg(const G2Affine& a);
G2Affine g2_mul(const G2Affine& a, const U256& k);  // affine double-and-add (host only)
bool g2_on_curve(const G2Affine& a);
using EllCoeff = Fp6;  // (c0, c1, c2) as in arkworks' (ell_0, ell_vw, ell_vv)
EllCoeff g2_double_in_place(G2Proj& r);
EllCoeff g2_add_in_place(G2Proj& r, const G2Affine& q);
G2Affine g2_mul_by_char(const G2Affine& r);
std::vector<EllCoeff> ell_coeffs(const G2Affine& q);

// Bn254::multi_miller_loop / final_exponentiation (ark-ec bn), pairing = FE(ML)
Fp12 miller_loop(const std::vector<G1Affine>& ps, const std::vector<G2Affine>& qs);
Fp12 final_exponentiation(const Fp12& f);
Fp12 fp12_cyclotomic_exp_x(const Fp12& f);  // f^X

// ---- Groth16 verifying key / proof (ark-groth16 VerifyingKey<Bn254>)
struct VerifyingKey {
  G1Affine alpha_g1;
  G2Affine beta_g2, gamma_g2, delta_g2;
  std::vector<G1Affine> gamma_abc_g1;
};
struct Proof { G1Affine a; G2Affine b; G1Affine c; };
// Deterministic synthetic key/proof from explicit scalars: the pairing
// equation holds by construction for public input x.
void synthetic_groth16(uint64_t seed, const U256& public_x, VerifyingKey& vk, Proof& proof);
bool groth16_verify_host(const VerifyingKey& vk, const Proof& pr, const std::vector<U256>& publics);

}  // namespace host
}  // namespace gsv

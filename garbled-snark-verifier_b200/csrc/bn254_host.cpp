// bn254_host.cpp -- see bn254_host.h.
#include "bn254_host.h"

#include <stdexcept>

namespace gsv {
namespace host {

// =============================================================================== Fp
const FpCtx& FpCtx::get() {
  static const FpCtx c = [] {
    FpCtx k;
    k.p = U256::from_dec("21888242871839275222246405745257275088696311157297823662689037894645226208583");
    // -p^-1 mod 2^64 by Newton iteration
    uint64_t p0 = k.p.l[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
    k.inv = ~x + 1;
    // (2^256)^2 mod p: 512 modular doublings of 1
    U256 r(1);
    for (int i = 0; i < 512; i++) r = addmod(r, r, k.p);
    k.r2 = r;
    return k;
  }();
  return c;
}

static U256 mont_mul(const U256& a, const U256& b) {
  const FpCtx& C = FpCtx::get();
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    unsigned __int128 carry = 0;
    for (int j = 0; j < 4; j++) {
      unsigned __int128 v = (unsigned __int128)a.l[j] * b.l[i] + t[j] + carry;
      t[j] = (uint64_t)v;
      carry = v >> 64;
    }
    unsigned __int128 v = (unsigned __int128)t[4] + carry;
    t[4] = (uint64_t)v;
    t[5] = (uint64_t)(v >> 64);
    uint64_t m = t[0] * C.inv;
    carry = ((unsigned __int128)m * C.p.l[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      unsigned __int128 w = (unsigned __int128)m * C.p.l[j] + t[j] + carry;
      t[j - 1] = (uint64_t)w;
      carry = w >> 64;
    }
    v = (unsigned __int128)t[4] + carry;
    t[3] = (uint64_t)v;
    t[4] = t[5] + (uint64_t)(v >> 64);
    t[5] = 0;
  }
  U256 r;
  for (int i = 0; i < 4; i++) r.l[i] = t[i];
  if (t[4] || !(r < C.p)) r = sub(r, C.p);
  return r;
}

Fp Fp::from_u256(const U256& v) {
  const FpCtx& C = FpCtx::get();
  U256 x = v;
  while (!(x < C.p)) x = sub(x, C.p);
  Fp f;
  f.m = mont_mul(x, C.r2);
  return f;
}
U256 Fp::to_u256() const { return mont_mul(m, U256(1)); }
Fp operator+(const Fp& a, const Fp& b) { Fp r; r.m = addmod(a.m, b.m, FpCtx::get().p); return r; }
Fp operator-(const Fp& a, const Fp& b) { Fp r; r.m = submod(a.m, b.m, FpCtx::get().p); return r; }
Fp operator-(const Fp& a) { Fp z; return z - a; }
Fp operator*(const Fp& a, const Fp& b) { Fp r; r.m = mont_mul(a.m, b.m); return r; }
Fp fp_pow(const Fp& a, const U256& e) {
  Fp r = Fp::from_u64(1), x = a;
  unsigned n = e.bits();
  for (unsigned i = 0; i < n; i++) {
    if (e.bit(i)) r = r * x;
    x = x * x;
  }
  return r;
}
Fp fp_inv(const Fp& a) {
  if (a.is_zero()) throw std::domain_error("inverse of zero in Fp");
  return fp_pow(a, sub(FpCtx::get().p, U256(2)));
}

// =============================================================================== Fp2
Fp2 operator+(const Fp2& a, const Fp2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
Fp2 operator-(const Fp2& a, const Fp2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
Fp2 operator-(const Fp2& a) { return {-a.c0, -a.c1}; }
Fp2 operator*(const Fp2& a, const Fp2& b) {
  return {a.c0 * b.c0 - a.c1 * b.c1, a.c0 * b.c1 + a.c1 * b.c0};
}
Fp2 fp2_scale(const Fp2& a, const Fp& k) { return {a.c0 * k, a.c1 * k}; }
Fp2 fp2_conj(const Fp2& a) { return {a.c0, -a.c1}; }
Fp2 fp2_inv(const Fp2& a) {
  Fp n = fp_inv(a.c0 * a.c0 + a.c1 * a.c1);
  return {a.c0 * n, -(a.c1 * n)};
}
Fp2 fp2_one() { return {Fp::from_u64(1), Fp()}; }
Fp2 fp2_xi() { return {Fp::from_u64(9), Fp::from_u64(1)}; }
Fp2 fp2_mul_xi(const Fp2& a) { return a * fp2_xi(); }
Fp2 fp2_pow(const Fp2& a, const U256& e) {
  Fp2 r = fp2_one(), x = a;
  unsigned n = e.bits();
  for (unsigned i = 0; i < n; i++) {
    if (e.bit(i)) r = r * x;
    x = x * x;
  }
  return r;
}

// =============================================================================== Fp6 / Fp12
Fp6 operator+(const Fp6& a, const Fp6& b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
Fp6 operator-(const Fp6& a, const Fp6& b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
Fp6 operator-(const Fp6& a) { return {-a.c0, -a.c1, -a.c2}; }
Fp6 operator*(const Fp6& a, const Fp6& b) {
  Fp2 c0 = a.c0 * b.c0 + fp2_mul_xi(a.c1 * b.c2 + a.c2 * b.c1);
  Fp2 c1 = a.c0 * b.c1 + a.c1 * b.c0 + fp2_mul_xi(a.c2 * b.c2);
  Fp2 c2 = a.c0 * b.c2 + a.c1 * b.c1 + a.c2 * b.c0;
  return {c0, c1, c2};
}
Fp6 fp6_mul_by_v(const Fp6& a) { return {fp2_mul_xi(a.c2), a.c0, a.c1}; }
Fp6 fp6_inv(const Fp6& a) {
  Fp2 t0 = a.c0 * a.c0 - fp2_mul_xi(a.c1 * a.c2);
  Fp2 t1 = fp2_mul_xi(a.c2 * a.c2) - a.c0 * a.c1;
  Fp2 t2 = a.c1 * a.c1 - a.c0 * a.c2;
  Fp2 n = fp2_inv(a.c0 * t0 + fp2_mul_xi(a.c2 * t1 + a.c1 * t2));
  return {t0 * n, t1 * n, t2 * n};
}
Fp6 fp6_frobenius(const Fp6& a, unsigned i) {
  const Params& P = Params::get();
  auto fr = [&](const Fp2& x) { return (i & 1) ? fp2_conj(x) : x; };
  return {fr(a.c0), fr(a.c1) * P.frob_fp6_c1[i % 6], fr(a.c2) * P.frob_fp6_c2[i % 6]};
}

Fp12 operator*(const Fp12& a, const Fp12& b) {
  Fp6 a0b0 = a.c0 * b.c0, a1b1 = a.c1 * b.c1;
  return {a0b0 + fp6_mul_by_v(a1b1), a.c0 * b.c1 + a.c1 * b.c0};
}
Fp12 fp12_one() {
  Fp12 r;
  r.c0.c0 = fp2_one();
  return r;
}
Fp12 fp12_inv(const Fp12& a) {
  Fp6 n = fp6_inv(a.c0 * a.c0 - fp6_mul_by_v(a.c1 * a.c1));
  return {a.c0 * n, -(a.c1 * n)};
}
Fp12 fp12_conj(const Fp12& a) { return {a.c0, -a.c1}; }
Fp12 fp12_frobenius(const Fp12& a, unsigned i) {
  const Params& P = Params::get();
  Fp6 c0 = fp6_frobenius(a.c0, i), c1 = fp6_frobenius(a.c1, i);
  const Fp2& k = P.frob_fp12_c1[i % 12];
  return {c0, Fp6{c1.c0 * k, c1.c1 * k, c1.c2 * k}};
}
Fp12 fp12_mul_by_034(const Fp12& f, const Fp2& c0, const Fp2& c3, const Fp2& c4) {
  Fp12 o;
  o.c0.c0 = c0;
  o.c1.c0 = c3;
  o.c1.c1 = c4;
  return f * o;
}

// =============================================================================== parameters
static U256 div_small(const U256& a, uint64_t d) {
  U256 q;
  unsigned __int128 rem = 0;
  for (int i = 3; i >= 0; i--) {
    unsigned __int128 cur = (rem << 64) | a.l[i];
    q.l[i] = (uint64_t)(cur / d);
    rem = cur % d;
  }
  return q;
}

const Params& Params::get() {
  static const Params P = [] {
    Params k;
    k.p = FpCtx::get().p;
    k.r = U256::from_dec("21888242871839275222246405745257275088548364400416034343698204186575808495617");
    k.x = 4965661367192848881ull;
    static const int8_t ATE[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0,
                                   0, 1, 0, -1, 0, 0, 0, 0, 1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0,
                                   -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1};
    k.ate_loop.assign(ATE, ATE + 65);
    {
      // self-check: sum d_i 2^i == 6x + 2
      __int128 s = 0;
      for (int i = 64; i >= 0; i--) s = s * 2 + ATE[i];
      if (s != (__int128)6 * k.x + 2) throw std::logic_error("ATE_LOOP_COUNT does not encode 6x+2");
    }
    {
      // ark_ff::biginteger::arithmetic::find_naf(X)
      unsigned __int128 num = k.x;
      while (num != 0) {
        int8_t z = 0;
        if (num & 1) {
          z = (int8_t)(2 - (int)(num % 4));
          if (z > 0) num -= 1; else num += 1;
        }
        k.x_naf.push_back(z);
        num >>= 1;
      }
    }
    k.g1_b = Fp::from_u64(3);
    k.g2_b = fp2_scale(fp2_inv(fp2_xi()), Fp::from_u64(3));
    const U256 pm1 = sub(k.p, U256(1));
    const Fp2 xi = fp2_xi();
    const Fp2 g6 = fp2_pow(xi, div_small(pm1, 6));  // xi^((p-1)/6)
    const Fp2 g3 = g6 * g6;                          // xi^((p-1)/3)
    const Fp2 g2 = g3 * g6;                          // xi^((p-1)/2)
    k.twist_mul_by_q_x = g3;
    k.twist_mul_by_q_y = g2;
    k.frob_fp2_c1[0] = fp2_one();
    k.frob_fp2_c1[1] = -fp2_one();
    // xi^((p^i - 1)/d) = conj(xi^((p^(i-1) - 1)/d)) * xi^((p-1)/d)
    k.frob_fp6_c1[0] = fp2_one();
    k.frob_fp12_c1[0] = fp2_one();
    for (int i = 1; i < 6; i++) k.frob_fp6_c1[i] = fp2_conj(k.frob_fp6_c1[i - 1]) * g3;
    for (int i = 0; i < 6; i++) k.frob_fp6_c2[i] = k.frob_fp6_c1[i] * k.frob_fp6_c1[i];
    for (int i = 1; i < 12; i++) k.frob_fp12_c1[i] = fp2_conj(k.frob_fp12_c1[i - 1]) * g6;
    return k;
  }();
  return P;
}

// =============================================================================== G1
G1Jac g1_zero() { return {Fp::from_u64(1), Fp::from_u64(1), Fp()}; }
G1Jac g1_from_affine(const G1Affine& a) {
  if (a.inf) return g1_zero();
  return {a.x, a.y, Fp::from_u64(1)};
}
G1Affine g1_to_affine(const G1Jac& p) {
  if (p.z.is_zero()) return {Fp(), Fp(), true};
  Fp zi = fp_inv(p.z), zi2 = zi * zi;
  return {p.x * zi2, p.y * zi2 * zi, false};
}
G1Jac g1_double(const G1Jac& p) {
  if (p.z.is_zero()) return p;
  // dbl-2009-l
  Fp a = p.x * p.x, b = p.y * p.y, c = b * b;
  Fp xb = p.x + b;
  Fp d = fp_dbl(xb * xb - a - c);
  Fp e = a + a + a, f = e * e;
  Fp z3 = fp_dbl(p.y * p.z);
  Fp x3 = f - fp_dbl(d);
  Fp c8 = fp_dbl(fp_dbl(fp_dbl(c)));
  Fp y3 = e * (d - x3) - c8;
  return {x3, y3, z3};
}
G1Jac g1_add(const G1Jac& p, const G1Jac& q) {
  if (p.z.is_zero()) return q;
  if (q.z.is_zero()) return p;
  // add-2007-bl
  Fp z1z1 = p.z * p.z, z2z2 = q.z * q.z;
  Fp u1 = p.x * z2z2, u2 = q.x * z1z1;
  Fp s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2 && s1 == s2) return g1_double(p);
  Fp h = u2 - u1;
  Fp i = fp_dbl(h);
  i = i * i;
  Fp j = h * i;
  Fp r = fp_dbl(s2 - s1);
  Fp v = u1 * i;
  Fp x3 = r * r - j - fp_dbl(v);
  Fp y3 = r * (v - x3) - fp_dbl(s1 * j);
  Fp z3 = fp_dbl(p.z * q.z) * h;
  return {x3, y3, z3};
}
G1Jac g1_mul(const G1Jac& p, const U256& k) {
  G1Jac acc = g1_zero();
  for (int i = (int)k.bits() - 1; i >= 0; i--) {
    acc = g1_double(acc);
    if (k.bit((unsigned)i)) acc = g1_add(acc, p);
  }
  return acc;
}
G1Affine g1_generator() { return {Fp::from_u64(1), Fp::from_u64(2), false}; }
bool g1_on_curve(const G1Affine& a) {
  if (a.inf) return true;
  return a.y * a.y == a.x * a.x * a.x + Params::get().g1_b;
}

// =============================================================================== G2
G2Affine g2_generator() {
  G2Affine g;
  g.x = {Fp::from_dec("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
         Fp::from_dec("11559732032986387107991004021392285783925812861821192530917403151452391805634")};
  g.y = {Fp::from_dec("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
         Fp::from_dec("4082367875863433681332203403145435568316851327593401208105741076214120093531")};
  return g;
}
G2Affine g2_neg(const G2Affine& a) { return {a.x, -a.y, a.inf}; }
bool g2_on_curve(const G2Affine& a) {
  if (a.inf) return true;
  return a.y * a.y == a.x * a.x * a.x + Params::get().g2_b;
}
static G2Affine g2_add_affine(const G2Affine& p, const G2Affine& q) {
  if (p.inf) return q;
  if (q.inf) return p;
  Fp2 lam;
  if (p.x == q.x) {
    if (!(p.y == q.y) || p.y.is_zero()) return {Fp2(), Fp2(), true};
    Fp2 x2 = p.x * p.x;
    lam = (x2 + x2 + x2) * fp2_inv(p.y + p.y);
  } else {
    lam = (q.y - p.y) * fp2_inv(q.x - p.x);
  }
  Fp2 x3 = lam * lam - p.x - q.x;
  Fp2 y3 = lam * (p.x - x3) - p.y;
  return {x3, y3, false};
}
G2Affine g2_mul(const G2Affine& a, const U256& k) {
  G2Affine acc{Fp2(), Fp2(), true};
  for (int i = (int)k.bits() - 1; i >= 0; i--) {
    acc = g2_add_affine(acc, acc);
    if (k.bit((unsigned)i)) acc = g2_add_affine(acc, a);
  }
  return acc;
}

// pairing.rs:30-52 (arkworks bn::g2::doubling_step)
EllCoeff g2_double_in_place(G2Proj& r) {
  const Params& P = Params::get();
  const Fp half = fp_inv(Fp::from_u64(2));
  Fp2 a = fp2_scale(r.x * r.y, half);
  Fp2 b = r.y * r.y;
  Fp2 c = r.z * r.z;
  Fp2 e = P.g2_b * (fp2_dbl(c) + c);
  Fp2 f = fp2_dbl(e) + e;
  Fp2 g = fp2_scale(b + f, half);
  Fp2 yz = r.y + r.z;
  Fp2 h = yz * yz - (b + c);
  Fp2 i = e - b;
  Fp2 j = r.x * r.x;
  Fp2 e2 = e * e;
  G2Proj n{a * (b - f), g * g - (fp2_dbl(e2) + e2), b * h};
  r = n;
  return {-h, fp2_dbl(j) + j, i};
}
// pairing.rs:54-75 (arkworks bn::g2::addition_step)
EllCoeff g2_add_in_place(G2Proj& r, const G2Affine& q) {
  Fp2 theta = r.y - q.y * r.z;
  Fp2 lambda = r.x - q.x * r.z;
  Fp2 c = theta * theta;
  Fp2 d = lambda * lambda;
  Fp2 e = lambda * d;
  Fp2 f = r.z * c;
  Fp2 g = r.x * d;
  Fp2 h = e + f - fp2_dbl(g);
  Fp2 j = theta * q.x - lambda * q.y;
  G2Proj n{lambda * h, theta * (g - h) - e * r.y, r.z * e};
  r = n;
  return {lambda, -theta, j};
}
G2Affine g2_mul_by_char(const G2Affine& r) {
  const Params& P = Params::get();
  return {fp2_conj(r.x) * P.twist_mul_by_q_x, fp2_conj(r.y) * P.twist_mul_by_q_y, r.inf};
}
std::vector<EllCoeff> ell_coeffs(const G2Affine& q) {
  const Params& P = Params::get();
  std::vector<EllCoeff> ellc;
  G2Proj r{q.x, q.y, fp2_one()};
  G2Affine neg_q = g2_neg(q);
  for (int i = (int)P.ate_loop.size() - 2; i >= 0; i--) {  // .iter().rev().skip(1)
    ellc.push_back(g2_double_in_place(r));
    if (P.ate_loop[i] == 1) ellc.push_back(g2_add_in_place(r, q));
    else if (P.ate_loop[i] == -1) ellc.push_back(g2_add_in_place(r, neg_q));
  }
  G2Affine q1 = g2_mul_by_char(q);
  G2Affine q2 = g2_mul_by_char(q1);
  q2.y = -q2.y;
  ellc.push_back(g2_add_in_place(r, q1));
  ellc.push_back(g2_add_in_place(r, q2));
  return ellc;
}

// ark-ec bn::Bn::multi_miller_loop with the D-type twist line evaluation (ell)
Fp12 miller_loop(const std::vector<G1Affine>& ps, const std::vector<G2Affine>& qs) {
  const Params& P = Params::get();
  std::vector<std::vector<EllCoeff>> cs;
  for (const G2Affine& q : qs) cs.push_back(ell_coeffs(q));
  std::vector<size_t> it(ps.size(), 0);
  Fp12 f = fp12_one();
  auto ell_all = [&]() {
    for (size_t k = 0; k < ps.size(); k++) {
      const EllCoeff& c = cs[k][it[k]++];
      f = fp12_mul_by_034(f, fp2_scale(c.c0, ps[k].y), fp2_scale(c.c1, ps[k].x), c.c2);
    }
  };
  const int n = (int)P.ate_loop.size();
  for (int i = n - 1; i >= 1; i--) {
    if (i != n - 1) f = fp12_sq(f);
    ell_all();
    if (P.ate_loop[i - 1] == 1 || P.ate_loop[i - 1] == -1) ell_all();
  }
  ell_all();
  ell_all();
  return f;
}
Fp12 fp12_cyclotomic_exp_x(const Fp12& f) {
  const Params& P = Params::get();
  Fp12 r = fp12_one();
  for (int i = 63; i >= 0; i--) {
    r = fp12_sq(r);
    if ((P.x >> i) & 1) r = r * f;
  }
  return r;
}
// final_exponentiation.rs:37-64 (== ark-ec bn final exponentiation)
Fp12 final_exponentiation(const Fp12& f) {
  auto exp_by_neg_x = [](const Fp12& v) { return fp12_conj(fp12_cyclotomic_exp_x(v)); };
  Fp12 u = fp12_inv(f) * fp12_conj(f);
  Fp12 r = fp12_frobenius(u, 2) * u;
  Fp12 y0 = exp_by_neg_x(r);
  Fp12 y1 = fp12_sq(y0);
  Fp12 y2 = fp12_sq(y1);
  Fp12 y3 = y2 * y1;
  Fp12 y4 = exp_by_neg_x(y3);
  Fp12 y5 = fp12_sq(y4);
  Fp12 y6 = exp_by_neg_x(y5);
  Fp12 y7 = fp12_conj(y3);
  Fp12 y8 = fp12_conj(y6);
  Fp12 y9 = y8 * y4;
  Fp12 y10 = y9 * y7;
  Fp12 y11 = y10 * y1;
  Fp12 y12 = y10 * y4;
  Fp12 y13 = y12 * r;
  Fp12 y14 = fp12_frobenius(y11, 1);
  Fp12 y15 = y14 * y13;
  Fp12 y16 = fp12_frobenius(y10, 2);
  Fp12 y17 = y16 * y15;
  Fp12 r2 = fp12_conj(r);
  Fp12 y18 = r2 * y11;
  Fp12 y19 = fp12_frobenius(y18, 3);
  return y19 * y17;
}

// =============================================================================== Groth16
static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static U256 rand_scalar(uint64_t& s, const U256& r) {
  U256 v;
  for (int i = 0; i < 4; i++) v.l[i] = splitmix(s);
  v.l[3] &= 0x0FFFFFFFFFFFFFFFull;  // < 2^252 < r
  if (v.is_zero()) v = U256(1);
  (void)r;
  return v;
}

void synthetic_groth16(uint64_t seed, const U256& public_x, VerifyingKey& vk, Proof& proof) {
  const Params& P = Params::get();
  const U256& r = P.r;
  uint64_t s = seed * 0x2545F4914F6CDD1Dull + 0x1234567;
  U256 alpha = rand_scalar(s, r), beta = rand_scalar(s, r), gamma = rand_scalar(s, r), delta = rand_scalar(s, r);
  U256 ic0 = rand_scalar(s, r), ic1 = rand_scalar(s, r), a = rand_scalar(s, r), b = rand_scalar(s, r);
  U256 x = public_x;
  while (!(x < r)) x = sub(x, r);
  // c = (a b - alpha beta - (ic0 + x ic1) gamma) / delta  (mod r)
  U256 msm = addmod(ic0, mulmod(x, ic1, r), r);
  U256 num = submod(submod(mulmod(a, b, r), mulmod(alpha, beta, r), r), mulmod(msm, gamma, r), r);
  U256 c = mulmod(num, invmod(delta, r), r);
  G1Jac g1 = g1_from_affine(g1_generator());
  G2Affine g2 = g2_generator();
  vk.alpha_g1 = g1_to_affine(g1_mul(g1, alpha));
  vk.beta_g2 = g2_mul(g2, beta);
  vk.gamma_g2 = g2_mul(g2, gamma);
  vk.delta_g2 = g2_mul(g2, delta);
  vk.gamma_abc_g1 = {g1_to_affine(g1_mul(g1, ic0)), g1_to_affine(g1_mul(g1, ic1))};
  proof.a = g1_to_affine(g1_mul(g1, a));
  proof.b = g2_mul(g2, b);
  proof.c = g1_to_affine(g1_mul(g1, c));
}

bool groth16_verify_host(const VerifyingKey& vk, const Proof& pr, const std::vector<U256>& publics) {
  G1Jac acc = g1_from_affine(vk.gamma_abc_g1[0]);
  for (size_t i = 0; i < publics.size(); i++)
    acc = g1_add(acc, g1_mul(g1_from_affine(vk.gamma_abc_g1[i + 1]), publics[i]));
  G1Affine msm = g1_to_affine(acc);
  Fp12 f = final_exponentiation(miller_loop({msm, pr.c, pr.a}, {g2_neg(vk.gamma_g2), g2_neg(vk.delta_g2), pr.b}));
  Fp12 alpha_beta = fp12_inv(final_exponentiation(miller_loop({vk.alpha_g1}, {g2_neg(vk.beta_g2)})));
  return f == alpha_beta;
}

}  // namespace host
}  // namespace gsv

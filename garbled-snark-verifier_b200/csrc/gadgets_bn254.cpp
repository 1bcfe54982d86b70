// gadgets_bn254.cpp -- emission-order restatement of the reference's pairing / Groth16 gadgets:
// src/gadgets/bn254/{fp254impl,fq,fq2,fq6,fq12,g1,pairing,final_exponentiation}.rs and
// src/gadgets/groth16.rs.  Same rules as gadgets.cpp: statement order == Rust program order,
// every #[component] is a Builder::component with the reference's key fields.
#include <algorithm>
#include <stdexcept>

#include "gadgets.h"

namespace gsv {

using host::Fp;
using host::Fp2;
using host::Fp6;
using host::Fp12;

static inline BigInt slice(const BigInt& v, size_t lo, size_t hi) { return BigInt(v.begin() + lo, v.begin() + hi); }
static inline BigInt concat(const BigInt& a, const BigInt& b) {
  BigInt r(a);
  r.insert(r.end(), b.begin(), b.end());
  return r;
}

// ---- constants: standard integer of a field element / its Montgomery (R = 2^254) image
U256 mont254(const U256& x) {
  const FqConsts& K = FqConsts::get();
  return mulmod(x, K.not_mod /* R mod p = R - p */, K.p);
}
static U256 ival(const Fp& x) { return x.to_u256(); }
static Fp fp_mont(const Fp& x) { return Fp::from_u256(mont254(x.to_u256())); }
static Fp2 fp2_mont(const Fp2& x) { return {fp_mont(x.c0), fp_mont(x.c1)}; }
static Fp6 fp6_mont(const Fp6& x) { return {fp2_mont(x.c0), fp2_mont(x.c1), fp2_mont(x.c2)}; }
static std::string hex2(const Fp2& x) { return ival(x.c0).to_hex() + ival(x.c1).to_hex(); }
static std::string hex6(const Fp6& x) { return hex2(x.c0) + hex2(x.c1) + hex2(x.c2); }

Fq fq_constant(const U256& v) { return bn_constant(FQ_BITS, v); }       // Fq::new_constant
static Fq2 fq2_constant(const Fp2& v) { return {fq_constant(ival(v.c0)), fq_constant(ival(v.c1))}; }
// Fq12::new_constant / new_fq12_constant_montgomery: standard value -> Montgomery constant wires
static Fq12 fq12_constant_montgomery(const Fp12& v) {
  auto c6 = [](const Fp6& x) {
    Fp6 m = fp6_mont(x);
    return Fq6{fq2_constant(m.c0), fq2_constant(m.c1), fq2_constant(m.c2)};
  };
  return {c6(v.c0), c6(v.c1)};
}

// =================================================================== basic / bigint additions
// basic.rs:81-113
Wire basic_multiplexer(Builder& c, const Wires& a, const Wires& s, size_t w) {
  if (a.size() != ((size_t)1 << w) || s.size() != w) throw std::logic_error("multiplexer: bad shape");
  size_t n = a.size();
  return c.component("basic::multiplexer|w=" + std::to_string(w), concat(a, s), 1, [n, w](Builder& c, const Wires& in) {
    Wires cur(in.begin(), in.begin() + n);
    size_t cur_len = n;
    for (size_t k = 0; k < w; k++) {
      Wire sel = in[n + k];
      size_t j = 0;
      for (size_t i = 0; i < cur_len; i += 2, j++) cur[j] = selector(c, cur[i + 1], cur[i], sel);
      cur_len /= 2;
    }
    return Wires{cur[0]};
  })[0];
}
// bigint/cmp.rs:171-193
BigInt bn_multiplexer(Builder& c, const std::vector<BigInt>& a, const Wires& s, size_t w) {
  size_t n = (size_t)1 << w;
  if (a.size() != n) throw std::logic_error("bn_multiplexer: bad shape");
  size_t n_bits = a[0].size();
  Wires in;
  in.reserve(n * n_bits + w);
  for (const BigInt& x : a) in.insert(in.end(), x.begin(), x.end());
  in.insert(in.end(), s.begin(), s.end());
  return c.component("bigint::multiplexer|w=" + std::to_string(w), in, n_bits, [n, n_bits, w](Builder& c, const Wires& in) {
    Wires s(in.begin() + n * n_bits, in.end());
    BigInt bits(n_bits);
    Wires ith(n);
    for (size_t i = 0; i < n_bits; i++) {
      for (size_t k = 0; k < n; k++) ith[k] = in[k * n_bits + i];
      bits[i] = basic_multiplexer(c, ith, s, w);
    }
    return bits;
  });
}
// bigint/add.rs:136-142
BigInt bn_double_without_overflow(Builder& c, const BigInt& a) {
  size_t n = a.size();
  return c.component("bigint::double_without_overflow", a, n, [n](Builder&, const Wires& a) {
    BigInt bits;
    bits.push_back(WIRE_FALSE);
    bits.insert(bits.end(), a.begin(), a.begin() + (n - 1));
    return bits;
  });
}
// bigint/cmp.rs:25-41
BigInt bn_self_or_zero_inv(Builder& c, const BigInt& a, Wire s) {
  size_t n = a.size();
  BigInt in(a);
  in.push_back(s);
  return c.component("bigint::self_or_zero_inv", in, n, [n](Builder& c, const Wires& in) {
    BigInt bits(n);
    for (size_t i = 0; i < n; i++) {
      Wire w = c.issue_wire();
      c.add_gate(NIMP, in[i], in[n], w);  // and_variant [false,true,false]
      bits[i] = w;
    }
    return bits;
  });
}
// bigint/cmp.rs:43-58
Wire bn_equal(Builder& c, const BigInt& a, const BigInt& b) {
  size_t n = a.size();
  return c.component("bigint::equal", concat(a, b), 1, [n](Builder& c, const Wires& in) {
    BigInt x(n);
    for (size_t i = 0; i < n; i++) {
      Wire w = c.issue_wire();
      c.add_gate(XOR, in[i], in[n + i], w);
      x[i] = w;
    }
    return Wires{bn_equal_constant(c, x, U256())};
  })[0];
}
// bigint/add.rs:155-190 (plain function: gates land in the caller's frame)
static void bn_odd_part(Builder& c, const BigInt& a, BigInt& odd_out, BigInt& k_out) {
  size_t n = a.size();
  BigInt select_bn = c.issue_wires(n - 1);
  select_bn.insert(select_bn.begin(), a[0]);
  for (size_t i = 1; i < n; i++) c.add_gate(OR, select_bn[i - 1], a[i], select_bn[i]);
  BigInt k = c.issue_wires(n - 1);
  k.insert(k.begin(), a[0]);
  for (size_t i = 1; i < n; i++) c.add_gate(NCIMP, select_bn[i - 1], a[i], k[i]);
  BigInt odd_acc = a;
  for (size_t i = 0; i < n; i++) {
    BigInt half_res = bn_half(odd_acc);
    odd_acc = bn_select(c, odd_acc, half_res, select_bn[i]);
  }
  odd_out = odd_acc;
  k_out = k;
}

// =================================================================== fp254impl.rs (rest) / fq.rs
// fp254impl.rs:253-270
Fq fq_mul_by_constant_montgomery(Builder& c, const Fq& a, const U256& b) {
  return c.component("fp254::mul_by_constant_montgomery|b=" + b.to_hex(), a, FQ_BITS, [b](Builder& c, const Wires& a) {
    if (b.is_zero()) return bn_constant(a.size(), U256());
    if (b == mont254(U256(1))) return Wires(a);
    BigInt m = bn_mul_by_constant(c, a, b);
    return fq_montgomery_reduce(c, m);
  });
}

// fp254impl.rs:334-661
Fq fq_inverse(Builder& c, const Fq& a) {
  return c.component("fp254::inverse", a, FQ_BITS, [](Builder& c, const Wires& a) {
    const size_t N = FQ_BITS;
    const size_t PER_CHUNK = 4;
    BigInt odd_part, even_part;
    bn_odd_part(c, a, odd_part, even_part);
    Fq neg_odd_part = fq_neg(c, odd_part);
    BigInt u = bn_half(neg_odd_part);
    BigInt v = odd_part;
    BigInt k = bn_constant(N, U256(1));
    BigInt r = bn_constant(N, U256(1));
    BigInt s = bn_constant(N, U256(2));

    auto pack5 = [](const BigInt& u, const BigInt& v, const BigInt& r, const BigInt& s, const BigInt& k) {
      Wires w;
      w.reserve(5 * u.size());
      for (const BigInt* x : {&u, &v, &r, &s, &k}) w.insert(w.end(), x->begin(), x->end());
      return w;
    };
    for (size_t chunk0 = 0; chunk0 < 2 * N; chunk0 += PER_CHUNK) {
      const size_t n_it = std::min(PER_CHUNK, 2 * N - chunk0);
      Wires out = c.component("inverse_iteration", pack5(u, v, r, s, k), 5 * N, [N, n_it, pack5](Builder& c, const Wires& in) {
        BigInt u = slice(in, 0, N), v = slice(in, N, 2 * N), r = slice(in, 2 * N, 3 * N), s = slice(in, 3 * N, 4 * N),
               k = slice(in, 4 * N, 5 * N);
        for (size_t it = 0; it < n_it; it++) {
          Wire not_x1 = u[0];
          Wire not_x2 = v[0];
          Wire x3 = bn_greater_than(c, u, v);
          Wire p2 = c.issue_wire();
          c.add_gate(NIMP, not_x1, not_x2, p2);
          Wire p3 = c.issue_wire();
          Wire wires_2 = c.issue_wire();
          c.add_gate(AND, not_x1, not_x2, wires_2);
          c.add_gate(AND, wires_2, x3, p3);
          Wire p4 = c.issue_wire();
          c.add_gate(NIMP, wires_2, x3, p4);
          // part 1
          BigInt u1 = bn_half(u);
          BigInt v1 = v;
          BigInt r1 = r;
          BigInt s1 = bn_double_without_overflow(c, s);
          BigInt k1 = bn_add_constant_without_carry(c, k, U256(1));
          // part 2
          BigInt u2 = u;
          BigInt v2 = bn_half(v);
          BigInt r2 = bn_double_without_overflow(c, r);
          BigInt s2 = s;
          BigInt k2 = bn_add_constant_without_carry(c, k, U256(1));
          // part 3
          BigInt u3 = bn_sub_without_borrow(c, u1, v2);
          BigInt v3 = v;
          BigInt r3 = bn_add_without_carry(c, r, s);
          BigInt s3 = bn_double_without_overflow(c, s);
          BigInt k3 = bn_add_constant_without_carry(c, k, U256(1));
          // part 4
          BigInt u4 = u;
          BigInt v4 = bn_sub_without_borrow(c, v2, u1);
          BigInt r4 = bn_double_without_overflow(c, r);
          BigInt s4 = bn_add_without_carry(c, r, s);
          BigInt k4 = bn_add_constant_without_carry(c, k, U256(1));

          auto blend = [&](const BigInt& x1, const BigInt& x2, const BigInt& x3v, const BigInt& x4) {
            BigInt w1 = bn_self_or_zero_inv(c, x1, not_x1);
            BigInt w2 = bn_self_or_zero(c, x2, p2);
            BigInt w3 = bn_self_or_zero(c, x3v, p3);
            BigInt w4 = bn_self_or_zero(c, x4, p4);
            BigInt a1 = bn_add_without_carry(c, w1, w2);
            BigInt a2 = bn_add_without_carry(c, a1, w3);
            return bn_add_without_carry(c, a2, w4);
          };
          BigInt new_u = blend(u1, u2, u3, u4);
          BigInt new_v = blend(v1, v2, v3, v4);
          BigInt new_r = blend(r1, r2, r3, r4);
          BigInt new_s = blend(s1, s2, s3, s4);
          BigInt new_k = blend(k1, k2, k3, k4);

          Wire v_equals_one = bn_equal_constant(c, v, U256(1));
          u = bn_select(c, u, new_u, v_equals_one);
          v = bn_select(c, v, new_v, v_equals_one);
          r = bn_select(c, r, new_r, v_equals_one);
          s = bn_select(c, s, new_s, v_equals_one);
          k = bn_select(c, k, new_k, v_equals_one);
        }
        return pack5(u, v, r, s, k);
      });
      u = slice(out, 0, N); v = slice(out, N, 2 * N); r = slice(out, 2 * N, 3 * N); s = slice(out, 3 * N, 4 * N);
      k = slice(out, 4 * N, 5 * N);
    }

    // inverse::divide_result_by_even_part
    BigInt s_div = c.component("inverse::divide_result_by_even_part", concat(s, even_part), N, [N, PER_CHUNK](Builder& c, const Wires& in) {
      BigInt s = slice(in, 0, N), even_part = slice(in, N, 2 * N);
      size_t chunk_idx = 0;
      for (size_t chunk0 = 0; chunk0 < N; chunk0 += PER_CHUNK, chunk_idx++) {
        const size_t n_it = std::min(PER_CHUNK, N - chunk0);
        Wires out = c.component("inverse::divide_result_by_even_part::chunk|chunk_idx=" + std::to_string(chunk_idx),
                                concat(s, even_part), 2 * N, [N, n_it](Builder& c, const Wires& in) {
          BigInt s = slice(in, 0, N), even_part = slice(in, N, 2 * N);
          for (size_t it = 0; it < n_it; it++) {
            BigInt updated_s = fq_half(c, s);
            BigInt updated_even_part = fq_half(c, even_part);
            Wire selector = bn_equal_constant(c, even_part, U256(1));
            s = bn_select(c, s, updated_s, selector);
            even_part = bn_select(c, even_part, updated_even_part, selector);
          }
          return concat(s, even_part);
        });
        s = slice(out, 0, N);
        even_part = slice(out, N, 2 * N);
      }
      return s;
    });

    // inverse::divide_result_by_2^k
    return c.component("inverse::divide_result_by_2^k", concat(s_div, k), N, [N, PER_CHUNK](Builder& c, const Wires& in) {
      const FqConsts& K = FqConsts::get();
      BigInt s = slice(in, 0, N), k = slice(in, N, 2 * N);
      for (size_t chunk0 = 0; chunk0 < 2 * N; chunk0 += PER_CHUNK) {
        const size_t n_it = std::min(PER_CHUNK, 2 * N - chunk0);
        Wires out = c.component("inverse::divide_result_by_2^k::chunk", concat(s, k), 2 * N, [N, n_it, &K](Builder& c, const Wires& in) {
          BigInt s = slice(in, 0, N), k = slice(in, N, 2 * N);
          for (size_t it = 0; it < n_it; it++) {
            BigInt updated_s = fq_half(c, s);
            BigInt updated_k = fq_add_constant(c, k, sub(K.p, U256(1)));  // Fq::from(-1)
            Wire selector = bn_equal_constant(c, k, U256());              // Self::equal_constant(k, ZERO)
            s = bn_select(c, s, updated_s, selector);
            k = bn_select(c, k, updated_k, selector);
          }
          return concat(s, k);
        });
        s = slice(out, 0, N);
        k = slice(out, N, 2 * N);
      }
      return s;
    });
  });
}
// fp254impl.rs:680-690
Fq fq_inverse_montgomery(Builder& c, const Fq& a) {
  const FqConsts& K = FqConsts::get();
  Fq b = fq_inverse(c, a);
  U256 r1 = K.not_mod;  // R mod p
  U256 r3 = mulmod(mulmod(r1, r1, K.p), r1, K.p);
  return fq_mul_by_constant_montgomery(c, b, r3);
}
// fp254impl.rs:692-725
Fq fq_exp_by_constant_montgomery(Builder& c, const Fq& a, const U256& exp) {
  return c.component("fp254::exp_by_constant_montgomery|exp=" + exp.to_hex(), a, FQ_BITS, [exp](Builder& c, const Wires& a) {
    if (exp.is_zero()) return bn_constant(a.size(), U256(1));
    if (exp == U256(1)) return Wires(a);
    int i = (int)exp.bits() - 1;  // index of the top set bit
    Fq result = a;
    for (int b = i - 1; b >= 0; b--) {
      Fq sq = fq_square_montgomery(c, result);
      if (exp.bit((unsigned)b)) result = fq_mul_montgomery(c, a, sq);
      else result = sq;
    }
    return result;
  });
}
// fq.rs:291-299
Fq fq_sqrt_montgomery(Builder& c, const Fq& a) {
  static const U256 e = U256::from_dec("5472060717959818805561601436314318772174077789324455915672259473661306552146");
  return fq_exp_by_constant_montgomery(c, a, e);
}
// fq.rs:177-193
Wire fq_is_qnr_montgomery(Builder& c, const Fq& x) {
  const FqConsts& K = FqConsts::get();
  Fq y = fq_exp_by_constant_montgomery(c, x, shr1(sub(K.p, U256(1))));
  Fq neg_one_mont = fq_constant(mont254(sub(K.p, U256(1))));
  return bn_equal(c, y, neg_one_mont);
}
Fq fq_multiplexer(Builder& c, const std::vector<Fq>& a, const Wires& s, size_t w) { return bn_multiplexer(c, a, s, w); }
static Wire fq_equal_constant(Builder& c, const Fq& a, const U256& b) { return bn_equal_constant(c, a, b); }

// =================================================================== fq2.rs (rest)
static Fq2 fq2_neg(Builder& c, const Fq2& a) {
  Fq c0 = fq_neg(c, a.c0);
  Fq c1 = fq_neg(c, a.c1);
  return {c0, c1};
}
static Fq2 fq2_half(Builder& c, const Fq2& a) {
  Fq c0 = fq_half(c, a.c0);
  Fq c1 = fq_half(c, a.c1);
  return {c0, c1};
}
// fq2.rs:171-178 (b: constant whose integer value is added, i.e. already Montgomery when the caller says so)
static Fq2 fq2_add_constant(Builder& c, const Fq2& a, const Fp2& b) {
  Fq c0 = fq_add_constant(c, a.c0, ival(b.c0));
  Fq c1 = fq_add_constant(c, a.c1, ival(b.c1));
  return {c0, c1};
}
// fq2.rs:260-283
static Fq2 fq2_mul_by_constant_montgomery(Builder& c, const Fq2& a, const Fp2& b) {
  if (b == host::fp2_one()) return a;
  Fq a_sum = fq_add(c, a.c0, a.c1);
  Fq a0_b0 = fq_mul_by_constant_montgomery(c, a.c0, ival(b.c0));
  Fq a1_b1 = fq_mul_by_constant_montgomery(c, a.c1, ival(b.c1));
  Fq sum_mul_sum = fq_mul_by_constant_montgomery(c, a_sum, ival(b.c0 + b.c1));
  Fq c0 = fq_sub(c, a0_b0, a1_b1);
  Fq a0b0_plus_a1b1 = fq_add(c, a0_b0, a1_b1);
  Fq c1 = fq_sub(c, sum_mul_sum, a0b0_plus_a1b1);
  return {c0, c1};
}
// fq2.rs:285-295
static Fq2 fq2_mul_by_fq_montgomery(Builder& c, const Fq2& a, const Fq& b) {
  Fq c0 = fq_mul_montgomery(c, a.c0, b);
  Fq c1 = fq_mul_montgomery(c, a.c1, b);
  return {c0, c1};
}
// fq2.rs:309-322
static Fq2 fq2_mul_constant_by_fq_montgomery(Builder& c, const Fp2& a, const Fq& b) {
  Wires out = c.component("fq2::mul_constant_by_fq_montgomery|a=" + hex2(a), b, 2 * FQ_BITS, [a](Builder& c, const Wires& b) {
    U256 a0_m = mont254(ival(a.c0)), a1_m = mont254(ival(a.c1));
    Fq c0 = fq_mul_by_constant_montgomery(c, b, a0_m);
    Fq c1 = fq_mul_by_constant_montgomery(c, b, a1_m);
    return concat(c0, c1);
  });
  return fq2_from_wires(out.data());
}
// fq2.rs:341-354
static Fq2 fq2_square_montgomery(Builder& c, const Fq2& a) {
  Fq a0_plus_a1 = fq_add(c, a.c0, a.c1);
  Fq a0_minus_a1 = fq_sub(c, a.c0, a.c1);
  Fq a0_a1 = fq_mul_montgomery(c, a.c0, a.c1);
  Fq c0 = fq_mul_montgomery(c, a0_plus_a1, a0_minus_a1);
  Fq c1 = fq_double(c, a0_a1);
  return {c0, c1};
}
// fq2.rs:356-372
static Fq2 fq2_inverse_montgomery(Builder& c, const Fq2& a) {
  Wires out = c.component("fq2::inverse_montgomery", to_wires(a), 2 * FQ_BITS, [](Builder& c, const Wires& in) {
    Fq2 a = fq2_from_wires(in.data());
    Fq a0_square = fq_square_montgomery(c, a.c0);
    Fq a1_square = fq_square_montgomery(c, a.c1);
    Fq norm = fq_add(c, a0_square, a1_square);
    Fq inverse_norm = fq_inverse_montgomery(c, norm);
    Fq c0 = fq_mul_montgomery(c, a.c0, inverse_norm);
    Fq neg_a1 = fq_neg(c, a.c1);
    Fq c1 = fq_mul_montgomery(c, neg_a1, inverse_norm);
    return concat(c0, c1);
  });
  return fq2_from_wires(out.data());
}
// fq2.rs:374-384
static Fq2 fq2_frobenius_montgomery(Builder& c, const Fq2& a, size_t i) {
  const host::Params& P = host::Params::get();
  const Fp coef = P.frob_fp2_c1[i % 2].c0;  // FROBENIUS_COEFF_FP2_C1 are Fq elements (1, -1)
  Fq c1 = fq_mul_by_constant_montgomery(c, a.c1, mont254(ival(coef)));
  return {a.c0, c1};
}
// fq2.rs:425-447
static Fq2 fq2_sqrt_general_montgomery(Builder& c, const Fq2& a) {
  Wires out = c.component("fq2::sqrt_general_montgomery", to_wires(a), 2 * FQ_BITS, [](Builder& c, const Wires& in) {
    Fq2 a = fq2_from_wires(in.data());
    // norm_montgomery
    Fq c0_square = fq_square_montgomery(c, a.c0);
    Fq c1_square = fq_square_montgomery(c, a.c1);
    Fq alpha = fq_add(c, c0_square, c1_square);
    Fq alpha_sqrt = fq_sqrt_montgomery(c, alpha);
    Fq delta_plus = fq_add(c, alpha_sqrt, a.c0);
    Fq delta = fq_half(c, delta_plus);
    Wire is_qnr = fq_is_qnr_montgomery(c, delta);
    Fq delta_alt = fq_sub(c, delta, alpha_sqrt);
    Fq delta_final = bn_select(c, delta_alt, delta, is_qnr);
    Fq c0_final = fq_sqrt_montgomery(c, delta_final);
    Fq c0_inv = fq_inverse_montgomery(c, c0_final);
    Fq c1_half = fq_half(c, a.c1);
    Fq c1_final = fq_mul_montgomery(c, c0_inv, c1_half);
    return concat(c0_final, c1_final);
  });
  return fq2_from_wires(out.data());
}
// fq2.rs:146-158
static Wire fq2_equal_constant(Builder& c, const Fq2& a, const Fp2& b) {
  Wire u = fq_equal_constant(c, a.c0, ival(b.c0));
  Wire v = fq_equal_constant(c, a.c1, ival(b.c1));
  Wire w = c.issue_wire();
  c.add_gate(AND, u, v, w);
  return w;
}

// =================================================================== fq6.rs (rest)
static Fq6 fq6_neg(Builder& c, const Fq6& a) {
  Fq2 c0 = fq2_neg(c, a.c0);
  Fq2 c1 = fq2_neg(c, a.c1);
  Fq2 c2 = fq2_neg(c, a.c2);
  return {c0, c1, c2};
}
// fq6.rs:327-333
static Fq6 fq6_mul_by_fq2_montgomery(Builder& c, const Fq6& a, const Fq2& b) {
  Fq2 c0 = fq2_mul_montgomery(c, a.c0, b);
  Fq2 c1 = fq2_mul_montgomery(c, a.c1, b);
  Fq2 c2 = fq2_mul_montgomery(c, a.c2, b);
  return {c0, c1, c2};
}
// fq6.rs:335-345
static Fq6 fq6_mul_by_constant_fq2_montgomery(Builder& c, const Fq6& a, const Fp2& b) {
  Fq2 c0 = fq2_mul_by_constant_montgomery(c, a.c0, b);
  Fq2 c1 = fq2_mul_by_constant_montgomery(c, a.c1, b);
  Fq2 c2 = fq2_mul_by_constant_montgomery(c, a.c2, b);
  return {c0, c1, c2};
}
// fq6.rs:351-380
static Fq6 fq6_mul_by_01_montgomery(Builder& c, const Fq6& a, const Fq2& c0, const Fq2& c1) {
  Fq2 wires_1 = fq2_mul_montgomery(c, a.c0, c0);
  Fq2 wires_2 = fq2_mul_montgomery(c, a.c1, c1);
  Fq2 wires_3 = fq2_add(c, a.c1, a.c2);
  Fq2 wires_4 = fq2_mul_montgomery(c, wires_3, c1);
  Fq2 wires_5 = fq2_sub(c, wires_4, wires_2);
  Fq2 wires_6 = fq2_mul_by_nonresidue(c, wires_5);
  Fq2 wires_7 = fq2_add(c, wires_6, wires_1);
  Fq2 wires_8 = fq2_add(c, a.c0, a.c1);
  Fq2 wires_9 = fq2_add(c, c0, c1);
  Fq2 wires_10 = fq2_mul_montgomery(c, wires_8, wires_9);
  Fq2 wires_11 = fq2_sub(c, wires_10, wires_1);
  Fq2 wires_12 = fq2_sub(c, wires_11, wires_2);
  Fq2 wires_13 = fq2_add(c, a.c0, a.c2);
  Fq2 wires_14 = fq2_mul_montgomery(c, wires_13, c0);
  Fq2 wires_15 = fq2_sub(c, wires_14, wires_1);
  Fq2 wires_16 = fq2_add(c, wires_15, wires_2);
  return {wires_7, wires_12, wires_16};
}
// fq6.rs:382-412
static Fq6 fq6_mul_by_01_constant1_montgomery(Builder& c, const Fq6& a, const Fq2& c0, const Fp2& c1) {
  Fq2 wires_1 = fq2_mul_montgomery(c, a.c0, c0);
  Fq2 wires_2 = fq2_mul_by_constant_montgomery(c, a.c1, c1);
  Fq2 wires_3 = fq2_add(c, a.c1, a.c2);
  Fq2 wires_4 = fq2_mul_by_constant_montgomery(c, wires_3, c1);
  Fq2 wires_5 = fq2_sub(c, wires_4, wires_2);
  Fq2 wires_6 = fq2_mul_by_nonresidue(c, wires_5);
  Fq2 wires_7 = fq2_add(c, wires_6, wires_1);
  Fq2 wires_8 = fq2_add(c, a.c0, a.c1);
  Fq2 wires_9 = fq2_add_constant(c, c0, c1);
  Fq2 wires_10 = fq2_mul_montgomery(c, wires_8, wires_9);
  Fq2 wires_11 = fq2_sub(c, wires_10, wires_1);
  Fq2 wires_12 = fq2_sub(c, wires_11, wires_2);
  Fq2 wires_13 = fq2_add(c, a.c0, a.c2);
  Fq2 wires_14 = fq2_mul_montgomery(c, wires_13, c0);
  Fq2 wires_15 = fq2_sub(c, wires_14, wires_1);
  Fq2 wires_16 = fq2_add(c, wires_15, wires_2);
  return {wires_7, wires_12, wires_16};
}
// fq6.rs:423-451
static Fq6 fq6_square_montgomery(Builder& c, const Fq6& a) {
  Fq2 s_0 = fq2_square_montgomery(c, a.c0);
  Fq2 wires_1 = fq2_add(c, a.c0, a.c2);
  Fq2 wires_2 = fq2_add(c, wires_1, a.c1);
  Fq2 wires_3 = fq2_sub(c, wires_1, a.c1);
  Fq2 s_1 = fq2_square_montgomery(c, wires_2);
  Fq2 s_2 = fq2_square_montgomery(c, wires_3);
  Fq2 wires_4 = fq2_mul_montgomery(c, a.c1, a.c2);
  Fq2 s_3 = fq2_double(c, wires_4);
  Fq2 s_4 = fq2_square_montgomery(c, a.c2);
  Fq2 wires_5 = fq2_add(c, s_1, s_2);
  Fq2 t_1 = fq2_half(c, wires_5);
  Fq2 wires_6 = fq2_mul_by_nonresidue(c, s_3);
  Fq2 res_c0 = fq2_add(c, s_0, wires_6);
  Fq2 wires_7 = fq2_mul_by_nonresidue(c, s_4);
  Fq2 wires_8 = fq2_sub(c, s_1, s_3);
  Fq2 wires_9 = fq2_sub(c, wires_8, t_1);
  Fq2 res_c1 = fq2_add(c, wires_9, wires_7);
  Fq2 wires_10 = fq2_sub(c, t_1, s_0);
  Fq2 res_c2 = fq2_sub(c, wires_10, s_4);
  return {res_c0, res_c1, res_c2};
}
// fq6.rs:453-487
static Fq6 fq6_inverse_montgomery(Builder& c, const Fq6& r) {
  const Fq2 &a = r.c0, &b = r.c1, &cc = r.c2;
  Fq2 a_square = fq2_square_montgomery(c, a);
  Fq2 b_square = fq2_square_montgomery(c, b);
  Fq2 c_square = fq2_square_montgomery(c, cc);
  Fq2 ab = fq2_mul_montgomery(c, a, b);
  Fq2 ac = fq2_mul_montgomery(c, a, cc);
  Fq2 bc = fq2_mul_montgomery(c, b, cc);
  Fq2 bc_beta = fq2_mul_by_nonresidue(c, bc);
  Fq2 a_square_minus_bc_beta = fq2_sub(c, a_square, bc_beta);
  Fq2 c_square_beta = fq2_mul_by_nonresidue(c, c_square);
  Fq2 c_square_beta_minus_ab = fq2_sub(c, c_square_beta, ab);
  Fq2 b_square_minus_ac = fq2_sub(c, b_square, ac);
  Fq2 wires_1 = fq2_mul_montgomery(c, c_square_beta_minus_ab, cc);
  Fq2 wires_2 = fq2_mul_montgomery(c, b_square_minus_ac, b);
  Fq2 wires_1_plus_wires_2 = fq2_add(c, wires_1, wires_2);
  Fq2 wires_3 = fq2_mul_by_nonresidue(c, wires_1_plus_wires_2);
  Fq2 wires_4 = fq2_mul_montgomery(c, a, a_square_minus_bc_beta);
  Fq2 norm = fq2_add(c, wires_4, wires_3);
  Fq2 inverse_norm = fq2_inverse_montgomery(c, norm);
  Fq2 res_c0 = fq2_mul_montgomery(c, a_square_minus_bc_beta, inverse_norm);
  Fq2 res_c1 = fq2_mul_montgomery(c, c_square_beta_minus_ab, inverse_norm);
  Fq2 res_c2 = fq2_mul_montgomery(c, b_square_minus_ac, inverse_norm);
  return {res_c0, res_c1, res_c2};
}
// fq6.rs:489-516
static Fq6 fq6_frobenius_montgomery(Builder& c, const Fq6& a, size_t i) {
  const host::Params& P = host::Params::get();
  Fq2 f0 = fq2_frobenius_montgomery(c, a.c0, i);
  Fq2 f1 = fq2_frobenius_montgomery(c, a.c1, i);
  Fq2 f2 = fq2_frobenius_montgomery(c, a.c2, i);
  Fq2 f1u = fq2_mul_by_constant_montgomery(c, f1, fp2_mont(P.frob_fp6_c1[i % 6]));
  Fq2 f2u = fq2_mul_by_constant_montgomery(c, f2, fp2_mont(P.frob_fp6_c2[i % 6]));
  return {f0, f1u, f2u};
}
// fq6.rs:139-152
static Wire fq6_equal_constant(Builder& c, const Fq6& a, const Fp6& b) {
  Wire u = fq2_equal_constant(c, a.c0, b.c0);
  Wire v = fq2_equal_constant(c, a.c1, b.c1);
  Wire w = fq2_equal_constant(c, a.c2, b.c2);
  Wire x = c.issue_wire();
  Wire y = c.issue_wire();
  c.add_gate(AND, u, v, x);
  c.add_gate(AND, x, w, y);
  return y;
}

// =================================================================== fq12.rs (rest)
// fq12.rs:311-324
static Fq12 fq12_square_montgomery(Builder& c, const Fq12& a) {
  Wires out = c.component("fq12::square_montgomery", to_wires(a), 3048, [](Builder& c, const Wires& in) {
    Fq12 a = fq12_from_wires(in.data());
    Fq6 w1 = fq6_add(c, a.c0, a.c1);
    Fq6 w2 = fq6_mul_by_nonresidue(c, a.c1);
    Fq6 w3 = fq6_add(c, a.c0, w2);
    Fq6 w4 = fq6_mul_montgomery(c, a.c0, a.c1);
    Fq6 w5 = fq6_mul_montgomery(c, w1, w3);
    Fq6 w6 = fq6_mul_by_nonresidue(c, w4);
    Fq6 w7 = fq6_add(c, w4, w6);
    Fq6 c0 = fq6_sub(c, w5, w7);
    Fq6 c1 = fq6_double(c, w4);
    return to_wires(Fq12{c0, c1});
  });
  return fq12_from_wires(out.data());
}
// fq12.rs:326-392
static Fq12 fq12_cyclotomic_square_montgomery(Builder& c, const Fq12& a) {
  const Fq2 c0 = a.c0.c0, c1 = a.c0.c1, c2 = a.c0.c2, c3 = a.c1.c0, c4 = a.c1.c1, c5 = a.c1.c2;
  Fq2 xy = fq2_mul_montgomery(c, c0, c4);
  Fq2 x_plus_y = fq2_add(c, c0, c4);
  Fq2 y_beta = fq2_mul_by_nonresidue(c, c4);
  Fq2 x_plus_y_beta = fq2_add(c, c0, y_beta);
  Fq2 xy_beta = fq2_mul_by_nonresidue(c, xy);
  Fq2 w1 = fq2_mul_montgomery(c, x_plus_y, x_plus_y_beta);
  Fq2 w2 = fq2_add(c, xy, xy_beta);
  Fq2 t0 = fq2_sub(c, w1, w2);
  Fq2 t1 = fq2_double(c, xy);

  xy = fq2_mul_montgomery(c, c2, c3);
  x_plus_y = fq2_add(c, c2, c3);
  y_beta = fq2_mul_by_nonresidue(c, c2);
  x_plus_y_beta = fq2_add(c, c3, y_beta);
  xy_beta = fq2_mul_by_nonresidue(c, xy);
  w1 = fq2_mul_montgomery(c, x_plus_y, x_plus_y_beta);
  w2 = fq2_add(c, xy, xy_beta);
  Fq2 t2 = fq2_sub(c, w1, w2);
  Fq2 t3 = fq2_double(c, xy);

  xy = fq2_mul_montgomery(c, c1, c5);
  x_plus_y = fq2_add(c, c1, c5);
  y_beta = fq2_mul_by_nonresidue(c, c5);
  x_plus_y_beta = fq2_add(c, c1, y_beta);
  xy_beta = fq2_mul_by_nonresidue(c, xy);
  w1 = fq2_mul_montgomery(c, x_plus_y, x_plus_y_beta);
  w2 = fq2_add(c, xy, xy_beta);
  Fq2 t4 = fq2_sub(c, w1, w2);
  Fq2 t5 = fq2_double(c, xy);

  w1 = fq2_sub(c, t0, c0);
  w2 = fq2_double(c, w1);
  Fq2 z0 = fq2_add(c, w2, t0);
  w1 = fq2_sub(c, t2, c1);
  w2 = fq2_double(c, w1);
  Fq2 z4 = fq2_add(c, w2, t2);
  w1 = fq2_sub(c, t4, c2);
  w2 = fq2_double(c, w1);
  Fq2 z3 = fq2_add(c, w2, t4);
  Fq2 t5_beta = fq2_mul_by_nonresidue(c, t5);
  w1 = fq2_add(c, t5_beta, c3);
  w2 = fq2_double(c, w1);
  Fq2 z2 = fq2_add(c, w2, t5_beta);
  w1 = fq2_add(c, t1, c4);
  w2 = fq2_double(c, w1);
  Fq2 z1 = fq2_add(c, w2, t1);
  w1 = fq2_add(c, t3, c5);
  w2 = fq2_double(c, w1);
  Fq2 z5 = fq2_add(c, w2, t3);
  return {Fq6{z0, z4, z3}, Fq6{z2, z1, z5}};
}
// fq12.rs:413-428
static Fq12 fq12_inverse_montgomery(Builder& c, const Fq12& a) {
  Wires out = c.component("fq12::inverse_montgomery", to_wires(a), 3048, [](Builder& c, const Wires& in) {
    Fq12 a = fq12_from_wires(in.data());
    Fq6 a_c0_square = fq6_square_montgomery(c, a.c0);
    Fq6 a_c1_square = fq6_square_montgomery(c, a.c1);
    Fq6 a_c1_square_beta = fq6_mul_by_nonresidue(c, a_c1_square);
    Fq6 norm = fq6_sub(c, a_c0_square, a_c1_square_beta);
    Fq6 inverse_norm = fq6_inverse_montgomery(c, norm);
    Fq6 res_c0 = fq6_mul_montgomery(c, a.c0, inverse_norm);
    Fq6 neg_a_c1 = fq6_neg(c, a.c1);
    Fq6 res_c1 = fq6_mul_montgomery(c, inverse_norm, neg_a_c1);
    return to_wires(Fq12{res_c0, res_c1});
  });
  return fq12_from_wires(out.data());
}
// fq12.rs:430-442
static Fq12 fq12_frobenius_montgomery(Builder& c, const Fq12& a, size_t i) {
  const host::Params& P = host::Params::get();
  Fq6 f0 = fq6_frobenius_montgomery(c, a.c0, i);
  Fq6 f1 = fq6_frobenius_montgomery(c, a.c1, i);
  Fq6 x = fq6_mul_by_constant_fq2_montgomery(c, f1, fp2_mont(P.frob_fp12_c1[i % 12]));
  return {f0, x};
}
// fq12.rs:444-447
static Fq12 fq12_conjugate(Builder& c, const Fq12& a) {
  Fq6 new_c1 = fq6_neg(c, a.c1);
  return {a.c0, new_c1};
}
// fq12.rs:267-285
static Fq12 fq12_mul_by_034_montgomery(Builder& c, const Fq12& a, const Fq2& c0, const Fq2& c3, const Fq2& c4) {
  Wires in = concat(concat(to_wires(a), to_wires(c0)), concat(to_wires(c3), to_wires(c4)));
  Wires out = c.component("fq12::mul_by_034_montgomery", in, 3048, [](Builder& c, const Wires& in) {
    Fq12 a = fq12_from_wires(in.data());
    Fq2 c0 = fq2_from_wires(in.data() + 3048), c3 = fq2_from_wires(in.data() + 3556), c4 = fq2_from_wires(in.data() + 4064);
    Fq6 w1 = fq6_mul_by_01_montgomery(c, a.c1, c3, c4);
    Fq6 w2 = fq6_mul_by_nonresidue(c, w1);
    Fq6 w3 = fq6_mul_by_fq2_montgomery(c, a.c0, c0);
    Fq6 new_c0 = fq6_add(c, w2, w3);
    Fq6 w4 = fq6_add(c, a.c0, a.c1);
    Fq2 w5 = fq2_add(c, c3, c0);
    Fq6 w6 = fq6_mul_by_01_montgomery(c, w4, w5, c4);
    Fq6 w7 = fq6_add(c, w1, w3);
    Fq6 new_c1 = fq6_sub(c, w6, w7);
    return to_wires(Fq12{new_c0, new_c1});
  });
  return fq12_from_wires(out.data());
}
// fq12.rs:287-309
static Fq12 fq12_mul_by_034_constant4_montgomery(Builder& c, const Fq12& a, const Fq2& c0, const Fq2& c3, const Fp2& c4) {
  Wires in = concat(to_wires(a), concat(to_wires(c0), to_wires(c3)));
  Wires out = c.component("fq12::mul_by_034_constant4_montgomery|c4=" + hex2(c4), in, 3048, [c4](Builder& c, const Wires& in) {
    Fq12 a = fq12_from_wires(in.data());
    Fq2 c0 = fq2_from_wires(in.data() + 3048), c3 = fq2_from_wires(in.data() + 3556);
    Fq6 w1 = fq6_mul_by_01_constant1_montgomery(c, a.c1, c3, c4);
    Fq6 w2 = fq6_mul_by_nonresidue(c, w1);
    Fq6 w3 = fq6_mul_by_fq2_montgomery(c, a.c0, c0);
    Fq6 new_c0 = fq6_add(c, w2, w3);
    Fq6 w4 = fq6_add(c, a.c0, a.c1);
    Fq2 w5 = fq2_add(c, c3, c0);
    Fq6 w6 = fq6_mul_by_01_constant1_montgomery(c, w4, w5, c4);
    Fq6 w7 = fq6_add(c, w1, w3);
    Fq6 new_c1 = fq6_sub(c, w6, w7);
    return to_wires(Fq12{new_c0, new_c1});
  });
  return fq12_from_wires(out.data());
}
// fq12.rs:158-168
static Wire fq12_equal_constant(Builder& c, const Fq12& a, const Fp12& b) {
  Wire u = fq6_equal_constant(c, a.c0, b.c0);
  Wire v = fq6_equal_constant(c, a.c1, b.c1);
  Wire w = c.issue_wire();
  c.add_gate(AND, u, v, w);
  return w;
}

// =================================================================== g1.rs
Wires to_wires(const G1P& p) { return concat(concat(p.x, p.y), p.z); }
G1P g1_from_wires(const Wire* w) { return {Fq(w, w + 254), Fq(w + 254, w + 508), Fq(w + 508, w + 762)}; }
// G1Projective::new_constant over Montgomery-converted Jacobian coordinates
static G1P g1_constant_montgomery(const host::G1Jac& p) {
  return {fq_constant(mont254(ival(p.x))), fq_constant(mont254(ival(p.y))), fq_constant(mont254(ival(p.z)))};
}
// g1.rs:159-235
G1P g1_add_montgomery(Builder& c, const G1P& p, const G1P& q) {
  Wires out = c.component("g1::add_montgomery", concat(to_wires(p), to_wires(q)), 762, [](Builder& c, const Wires& in) {
    G1P p = g1_from_wires(in.data()), q = g1_from_wires(in.data() + 762);
    const Fq &x1 = p.x, &y1 = p.y, &z1 = p.z, &x2 = q.x, &y2 = q.y, &z2 = q.z;
    Fq z1s = fq_square_montgomery(c, z1);
    Fq z2s = fq_square_montgomery(c, z2);
    Fq z1c = fq_mul_montgomery(c, z1s, z1);
    Fq z2c = fq_mul_montgomery(c, z2s, z2);
    Fq u1 = fq_mul_montgomery(c, x1, z2s);
    Fq u2 = fq_mul_montgomery(c, x2, z1s);
    Fq s1 = fq_mul_montgomery(c, y1, z2c);
    Fq s2 = fq_mul_montgomery(c, y2, z1c);
    Fq r = fq_sub(c, s1, s2);
    Fq h = fq_sub(c, u1, u2);
    Fq h2 = fq_square_montgomery(c, h);
    Fq g = fq_mul_montgomery(c, h, h2);
    Fq v = fq_mul_montgomery(c, u1, h2);
    Fq r2 = fq_square_montgomery(c, r);
    Fq r2g = fq_add(c, r2, g);
    Fq vd = fq_double(c, v);
    Fq x3 = fq_sub(c, r2g, vd);
    Fq vx3 = fq_sub(c, v, x3);
    Fq w = fq_mul_montgomery(c, r, vx3);
    Fq s1g = fq_mul_montgomery(c, s1, g);
    Fq y3 = fq_sub(c, w, s1g);
    Fq z1z2 = fq_mul_montgomery(c, z1, z2);
    Fq z3 = fq_mul_montgomery(c, z1z2, h);
    Wire z1_0 = fq_equal_constant(c, z1, U256());
    Wire z2_0 = fq_equal_constant(c, z2, U256());
    Fq zero = fq_constant(U256());
    Wires s{z1_0, z2_0};
    Fq x = fq_multiplexer(c, {x3, x2, x1, zero}, s, 2);
    Fq y = fq_multiplexer(c, {y3, y2, y1, zero}, s, 2);
    Fq z = fq_multiplexer(c, {z3, z2, z1, zero}, s, 2);
    return to_wires(G1P{x, y, z});
  });
  return g1_from_wires(out.data());
}
// g1.rs:275-306
static G1P g1_multiplexer(Builder& c, const std::vector<G1P>& a, const Wires& s, size_t w) {
  Wires in;
  in.reserve(a.size() * 762 + w);
  for (const G1P& p : a) {
    in.insert(in.end(), p.x.begin(), p.x.end());
    in.insert(in.end(), p.y.begin(), p.y.end());
    in.insert(in.end(), p.z.begin(), p.z.end());
  }
  in.insert(in.end(), s.begin(), s.end());
  const size_t n = a.size();
  Wires out = c.component("g1::multiplexer|w=" + std::to_string(w), in, 762, [n, w](Builder& c, const Wires& in) {
    std::vector<Fq> xs(n), ys(n), zs(n);
    for (size_t k = 0; k < n; k++) {
      G1P p = g1_from_wires(in.data() + k * 762);
      xs[k] = p.x; ys[k] = p.y; zs[k] = p.z;
    }
    Wires s(in.begin() + n * 762, in.end());
    Fq x = fq_multiplexer(c, xs, s, w);
    Fq y = fq_multiplexer(c, ys, s, w);
    Fq z = fq_multiplexer(c, zs, s, w);
    return to_wires(G1P{x, y, z});
  });
  return g1_from_wires(out.data());
}
// g1.rs:308-368 (W = 10)
static G1P g1_scalar_mul_by_constant_base_montgomery(Builder& c, const Wires& s, const host::G1Jac& base) {
  const size_t W = 10, FR_BITS = 254;
  if (s.size() != FR_BITS) throw std::logic_error("scalar must have 254 bits");
  std::string key = "g1::scalar_mul_by_constant_base_montgomery|base=" + ival(base.x).to_hex() + ival(base.y).to_hex() +
                    ival(base.z).to_hex();
  Wires out = c.component(key, s, 762, [base, W, FR_BITS](Builder& c, const Wires& s) {
    const size_t n = (size_t)1 << W;
    std::vector<host::G1Jac> bases;
    host::G1Jac p = host::g1_zero();
    for (size_t i = 0; i < n; i++) {
      bases.push_back(p);
      p = host::g1_add(p, base);
    }
    auto wires_of = [](const std::vector<host::G1Jac>& b) {
      std::vector<G1P> w;
      w.reserve(b.size());
      for (const host::G1Jac& q : b) w.push_back(g1_constant_montgomery(q));
      return w;
    };
    std::vector<G1P> bases_wires = wires_of(bases);
    std::vector<G1P> to_be_added;
    size_t index = 0;
    while (index < FR_BITS) {
      size_t w = std::min(W, FR_BITS - index);
      size_t m = (size_t)1 << w;
      Wires selector(s.begin() + index, s.begin() + index + w);
      std::vector<G1P> sub(bases_wires.begin(), bases_wires.begin() + m);
      to_be_added.push_back(g1_multiplexer(c, sub, selector, w));
      index += W;
      for (host::G1Jac& b : bases)
        for (size_t k = 0; k < w; k++) b = host::g1_add(b, b);
      bases_wires = wires_of(bases);
    }
    G1P acc = to_be_added[0];
    for (size_t i = 1; i < to_be_added.size(); i++) acc = g1_add_montgomery(c, acc, to_be_added[i]);
    return to_wires(acc);
  });
  return g1_from_wires(out.data());
}
// g1.rs:370-400
static G1P g1_msm_with_constant_bases_montgomery(Builder& c, const std::vector<Wires>& scalars,
                                                 const std::vector<host::G1Jac>& bases) {
  std::string key = "g1::msm_with_constant_bases_montgomery|bases=";
  for (const host::G1Jac& b : bases) key += ival(b.x).to_hex() + ival(b.y).to_hex() + ival(b.z).to_hex();
  Wires in;
  for (const Wires& s : scalars) in.insert(in.end(), s.begin(), s.end());
  const size_t n_s = scalars.size();
  Wires out = c.component(key, in, 762, [bases, n_s](Builder& c, const Wires& in) {
    if (n_s == 0) return to_wires(g1_constant_montgomery(host::g1_zero()));
    std::vector<G1P> to_be_added;
    for (size_t i = 0; i < n_s; i++)
      to_be_added.push_back(g1_scalar_mul_by_constant_base_montgomery(c, slice(in, i * 254, (i + 1) * 254), bases[i]));
    G1P acc = to_be_added[0];
    for (size_t i = 1; i < to_be_added.size(); i++) acc = g1_add_montgomery(c, acc, to_be_added[i]);
    return to_wires(acc);
  });
  return g1_from_wires(out.data());
}

// =================================================================== pairing.rs (Groth16 path)
Wires to_wires(const G2P& p) { return concat(concat(to_wires(p.x), to_wires(p.y)), to_wires(p.z)); }
G2P g2_from_wires(const Wire* w) { return {fq2_from_wires(w), fq2_from_wires(w + 508), fq2_from_wires(w + 1016)}; }

// pairing.rs:359-407
static void double_in_place_circuit_montgomery(Builder& c, const G2P& r, G2P& r_out, Fq6& coeffs) {
  Wires out = c.component("pairing::double_in_place_circuit_montgomery", to_wires(r), 1524 + 1524, [](Builder& c, const Wires& in) {
    const host::Params& P = host::Params::get();
    G2P r = g2_from_wires(in.data());
    const Fq2 &rx = r.x, &ry = r.y, &rz = r.z;
    Fq2 a = fq2_mul_montgomery(c, rx, ry);
    a = fq2_half(c, a);
    Fq2 b = fq2_square_montgomery(c, ry);
    Fq2 cc = fq2_square_montgomery(c, rz);
    Fq2 c_triple = fq2_triple(c, cc);
    Fq2 e = fq2_mul_by_constant_montgomery(c, c_triple, fp2_mont(P.g2_b));
    Fq2 f = fq2_triple(c, e);
    Fq2 g = fq2_add(c, b, f);
    g = fq2_half(c, g);
    Fq2 ryrz = fq2_add(c, ry, rz);
    Fq2 ryrzs = fq2_square_montgomery(c, ryrz);
    Fq2 bc = fq2_add(c, b, cc);
    Fq2 h = fq2_sub(c, ryrzs, bc);
    Fq2 i = fq2_sub(c, e, b);
    Fq2 j = fq2_square_montgomery(c, rx);
    Fq2 es = fq2_square_montgomery(c, e);
    Fq2 j_triple = fq2_triple(c, j);
    Fq2 bf = fq2_sub(c, b, f);
    Fq2 new_x = fq2_mul_montgomery(c, a, bf);
    Fq2 es_triple = fq2_triple(c, es);
    Fq2 gs = fq2_square_montgomery(c, g);
    Fq2 new_y = fq2_sub(c, gs, es_triple);
    Fq2 new_z = fq2_mul_montgomery(c, b, h);
    Fq2 hn = fq2_neg(c, h);
    return concat(to_wires(G2P{new_x, new_y, new_z}), to_wires(Fq6{hn, j_triple, i}));
  });
  r_out = g2_from_wires(out.data());
  coeffs = fq6_from_wires(out.data() + 1524);
}
// pairing.rs:409-462
static void add_in_place_montgomery(Builder& c, const G2P& r, const G2P& q, G2P& r_out, Fq6& coeffs) {
  Wires out = c.component("pairing::add_in_place_montgomery", concat(to_wires(r), to_wires(q)), 1524 + 1524, [](Builder& c, const Wires& in) {
    G2P r = g2_from_wires(in.data()), q = g2_from_wires(in.data() + 1524);
    const Fq2 &rx = r.x, &ry = r.y, &rz = r.z, &qx = q.x, &qy = q.y;
    Fq2 wires_1 = fq2_mul_montgomery(c, qy, rz);
    Fq2 theta = fq2_sub(c, ry, wires_1);
    Fq2 wires_2 = fq2_mul_montgomery(c, qx, rz);
    Fq2 lambda = fq2_sub(c, rx, wires_2);
    Fq2 cc = fq2_square_montgomery(c, theta);
    Fq2 d = fq2_square_montgomery(c, lambda);
    Fq2 e = fq2_mul_montgomery(c, lambda, d);
    Fq2 f = fq2_mul_montgomery(c, rz, cc);
    Fq2 g = fq2_mul_montgomery(c, rx, d);
    Fq2 wires_3 = fq2_add(c, e, f);
    Fq2 wires_4 = fq2_double(c, g);
    Fq2 h = fq2_sub(c, wires_3, wires_4);
    Fq2 neg_theta = fq2_neg(c, theta);
    Fq2 wires_5 = fq2_mul_montgomery(c, theta, qx);
    Fq2 wires_6 = fq2_mul_montgomery(c, lambda, qy);
    Fq2 j = fq2_sub(c, wires_5, wires_6);
    Fq2 new_r_x = fq2_mul_montgomery(c, lambda, h);
    Fq2 wires_7 = fq2_sub(c, g, h);
    Fq2 wires_8 = fq2_mul_montgomery(c, theta, wires_7);
    Fq2 wires_9 = fq2_mul_montgomery(c, e, ry);
    Fq2 new_r_y = fq2_sub(c, wires_8, wires_9);
    Fq2 new_r_z = fq2_mul_montgomery(c, rz, e);
    return concat(to_wires(G2P{new_r_x, new_r_y, new_r_z}), to_wires(Fq6{lambda, neg_theta, j}));
  });
  r_out = g2_from_wires(out.data());
  coeffs = fq6_from_wires(out.data() + 1524);
}
// pairing.rs:464-471
static G2P g2_affine_neg_evaluate(Builder& c, const G2P& q) {
  G2P result = q;
  result.y = fq2_neg(c, q.y);
  return result;
}
// pairing.rs:473-498
static G2P mul_by_char_montgomery(Builder& c, const G2P& r) {
  Wires out = c.component("pairing::mul_by_char_montgomery", to_wires(r), 1524, [](Builder& c, const Wires& in) {
    const host::Params& P = host::Params::get();
    G2P r = g2_from_wires(in.data());
    Fq2 s_x = fq2_frobenius_montgomery(c, r.x, 1);
    s_x = fq2_mul_by_constant_montgomery(c, s_x, fp2_mont(P.twist_mul_by_q_x));
    Fq2 s_y = fq2_frobenius_montgomery(c, r.y, 1);
    s_y = fq2_mul_by_constant_montgomery(c, s_y, fp2_mont(P.twist_mul_by_q_y));
    return to_wires(G2P{s_x, s_y, r.z});
  });
  return g2_from_wires(out.data());
}
// pairing.rs:507-545
static std::vector<Fq6> ell_coeffs_montgomery(Builder& c, const G2P& q) {
  const host::Params& P = host::Params::get();
  G2P neg_q = g2_affine_neg_evaluate(c, q);
  std::vector<Fq6> ellc;
  G2P r = q;
  for (int i = (int)P.ate_loop.size() - 2; i >= 0; i--) {
    G2P new_r;
    Fq6 coeffs;
    double_in_place_circuit_montgomery(c, r, new_r, coeffs);
    ellc.push_back(coeffs);
    r = new_r;
    if (P.ate_loop[i] == 1) {
      add_in_place_montgomery(c, r, q, new_r, coeffs);
      ellc.push_back(coeffs);
      r = new_r;
    } else if (P.ate_loop[i] == -1) {
      add_in_place_montgomery(c, r, neg_q, new_r, coeffs);
      ellc.push_back(coeffs);
      r = new_r;
    }
  }
  G2P q1 = mul_by_char_montgomery(c, q);
  G2P q2 = mul_by_char_montgomery(c, q1);
  q2 = g2_affine_neg_evaluate(c, q2);
  G2P new_r;
  Fq6 coeffs;
  add_in_place_montgomery(c, r, q1, new_r, coeffs);
  ellc.push_back(coeffs);
  r = new_r;
  add_in_place_montgomery(c, r, q2, new_r, coeffs);
  ellc.push_back(coeffs);
  return ellc;
}
// pairing.rs:923-942
static Fq12 ell_by_constant_montgomery(Builder& c, const Fq12& f, const Fp6& coeffs, const G1P& p) {
  Wires out = c.component("pairing::ell_by_constant_montgomery|coeffs=" + hex6(coeffs), concat(to_wires(f), to_wires(p)), 3048,
                          [coeffs](Builder& c, const Wires& in) {
    Fq12 f = fq12_from_wires(in.data());
    G1P p = g1_from_wires(in.data() + 3048);
    Fq2 new_c0 = fq2_mul_constant_by_fq_montgomery(c, coeffs.c0, p.y);
    Fq2 new_c1 = fq2_mul_constant_by_fq_montgomery(c, coeffs.c1, p.x);
    Fp2 c2_m = fp2_mont(coeffs.c2);
    return to_wires(fq12_mul_by_034_constant4_montgomery(c, f, new_c0, new_c1, c2_m));
  });
  return fq12_from_wires(out.data());
}
// pairing.rs:160-171
static Fq12 ell_montgomery(Builder& c, const Fq12& f, const Fq6& coeffs, const G1P& p) {
  Fq2 c0_fq2 = fq2_mul_by_fq_montgomery(c, coeffs.c0, p.y);
  Fq2 c3_fq2 = fq2_mul_by_fq_montgomery(c, coeffs.c1, p.x);
  return fq12_mul_by_034_montgomery(c, f, c0_fq2, c3_fq2, coeffs.c2);
}
// pairing.rs:944-1009
static Fq12 multi_miller_loop_groth16_evaluate_montgomery_fast(Builder& c, const G1P& p1, const G1P& p2, const G1P& p3,
                                                                const host::G2Affine& q1, const host::G2Affine& q2, const G2P& q3) {
  std::string key = "pairing::multi_miller_loop_groth16_evaluate_montgomery_fast|q1=" + hex2(q1.x) + hex2(q1.y) + "|q2=" +
                    hex2(q2.x) + hex2(q2.y);
  Wires in = concat(concat(to_wires(p1), to_wires(p2)), concat(to_wires(p3), to_wires(q3)));
  Wires out = c.component(key, in, 3048, [q1, q2](Builder& c, const Wires& in) {
    const host::Params& P = host::Params::get();
    G1P p1 = g1_from_wires(in.data()), p2 = g1_from_wires(in.data() + 762), p3 = g1_from_wires(in.data() + 1524);
    G2P q3 = g2_from_wires(in.data() + 2286);
    std::vector<Fp6> q1ell = host::ell_coeffs(q1);
    std::vector<Fp6> q2ell = host::ell_coeffs(q2);
    std::vector<Fq6> q3ell = ell_coeffs_montgomery(c, q3);
    size_t i1 = 0, i2 = 0, i3 = 0;
    Fq12 f = fq12_constant_montgomery(host::fp12_one());
    auto step = [&]() {
      f = ell_by_constant_montgomery(c, f, q1ell[i1++], p1);
      f = ell_by_constant_montgomery(c, f, q2ell[i2++], p2);
      f = ell_montgomery(c, f, q3ell[i3++], p3);
    };
    const int n = (int)P.ate_loop.size();
    for (int i = n - 1; i >= 1; i--) {
      if (i != n - 1) f = fq12_square_montgomery(c, f);
      step();
      int bit = P.ate_loop[i - 1];
      if (bit == 1 || bit == -1) step();
    }
    step();
    step();
    return to_wires(f);
  });
  return fq12_from_wires(out.data());
}

// =================================================================== final_exponentiation.rs
// final_exponentiation.rs:66-92
static Fq12 cyclotomic_exp_fast_inverse_montgomery_fast(Builder& c, const Fq12& f) {
  const host::Params& P = host::Params::get();
  Fq12 res = fq12_constant_montgomery(host::fp12_one());
  Fq12 f_inverse = fq12_inverse_montgomery(c, f);
  bool found_nonzero = false;
  for (int i = (int)P.x_naf.size() - 1; i >= 0; i--) {
    int value = P.x_naf[i];
    if (found_nonzero) res = fq12_cyclotomic_square_montgomery(c, res);
    if (value != 0) {
      found_nonzero = true;
      if (value > 0) res = fq12_mul_montgomery(c, res, f);
      else res = fq12_mul_montgomery(c, res, f_inverse);
    }
  }
  return res;
}
// final_exponentiation.rs:94-97
static Fq12 exp_by_neg_x_montgomery(Builder& c, const Fq12& f) {
  Fq12 f2 = cyclotomic_exp_fast_inverse_montgomery_fast(c, f);
  return fq12_conjugate(c, f2);
}
// final_exponentiation.rs:99-131
static Fq12 final_exponentiation_montgomery(Builder& c, const Fq12& f) {
  Wires out = c.component("final_exponentiation::final_exponentiation_montgomery", to_wires(f), 3048, [](Builder& c, const Wires& in) {
    Fq12 f = fq12_from_wires(in.data());
    Fq12 f_inv = fq12_inverse_montgomery(c, f);
    Fq12 f_conjugate = fq12_conjugate(c, f);
    Fq12 u = fq12_mul_montgomery(c, f_inv, f_conjugate);
    Fq12 u_frobenius = fq12_frobenius_montgomery(c, u, 2);
    Fq12 r = fq12_mul_montgomery(c, u_frobenius, u);
    Fq12 y0 = exp_by_neg_x_montgomery(c, r);
    Fq12 y1 = fq12_square_montgomery(c, y0);
    Fq12 y2 = fq12_square_montgomery(c, y1);
    Fq12 y3 = fq12_mul_montgomery(c, y1, y2);
    Fq12 y4 = exp_by_neg_x_montgomery(c, y3);
    Fq12 y5 = fq12_square_montgomery(c, y4);
    Fq12 y6 = exp_by_neg_x_montgomery(c, y5);
    Fq12 y7 = fq12_conjugate(c, y3);
    Fq12 y8 = fq12_conjugate(c, y6);
    Fq12 y9 = fq12_mul_montgomery(c, y8, y4);
    Fq12 y10 = fq12_mul_montgomery(c, y9, y7);
    Fq12 y11 = fq12_mul_montgomery(c, y10, y1);
    Fq12 y12 = fq12_mul_montgomery(c, y10, y4);
    Fq12 y13 = fq12_mul_montgomery(c, y12, r);
    Fq12 y14 = fq12_frobenius_montgomery(c, y11, 1);
    Fq12 y15 = fq12_mul_montgomery(c, y14, y13);
    Fq12 y16 = fq12_frobenius_montgomery(c, y10, 2);
    Fq12 y17 = fq12_mul_montgomery(c, y16, y15);
    Fq12 r2 = fq12_conjugate(c, r);
    Fq12 y18 = fq12_mul_montgomery(c, r2, y11);
    Fq12 y19 = fq12_frobenius_montgomery(c, y18, 3);
    return to_wires(fq12_mul_montgomery(c, y19, y17));
  });
  return fq12_from_wires(out.data());
}

// =================================================================== groth16.rs
// groth16.rs:26-47
static G1P projective_to_affine_montgomery(Builder& c, const G1P& p) {
  Wires out = c.component("groth16::projective_to_affine_montgomery", to_wires(p), 762, [](Builder& c, const Wires& in) {
    G1P p = g1_from_wires(in.data());
    Fq z_inverse = fq_inverse_montgomery(c, p.z);
    Fq z_inverse_square = fq_square_montgomery(c, z_inverse);
    Fq z_inverse_cube = fq_mul_montgomery(c, z_inverse, z_inverse_square);
    Fq new_x = fq_mul_montgomery(c, p.x, z_inverse_square);
    Fq new_y = fq_mul_montgomery(c, p.y, z_inverse_cube);
    return to_wires(G1P{new_x, new_y, fq_constant(mont254(U256(1)))});
  });
  return g1_from_wires(out.data());
}
// groth16.rs:57-110
Wire groth16_verify(Builder& c, const std::vector<Wires>& publics, const G1P& a, const G2P& b, const G1P& cc,
                    const host::VerifyingKey& vk) {
  std::vector<host::G1Jac> bases;
  for (size_t i = 0; i < publics.size(); i++) bases.push_back(host::g1_from_affine(vk.gamma_abc_g1[i + 1]));
  G1P msm_temp = g1_msm_with_constant_bases_montgomery(c, publics, bases);
  G1P gamma0 = g1_constant_montgomery(host::g1_from_affine(vk.gamma_abc_g1[0]));
  G1P msm = g1_add_montgomery(c, msm_temp, gamma0);
  G1P msm_affine = projective_to_affine_montgomery(c, msm);
  Fq12 f = multi_miller_loop_groth16_evaluate_montgomery_fast(c, msm_affine, cc, a, host::g2_neg(vk.gamma_g2),
                                                              host::g2_neg(vk.delta_g2), b);
  Fp12 alpha_beta = host::fp12_inv(host::final_exponentiation(host::miller_loop({vk.alpha_g1}, {host::g2_neg(vk.beta_g2)})));
  f = final_exponentiation_montgomery(c, f);
  Fp12 ab_m{fp6_mont(alpha_beta.c0), fp6_mont(alpha_beta.c1)};
  return fq12_equal_constant(c, f, ab_m);
}
// groth16.rs:115-143
static G1P decompress_g1_from_compressed(Builder& c, const Fq& x_m, Wire y_flag) {
  Wires in(x_m);
  in.push_back(y_flag);
  Wires out = c.component("groth16::decompress_g1_from_compressed", in, 762, [](Builder& c, const Wires& in) {
    Fq x_m = slice(in, 0, 254);
    Wire y_flag = in[254];
    Fq x2 = fq_square_montgomery(c, x_m);
    Fq x3 = fq_mul_montgomery(c, x2, x_m);
    Fq rhs = fq_add_constant(c, x3, mont254(U256(3)));
    Fq sy = fq_sqrt_montgomery(c, rhs);
    Fq sy_neg = fq_neg(c, sy);
    Fq y = bn_select(c, sy, sy_neg, y_flag);
    return to_wires(G1P{x_m, y, fq_constant(mont254(U256(1)))});
  });
  return g1_from_wires(out.data());
}
// groth16.rs:145-182
static G2P decompress_g2_from_compressed(Builder& c, const Fq2& x, Wire y_flag) {
  Wires in = to_wires(x);
  in.push_back(y_flag);
  Wires out = c.component("groth16::decompress_g2_from_compressed", in, 1524, [](Builder& c, const Wires& in) {
    const host::Params& P = host::Params::get();
    Fq2 x = fq2_from_wires(in.data());
    Wire y_flag = in[508];
    Fq2 x2 = fq2_square_montgomery(c, x);
    Fq2 x3 = fq2_mul_montgomery(c, x2, x);
    Fq2 y2 = fq2_add_constant(c, x3, fp2_mont(P.g2_b));
    Fq2 y = fq2_sqrt_general_montgomery(c, y2);
    Fq2 neg_y = fq2_neg(c, y);
    Fq final_y_0 = bn_select(c, y.c0, neg_y.c0, y_flag);
    Fq final_y_1 = bn_select(c, y.c1, neg_y.c1, y_flag);
    Fq2 z{fq_constant(mont254(U256(1))), fq_constant(mont254(U256()))};
    return to_wires(G2P{x, Fq2{final_y_0, final_y_1}, z});
  });
  return g2_from_wires(out.data());
}
// groth16.rs:250-268.  Inputs: public (254 each), a = (x, flag), b = (x.c0, x.c1, flag), c = (x, flag)
Wire groth16_verify_compressed(Builder& c, const Wires& in, size_t n_public, const host::VerifyingKey& vk) {
  size_t o = 0;
  std::vector<Wires> publics;
  for (size_t i = 0; i < n_public; i++, o += 254) publics.push_back(slice(in, o, o + 254));
  Fq ax = slice(in, o, o + 254); Wire aflag = in[o + 254]; o += 255;
  Fq2 bx = fq2_from_wires(in.data() + o); Wire bflag = in[o + 508]; o += 509;
  Fq cx = slice(in, o, o + 254); Wire cflag = in[o + 254]; o += 255;
  G1P a = decompress_g1_from_compressed(c, ax, aflag);
  G2P b = decompress_g2_from_compressed(c, bx, bflag);
  G1P cc = decompress_g1_from_compressed(c, cx, cflag);
  return groth16_verify(c, publics, a, b, cc, vk);
}

// =================================================================== roots for tests and workloads
uint32_t build_groth16_verify_compressed(Builder& b, const host::VerifyingKey& vk, size_t n_public) {
  size_t n_in = n_public * 254 + 255 + 509 + 255;
  return b.build_root("groth16_verify_compressed", n_in, [vk, n_public](Builder& c, const Wires& in) {
    return Wires{groth16_verify_compressed(c, in, n_public, vk)};
  });
}
// groth16_verify on uncompressed affine points with z = 1 constants (src/garbled_groth16.rs:108-176):
// inputs in EncodeInput order: public (254 each), a.x, a.y, b.x (c0, c1), b.y (c0, c1), c.x, c.y = 2 286 wires
// for one public input -- what examples/groth16_garble.rs garbles.
uint32_t build_groth16_verify(Builder& b, const host::VerifyingKey& vk, size_t n_public) {
  size_t n_in = n_public * 254 + 8 * 254;
  return b.build_root("groth16_verify", n_in, [vk, n_public](Builder& c, const Wires& in) {
    size_t o = 0;
    std::vector<Wires> publics;
    for (size_t i = 0; i < n_public; i++, o += 254) publics.push_back(slice(in, o, o + 254));
    const Fq one = fq_constant(mont254(U256(1))), zero = fq_constant(mont254(U256()));
    G1P a{slice(in, o, o + 254), slice(in, o + 254, o + 508), one};
    o += 508;
    G2P bq{fq2_from_wires(in.data() + o), fq2_from_wires(in.data() + o + 508), Fq2{one, zero}};
    o += 1016;
    G1P cc{slice(in, o, o + 254), slice(in, o + 254, o + 508), one};
    return Wires{groth16_verify(c, publics, a, bq, cc, vk)};
  });
}
uint32_t build_fq_inverse(Builder& b) {
  return b.build_root("fq_inverse_montgomery", 254, [](Builder& c, const Wires& in) { return fq_inverse_montgomery(c, in); });
}
uint32_t build_fq_sqrt(Builder& b) {
  return b.build_root("fq_sqrt_montgomery", 254, [](Builder& c, const Wires& in) { return fq_sqrt_montgomery(c, in); });
}
uint32_t build_fq2_sqrt(Builder& b) {
  return b.build_root("fq2_sqrt_general", 508, [](Builder& c, const Wires& in) {
    return to_wires(fq2_sqrt_general_montgomery(c, fq2_from_wires(in.data())));
  });
}
uint32_t build_g1_add(Builder& b) {
  return b.build_root("g1_add", 1524, [](Builder& c, const Wires& in) {
    return to_wires(g1_add_montgomery(c, g1_from_wires(in.data()), g1_from_wires(in.data() + 762)));
  });
}
uint32_t build_g1_msm1(Builder& b, const host::G1Affine& base) {
  return b.build_root("g1_msm1", 254, [base](Builder& c, const Wires& in) {
    return to_wires(g1_msm_with_constant_bases_montgomery(c, {in}, {host::g1_from_affine(base)}));
  });
}
uint32_t build_fq12_square(Builder& b) {
  return b.build_root("fq12_square", 3048, [](Builder& c, const Wires& in) {
    return to_wires(fq12_square_montgomery(c, fq12_from_wires(in.data())));
  });
}
uint32_t build_fq12_cyclotomic_square(Builder& b) {
  return b.build_root("fq12_cyclotomic_square", 3048, [](Builder& c, const Wires& in) {
    return to_wires(fq12_cyclotomic_square_montgomery(c, fq12_from_wires(in.data())));
  });
}
uint32_t build_fq12_inverse(Builder& b) {
  return b.build_root("fq12_inverse", 3048, [](Builder& c, const Wires& in) {
    return to_wires(fq12_inverse_montgomery(c, fq12_from_wires(in.data())));
  });
}
uint32_t build_fq12_frobenius(Builder& b, size_t i) {
  return b.build_root("fq12_frobenius" + std::to_string(i), 3048, [i](Builder& c, const Wires& in) {
    return to_wires(fq12_frobenius_montgomery(c, fq12_from_wires(in.data()), i));
  });
}
// single steps of the pairing layer as roots: small enough for the independent emission model (tests/golden)
uint32_t build_g2_double_step(Builder& b) {
  return b.build_root("g2_double_step", 1524, [](Builder& c, const Wires& in) {
    G2P r;
    Fq6 coeffs;
    double_in_place_circuit_montgomery(c, g2_from_wires(in.data()), r, coeffs);
    return concat(to_wires(r), to_wires(coeffs));
  });
}
uint32_t build_g2_add_step(Builder& b) {
  return b.build_root("g2_add_step", 3048, [](Builder& c, const Wires& in) {
    G2P r;
    Fq6 coeffs;
    add_in_place_montgomery(c, g2_from_wires(in.data()), g2_from_wires(in.data() + 1524), r, coeffs);
    return concat(to_wires(r), to_wires(coeffs));
  });
}
uint32_t build_g2_mul_by_char(Builder& b) {
  return b.build_root("g2_mul_by_char", 1524, [](Builder& c, const Wires& in) {
    return to_wires(mul_by_char_montgomery(c, g2_from_wires(in.data())));
  });
}
uint32_t build_ell(Builder& b) {
  return b.build_root("ell", 3048 + 1524 + 762, [](Builder& c, const Wires& in) {
    return to_wires(ell_montgomery(c, fq12_from_wires(in.data()), fq6_from_wires(in.data() + 3048), g1_from_wires(in.data() + 4572)));
  });
}
uint32_t build_ell_const(Builder& b) {
  // line coefficients (3 + 5u, 7 + 11u, 13 + 17u), standard form
  const host::Fp6 coeffs{{host::Fp::from_u64(3), host::Fp::from_u64(5)}, {host::Fp::from_u64(7), host::Fp::from_u64(11)},
                         {host::Fp::from_u64(13), host::Fp::from_u64(17)}};
  return b.build_root("ell_const", 3048 + 762, [coeffs](Builder& c, const Wires& in) {
    return to_wires(ell_by_constant_montgomery(c, fq12_from_wires(in.data()), coeffs, g1_from_wires(in.data() + 3048)));
  });
}
uint32_t build_g1_to_affine(Builder& b) {
  return b.build_root("g1_to_affine", 762, [](Builder& c, const Wires& in) {
    return to_wires(projective_to_affine_montgomery(c, g1_from_wires(in.data())));
  });
}
uint32_t build_decompress_g1(Builder& b) {
  return b.build_root("decompress_g1", 255, [](Builder& c, const Wires& in) {
    return to_wires(decompress_g1_from_compressed(c, slice(in, 0, 254), in[254]));
  });
}
uint32_t build_final_exponentiation(Builder& b) {
  return b.build_root("final_exponentiation", 3048, [](Builder& c, const Wires& in) {
    return to_wires(final_exponentiation_montgomery(c, fq12_from_wires(in.data())));
  });
}
uint32_t build_miller_loop_groth16(Builder& b, const host::G2Affine& q1, const host::G2Affine& q2) {
  return b.build_root("miller_loop_groth16", 3 * 762 + 1524, [q1, q2](Builder& c, const Wires& in) {
    return to_wires(multi_miller_loop_groth16_evaluate_montgomery_fast(
        c, g1_from_wires(in.data()), g1_from_wires(in.data() + 762), g1_from_wires(in.data() + 1524), q1, q2,
        g2_from_wires(in.data() + 2286)));
  });
}

}  // namespace gsv

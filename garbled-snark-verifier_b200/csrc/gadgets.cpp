// gadgets.cpp -- emission-order restatement of the reference's gadget library (see gadgets.h).
// Every function cites the Rust it follows.  Rust evaluates call arguments left to right;
// C++ does not, so every sequence of gadget calls below is written as explicit statements.
#include "gadgets.h"

#include <cassert>
#include <stdexcept>

namespace gsv {

static inline BigInt slice(const BigInt& v, size_t lo, size_t hi) { return BigInt(v.begin() + lo, v.begin() + hi); }
static inline BigInt concat(const BigInt& a, const BigInt& b) {
  BigInt r(a);
  r.insert(r.end(), b.begin(), b.end());
  return r;
}

// =================================================================== basic.rs
// basic.rs:7-16
Pair2 half_adder(Builder& c, Wire a, Wire b) {
  Wire result = c.issue_wire();
  Wire carry = c.issue_wire();
  c.add_gate(XOR, a, b, result);
  c.add_gate(AND, a, b, carry);
  return {result, carry};
}
// basic.rs:18-33
Pair2 full_adder(Builder& c, Wire a, Wire b, Wire cin) {
  Wire axc = c.issue_wire(), bxc = c.issue_wire(), result = c.issue_wire(), t = c.issue_wire(),
       carry = c.issue_wire();
  c.add_gate(XOR, a, cin, axc);
  c.add_gate(XOR, b, cin, bxc);
  c.add_gate(XOR, a, bxc, result);
  c.add_gate(AND, axc, bxc, t);
  c.add_gate(XOR, cin, t, carry);
  return {result, carry};
}
// basic.rs:35-47  (and_variant [true,false,false] == Ncimp, gate.rs:187-198)
Pair2 half_subtracter(Builder& c, Wire a, Wire b) {
  Wire result = c.issue_wire();
  Wire borrow = c.issue_wire();
  c.add_gate(XOR, a, b, result);
  c.add_gate(NCIMP, a, b, borrow);
  return {result, borrow};
}
// basic.rs:49-64
Pair2 full_subtracter(Builder& c, Wire a, Wire b, Wire cin) {
  Wire bxa = c.issue_wire(), bxc = c.issue_wire(), result = c.issue_wire(), t = c.issue_wire(),
       carry = c.issue_wire();
  c.add_gate(XOR, a, b, bxa);
  c.add_gate(XOR, b, cin, bxc);
  c.add_gate(XOR, bxa, cin, result);
  c.add_gate(AND, bxa, bxc, t);
  c.add_gate(XOR, cin, t, carry);
  return {result, carry};
}
// basic.rs:66-72  selector(a, b, s) = s ? a : b
Wire selector(Builder& c, Wire a, Wire b, Wire s) {
  Wire d = c.issue_wire(), f = c.issue_wire(), g = c.issue_wire();
  c.add_gate(NAND, a, s, d);
  c.add_gate(CIMP, s, b, f);  // and_variant(c, b, f, [true,false,true])
  c.add_gate(NAND, d, f, g);
  return g;
}

// =================================================================== bigint
BigInt bn_constant(size_t len, const U256& u) {
  auto bits = u.bits_le(len);
  BigInt w(len);
  for (size_t i = 0; i < len; i++) w[i] = bits[i] ? WIRE_TRUE : WIRE_FALSE;
  return w;
}

// bigint/add.rs:8-26
BigInt bn_add(Builder& c, const BigInt& a, const BigInt& b) {
  if (a.size() != b.size()) throw std::logic_error("bn_add: length mismatch");
  size_t n = a.size();
  return c.component("bigint::add", concat(a, b), n + 1, [n](Builder& c, const Wires& in) {
    const Wire* a = in.data();
    const Wire* b = in.data() + n;
    BigInt bits;
    bits.reserve(n + 1);
    Pair2 r = half_adder(c, a[0], b[0]);
    bits.push_back(r.first);
    Wire carry = r.second;
    for (size_t i = 1; i < n; i++) {
      Pair2 q = full_adder(c, a[i], b[i], carry);
      bits.push_back(q.first);
      carry = q.second;
    }
    bits.push_back(carry);
    return bits;
  });
}
// bigint/add.rs:28-36
BigInt bn_add_without_carry(Builder& c, const BigInt& a, const BigInt& b) {
  BigInt r = bn_add(c, a, b);
  r.pop_back();
  return r;
}
// bigint/add.rs:38-82
BigInt bn_add_constant(Builder& c, const BigInt& a, const U256& k) {
  if (k.is_zero()) throw std::logic_error("bn_add_constant: zero constant");
  size_t n = a.size();
  auto kb = k.bits_le(n);
  return c.component("bigint::add_constant|b=" + k.to_hex(), a, n + 1, [n, kb](Builder& c, const Wires& a) {
    size_t first_one = 0;
    while (!kb[first_one]) first_one++;
    BigInt bits;
    bits.reserve(n + 1);
    Wire carry = WIRE_DEAD;
    for (size_t i = 0; i < n; i++) {
      Wire a_i = a[i];
      if (i < first_one) {
        bits.push_back(a_i);
      } else if (i == first_one) {
        Wire w = c.issue_wire();
        c.add_gate(XOR, a_i, WIRE_TRUE, w);  // Gate::not_with_xor
        bits.push_back(w);
        carry = a_i;
      } else if (kb[i]) {
        Wire w1 = c.issue_wire();
        Wire w2 = c.issue_wire();
        c.add_gate(XNOR, a_i, carry, w1);
        c.add_gate(OR, a_i, carry, w2);
        bits.push_back(w1);
        carry = w2;
      } else {
        Wire w1 = c.issue_wire();
        Wire w2 = c.issue_wire();
        c.add_gate(XOR, a_i, carry, w1);
        c.add_gate(AND, a_i, carry, w2);
        bits.push_back(w1);
        carry = w2;
      }
    }
    bits.push_back(carry);
    return bits;
  });
}
// bigint/add.rs:84-92
BigInt bn_add_constant_without_carry(Builder& c, const BigInt& a, const U256& k) {
  BigInt r = bn_add_constant(c, a, k);
  r.pop_back();
  return r;
}
// bigint/add.rs:94-114
BigInt bn_sub(Builder& c, const BigInt& a, const BigInt& b) {
  if (a.size() != b.size()) throw std::logic_error("bn_sub: length mismatch");
  size_t n = a.size();
  return c.component("bigint::sub", concat(a, b), n + 1, [n](Builder& c, const Wires& in) {
    const Wire* a = in.data();
    const Wire* b = in.data() + n;
    BigInt bits;
    bits.reserve(n + 1);
    Pair2 r = half_subtracter(c, a[0], b[0]);
    bits.push_back(r.first);
    Wire borrow = r.second;
    for (size_t i = 1; i < n; i++) {
      Pair2 q = full_subtracter(c, a[i], b[i], borrow);
      borrow = q.second;
      bits.push_back(q.first);
    }
    bits.push_back(borrow);
    return bits;
  });
}
// bigint/add.rs:116-125
BigInt bn_sub_without_borrow(Builder& c, const BigInt& a, const BigInt& b) {
  size_t n = a.size();
  return c.component("bigint::sub_without_borrow", concat(a, b), n, [n](Builder& c, const Wires& in) {
    BigInt bits = bn_sub(c, slice(in, 0, n), slice(in, n, 2 * n));
    bits.pop_back();
    return bits;
  });
}
// bigint/add.rs:144-153 (pure re-wiring)
BigInt bn_half(const BigInt& a) {
  BigInt r(a.begin() + 1, a.end());
  r.push_back(WIRE_FALSE);
  return r;
}
// bigint/cmp.rs:10-22
BigInt bn_self_or_zero(Builder& c, const BigInt& a, Wire s) {
  size_t n = a.size();
  BigInt in(a);
  in.push_back(s);
  return c.component("bigint::self_or_zero", in, n, [n](Builder& c, const Wires& in) {
    BigInt bits(n);
    for (size_t i = 0; i < n; i++) {
      Wire w = c.issue_wire();
      c.add_gate(AND, in[i], in[n], w);
      bits[i] = w;
    }
    return bits;
  });
}
// bigint/cmp.rs:87-107
Wire bn_equal_zero(Builder& c, const BigInt& a) {
  size_t n = a.size();
  return c.component("bigint::equal_zero", a, 1, [n](Builder& c, const Wires& a) {
    if (n == 1) {
      Wire z = c.issue_wire();
      c.add_gate(XOR, a[0], WIRE_TRUE, z);
      return Wires{z};
    }
    Wire res = c.issue_wire();
    c.add_gate(XNOR, a[0], a[1], res);
    for (size_t i = 1; i < n; i++) {
      Wire next = c.issue_wire();
      c.add_gate(NCIMP, a[i], res, next);  // and_variant [true,false,false]
      res = next;
    }
    return Wires{res};
  })[0];
}
// bigint/cmp.rs:60-85
Wire bn_equal_constant(Builder& c, const BigInt& a, const U256& k) {
  size_t n = a.size();
  return c.component("bigint::equal_constant|b=" + k.to_hex(), a, 1, [n, k](Builder& c, const Wires& a) {
    if (k.is_zero()) return Wires{bn_equal_zero(c, a)};
    auto kb = k.bits_le(n);
    size_t one_ind = 0;
    while (!kb[one_ind]) one_ind++;
    Wire res = a[one_ind];
    for (size_t i = 0; i < n; i++) {
      if (i == one_ind) continue;
      Wire next = c.issue_wire();
      // and_variant(a_i, res, new_res, [!b_i, false, false])
      c.add_gate(kb[i] ? AND : NCIMP, a[i], res, next);
      res = next;
    }
    return Wires{res};
  })[0];
}
// bigint/cmp.rs:109-129
Wire bn_greater_than(Builder& c, const BigInt& a, const BigInt& b) {
  size_t n = a.size();
  return c.component("bigint::greater_than", concat(a, b), 1, [n](Builder& c, const Wires& in) {
    BigInt a = slice(in, 0, n);
    BigInt not_b(n);
    for (size_t i = 0; i < n; i++) {
      Wire w = c.issue_wire();
      c.add_gate(XOR, in[n + i], WIRE_TRUE, w);
      not_b[i] = w;
    }
    BigInt sum = bn_add(c, a, not_b);
    return Wires{sum.back()};
  })[0];
}
// bigint/cmp.rs:131-151
Wire bn_less_than_constant(Builder& c, const BigInt& a, const U256& k) {
  size_t n = a.size();
  return c.component("bigint::less_than_constant|b=" + k.to_hex(), a, 1, [n, k](Builder& c, const Wires& a) {
    BigInt not_a(n);
    for (size_t i = 0; i < n; i++) {
      Wire w = c.issue_wire();
      c.add_gate(XOR, a[i], WIRE_TRUE, w);
      not_a[i] = w;
    }
    BigInt sum = bn_add_constant(c, not_a, k);
    return Wires{sum.back()};
  })[0];
}
// bigint/cmp.rs:153-169
BigInt bn_select(Builder& c, const BigInt& a, const BigInt& b, Wire s) {
  if (a.size() != b.size()) throw std::logic_error("bn_select: length mismatch");
  size_t n = a.size();
  BigInt in = concat(a, b);
  in.push_back(s);
  return c.component("bigint::select", in, n, [n](Builder& c, const Wires& in) {
    BigInt bits(n);
    for (size_t i = 0; i < n; i++) bits[i] = selector(c, in[i], in[n + i], in[2 * n]);
    return bits;
  });
}

// bigint/mul.rs:8-13
static bool use_karatsuba(size_t len) { return len == 21 ? false : len > 19; }

// bigint/mul.rs:19-55
BigInt bn_mul_naive(Builder& c, const BigInt& a, const BigInt& b) {
  if (a.size() != b.size()) throw std::logic_error("bn_mul_naive: length mismatch");
  size_t len = a.size();
  return c.component("bigint::mul_naive", concat(a, b), 2 * len, [len](Builder& c, const Wires& in) {
    const Wire* a = in.data();
    const Wire* b = in.data() + len;
    BigInt result(2 * len, WIRE_FALSE);
    for (size_t i = 0; i < len; i++) {
      BigInt add0(result.begin() + i, result.begin() + i + len);
      BigInt add1;
      add1.reserve(len);
      for (size_t j = 0; j < len; j++) {
        Wire w = c.issue_wire();
        c.add_gate(AND, a[j], b[i], w);
        add1.push_back(w);
      }
      BigInt sum = bn_add(c, add0, add1);
      for (size_t j = 0; j <= len; j++) result[i + j] = sum[j];
    }
    return result;
  });
}

// bigint/mul.rs:57-183
BigInt bn_mul_karatsuba(Builder& c, const BigInt& a, const BigInt& b) {
  if (a.size() != b.size()) throw std::logic_error("bn_mul_karatsuba: length mismatch");
  size_t len = a.size();
  return c.component("bigint::mul_karatsuba", concat(a, b), 2 * len, [len](Builder& c, const Wires& in) {
    BigInt a = slice(in, 0, len), b = slice(in, len, 2 * len);
    if (len < 5) return bn_mul_naive(c, a, b);
    BigInt result(2 * len, WIRE_FALSE);
    size_t len_0 = len / 2, len_1 = (len + 1) / 2;
    BigInt a_0 = slice(a, 0, len_0), a_1 = slice(a, len_0, len);
    BigInt b_0 = slice(b, 0, len_0), b_1 = slice(b, len_0, len);
    BigInt sq_0 = use_karatsuba(len_0) ? bn_mul_karatsuba(c, a_0, b_0) : bn_mul_naive(c, a_0, b_0);
    BigInt sq_1 = use_karatsuba(len_1) ? bn_mul_karatsuba(c, a_1, b_1) : bn_mul_naive(c, a_1, b_1);
    BigInt ext_a0 = a_0, ext_b0 = b_0, ext_sq0 = sq_0;
    if (len_0 < len_1) {
      ext_a0.push_back(WIRE_FALSE);
      ext_b0.push_back(WIRE_FALSE);
      ext_sq0.push_back(WIRE_FALSE);
      ext_sq0.push_back(WIRE_FALSE);
    }
    BigInt sum_a = bn_add(c, ext_a0, a_1);
    BigInt sum_b = bn_add(c, ext_b0, b_1);
    BigInt sq_sum = bn_add(c, ext_sq0, sq_1);
    sq_sum.push_back(WIRE_FALSE);
    BigInt sum_mul =
        use_karatsuba(sum_a.size()) ? bn_mul_karatsuba(c, sum_a, sum_b) : bn_mul_naive(c, sum_a, sum_b);
    BigInt cross_full = bn_sub_without_borrow(c, sum_mul, sq_sum);
    BigInt cross = slice(cross_full, 0, len + 1);
    for (size_t i = 0; i < 2 * len_0; i++) result[i] = sq_0[i];
    BigInt segment = slice(result, len_0, len_0 + len + 1);
    BigInt new_segment = bn_add(c, segment, cross);
    for (size_t i = 0; i < len + 2; i++) result[len_0 + i] = new_segment[i];
    BigInt segment2 = slice(result, 2 * len_0, 2 * len);
    BigInt new_segment2 = bn_add(c, segment2, sq_1);
    for (size_t i = 0; i < 2 * len_1; i++) result[2 * len_0 + i] = new_segment2[i];
    return result;
  });
}

// bigint/mul.rs:185-206
BigInt bn_mul(Builder& c, const BigInt& a, const BigInt& b) {
  size_t len = a.size();
  if (len < 5) return bn_mul_naive(c, a, b);
  if (len > 4000) throw std::logic_error("bn_mul: too long");
  return use_karatsuba(len) ? bn_mul_karatsuba(c, a, b) : bn_mul_naive(c, a, b);
}

// bigint/mul.rs:208-239
BigInt bn_mul_by_constant(Builder& c, const BigInt& a, const U256& k) {
  size_t len = a.size();
  auto kb = k.bits_le(len);
  return c.component("bigint::mul_by_constant|c=" + k.to_hex(), a, 2 * len, [len, kb](Builder& c, const Wires& a) {
    BigInt acc(2 * len, WIRE_FALSE);
    for (size_t i = 0; i < len; i++) {
      if (!kb[i]) continue;
      BigInt addw(acc.begin() + i, acc.begin() + i + len);
      BigInt nb = bn_add(c, a, addw);  // note operand order: (a, acc slice)
      for (size_t j = 0; j <= len; j++) acc[i + j] = nb[j];
    }
    return acc;
  });
}

// bigint/mul.rs:241-329
BigInt bn_mul_by_constant_modulo_power_two(Builder& c, const BigInt& a, const U256& k, size_t power) {
  size_t len = a.size();
  auto kb = k.bits_le(len);
  std::string kh = k.to_hex();
  return c.component(
      "bigint::mul_by_constant_modulo_power_two|c=" + kh + "|power=" + std::to_string(power), a, power,
      [len, kb, kh, power](Builder& c, const Wires& a) {
        const size_t PER_CHUNK = 8;
        if (!(power < 2 * len)) throw std::logic_error("power must be < 2*len");
        std::vector<size_t> ones;
        for (size_t i = 0; i < len; i++)
          if (i < power && kb[i]) ones.push_back(i);
        BigInt result(power, WIRE_FALSE);
        if (ones.empty()) return result;
        for (size_t chunk_idx = 0; chunk_idx * PER_CHUNK < ones.size(); chunk_idx++) {
          std::vector<size_t> chunk(ones.begin() + chunk_idx * PER_CHUNK,
                                    ones.begin() + std::min(ones.size(), (chunk_idx + 1) * PER_CHUNK));
          BigInt prev = result;
          // reference key: ("mul_by_const_mod_2p", a_len, power, chunk_idx); the constant is added
          // here because the body (the chunk's bit positions) depends on it.
          std::string key = "mul_by_const_mod_2p|a_len=" + std::to_string(len) + "|power=" +
                            std::to_string(power) + "|chunk_idx=" + std::to_string(chunk_idx) + "|c=" + kh;
          result = c.component(key, concat(a, prev), power, [len, power, chunk](Builder& c, const Wires& in) {
            BigInt a = slice(in, 0, len);
            BigInt res = slice(in, len, len + power);
            for (size_t i : chunk) {
              size_t nb = std::min(power - i, len);
              if (nb == 0) continue;
              BigInt a_slice = slice(a, 0, nb);
              BigInt addw = slice(res, i, i + nb);
              BigInt new_bits = bn_add(c, a_slice, addw);
              if (i + nb < power) {
                for (size_t j = 0; j <= nb; j++) res[i + j] = new_bits[j];
              } else {
                for (size_t j = 0; j < nb; j++) res[i + j] = new_bits[j];
              }
            }
            return res;
          });
        }
        return result;
      });
}

// =================================================================== fp254impl.rs / fq.rs
const FqConsts& FqConsts::get() {
  static const FqConsts k = [] {
    FqConsts f;
    // fq.rs:57-62
    f.p = U256::from_dec("21888242871839275222246405745257275088696311157297823662689037894645226208583");
    f.m_inv = U256::from_dec("4759646384140481320982610724935209484903937857060724391493050186936685796471");
    f.r_inv = U256::from_dec("18289368484950178621272022062020525048389989670507786348948026221581485535495");
    // fp254impl.rs:23-24
    f.r = U256::from_dec("28948022309329048855892746252171976963317496166410141009864396001978282409984");
    f.not_mod = sub(f.r, f.p);  // 2^254 - p, fp254impl.rs:59-63
    // fq.rs:65-75: Fq(1)/Fq(2), Fq(1)/Fq(3), Fq(2)/Fq(3) as integers
    f.half_mod = invmod(U256(2), f.p);
    f.third = invmod(U256(3), f.p);
    f.two_third = mulmod(U256(2), f.third, f.p);
    return f;
  }();
  return k;
}

// fp254impl.rs:95-113
Fq fq_add(Builder& c, const Fq& a, const Fq& b) {
  return c.component("fp254::add", concat(a, b), FQ_BITS, [](Builder& c, const Wires& in) {
    const FqConsts& K = FqConsts::get();
    BigInt a = slice(in, 0, FQ_BITS), b = slice(in, FQ_BITS, 2 * FQ_BITS);
    BigInt wires1 = bn_add(c, a, b);
    Wire u = wires1.back();
    wires1.pop_back();
    BigInt wires2 = bn_add_constant(c, wires1, K.not_mod);
    wires2.pop_back();
    Wire v = bn_less_than_constant(c, wires1, K.p);
    Wire s = c.issue_wire();
    c.add_gate(NCIMP, u, v, s);
    return bn_select(c, wires1, wires2, s);
  });
}
// fp254impl.rs:115-139
Fq fq_add_constant(Builder& c, const Fq& a, const U256& k) {
  return c.component("fp254::add_constant|b=" + k.to_hex(), a, FQ_BITS, [k](Builder& c, const Wires& a) {
    const FqConsts& K = FqConsts::get();
    if (k.is_zero()) return Wires(a);
    BigInt wires1 = bn_add_constant(c, a, k);
    Wire u = wires1.back();
    wires1.pop_back();
    BigInt wires2 = bn_add_constant(c, wires1, K.not_mod);
    wires2.pop_back();
    Wire v = bn_less_than_constant(c, wires1, K.p);
    Wire s = c.issue_wire();
    c.add_gate(NCIMP, u, v, s);
    return bn_select(c, wires1, wires2, s);
  });
}
// fp254impl.rs:152-166
Fq fq_neg(Builder& c, const Fq& a) {
  return c.component("fp254::neg", a, FQ_BITS, [](Builder& c, const Wires& a) {
    const FqConsts& K = FqConsts::get();
    BigInt not_a = c.issue_wires(FQ_BITS);
    for (size_t i = 0; i < FQ_BITS; i++) c.add_gate(XOR, a[i], WIRE_TRUE, not_a[i]);
    // Fq(1) - Fq(not_modulus) as a standard-form integer
    U256 k = submod(U256(1), K.not_mod, K.p);
    return fq_add_constant(c, not_a, k);
  });
}
// fp254impl.rs:142-149
Fq fq_sub(Builder& c, const Fq& a, const Fq& b) {
  return c.component("fp254::sub", concat(a, b), FQ_BITS, [](Builder& c, const Wires& in) {
    BigInt a = slice(in, 0, FQ_BITS), b = slice(in, FQ_BITS, 2 * FQ_BITS);
    Fq neg_b = fq_neg(c, b);
    return fq_add(c, a, neg_b);
  });
}
// fp254impl.rs:169-189
Fq fq_double(Builder& c, const Fq& a) {
  return c.component("fp254::double", a, FQ_BITS, [](Builder& c, const Wires& a) {
    const FqConsts& K = FqConsts::get();
    BigInt shifted(a);
    Wire u = shifted.back();
    shifted.pop_back();
    shifted.insert(shifted.begin(), WIRE_FALSE);
    BigInt wires2 = bn_add_constant(c, shifted, K.not_mod);
    wires2.pop_back();
    Wire v = bn_less_than_constant(c, shifted, K.p);
    Wire s = c.issue_wire();
    c.add_gate(NCIMP, u, v, s);
    return bn_select(c, shifted, wires2, s);
  });
}
// fp254impl.rs:192-201
Fq fq_half(Builder& c, const Fq& a) {
  return c.component("fp254::half", a, FQ_BITS, [](Builder& c, const Wires& a) {
    const FqConsts& K = FqConsts::get();
    Wire sel = a[0];
    BigInt w1 = bn_half(a);
    BigInt w2 = bn_add_constant_without_carry(c, w1, K.half_mod);
    return bn_select(c, w2, w1, sel);
  });
}
// fp254impl.rs:727-732
Fq fq_triple(Builder& c, const Fq& a) {
  return c.component("fp254::triple", a, FQ_BITS, [](Builder& c, const Wires& a) {
    Fq a2 = fq_double(c, a);
    return fq_add(c, a2, a);
  });
}
// fp254impl.rs:734-793
Fq fq_div6(Builder& c, const Fq& a) {
  return c.component("fp254::div6", a, FQ_BITS, [](Builder& c, const Wires& a) {
    const FqConsts& K = FqConsts::get();
    Fq half = fq_half(c, a);
    BigInt result = c.issue_wires(FQ_BITS);  // pre-issued, never written (credits 0)
    Wire r1 = WIRE_FALSE, r2 = WIRE_FALSE;
    for (size_t i = 0; i < FQ_BITS; i++) {
      size_t j = FQ_BITS - 1 - i;  // msb to lsb
      Wire r2_and_hj = c.issue_wire();
      c.add_gate(AND, r2, half[j], r2_and_hj);
      Wire result_wire = c.issue_wire();
      c.add_gate(OR, r1, r2_and_hj, result_wire);
      result[j] = result_wire;
      Wire new_r1 = c.issue_wire();
      c.add_gate(XOR, r2, result_wire, new_r1);
      r1 = new_r1;
      Wire new_r2 = c.issue_wire();
      c.add_gate(XOR, half[j], result_wire, new_r2);
      r2 = new_r2;
      Wire edge_case = c.issue_wire();
      c.add_gate(NIMP, result_wire, half[j], edge_case);
      Wire new_r1b = c.issue_wire();
      c.add_gate(XOR, r1, edge_case, new_r1b);
      r1 = new_r1b;
    }
    BigInt plus_third = bn_add_constant_without_carry(c, result, K.third);
    result = bn_select(c, plus_third, result, r2);
    BigInt plus_two_third = bn_add_constant_without_carry(c, result, K.two_third);
    return bn_select(c, plus_two_third, result, r1);
  });
}
// fp254impl.rs:303-331
Fq fq_montgomery_reduce(Builder& c, const BigInt& x) {
  if (x.size() != 2 * FQ_BITS) throw std::logic_error("montgomery_reduce: bad width");
  return c.component("fp254::montgomery_reduce", x, FQ_BITS, [](Builder& c, const Wires& x) {
    const FqConsts& K = FqConsts::get();
    BigInt x_low = slice(x, 0, 254), x_high = slice(x, 254, 508);
    BigInt q = bn_mul_by_constant_modulo_power_two(c, x_low, K.m_inv, 254);
    BigInt prod = bn_mul_by_constant(c, q, K.p);
    BigInt sub = slice(prod, 254, 508);
    Wire bound_check = bn_greater_than(c, sub, x_high);
    BigInt modulus = bn_constant(x_high.size(), K.p);
    BigInt subtract_if_too_much = bn_self_or_zero(c, modulus, bound_check);
    BigInt new_sub = bn_sub_without_borrow(c, sub, subtract_if_too_much);
    return bn_sub_without_borrow(c, x_high, new_sub);
  });
}
// fp254impl.rs:216-226
Fq fq_mul_montgomery(Builder& c, const Fq& a, const Fq& b) {
  BigInt m = bn_mul(c, a, b);
  return fq_montgomery_reduce(c, m);
}
// fp254impl.rs:285-287
Fq fq_square_montgomery(Builder& c, const Fq& a) { return fq_mul_montgomery(c, a, a); }

// =================================================================== fq2.rs
Wires to_wires(const Fq2& a) { return concat(a.c0, a.c1); }
Wires to_wires(const Fq6& a) { return concat(concat(to_wires(a.c0), to_wires(a.c1)), to_wires(a.c2)); }
Wires to_wires(const Fq12& a) { return concat(to_wires(a.c0), to_wires(a.c1)); }
Fq2 fq2_from_wires(const Wire* w) { return Fq2{Fq(w, w + 254), Fq(w + 254, w + 508)}; }
Fq6 fq6_from_wires(const Wire* w) {
  return Fq6{fq2_from_wires(w), fq2_from_wires(w + 508), fq2_from_wires(w + 1016)};
}
Fq12 fq12_from_wires(const Wire* w) { return Fq12{fq6_from_wires(w), fq6_from_wires(w + 1524)}; }

// fq2.rs:160-169
Fq2 fq2_add(Builder& c, const Fq2& a, const Fq2& b) {
  Fq c0 = fq_add(c, a.c0, b.c0);
  Fq c1 = fq_add(c, a.c1, b.c1);
  return {c0, c1};
}
// fq2.rs:190-202
Fq2 fq2_sub(Builder& c, const Fq2& a, const Fq2& b) {
  Fq c0 = fq_sub(c, a.c0, b.c0);
  Fq c1 = fq_sub(c, a.c1, b.c1);
  return {c0, c1};
}
// fq2.rs:204-212
Fq2 fq2_double(Builder& c, const Fq2& a) {
  Fq c0 = fq_double(c, a.c0);
  Fq c1 = fq_double(c, a.c1);
  return {c0, c1};
}
// fq2.rs:224-231: add(a, double(a))
Fq2 fq2_triple(Builder& c, const Fq2& a) {
  Fq2 a2 = fq2_double(c, a);
  return fq2_add(c, a, a2);
}
// fq2.rs:386-395
Fq2 fq2_div6(Builder& c, const Fq2& a) {
  Fq c0 = fq_div6(c, a.c0);
  Fq c1 = fq_div6(c, a.c1);
  return {c0, c1};
}
// fq2.rs:233-258
Fq2 fq2_mul_montgomery(Builder& c, const Fq2& a, const Fq2& b) {
  Fq a_sum = fq_add(c, a.c0, a.c1);
  Fq b_sum = fq_add(c, b.c0, b.c1);
  Fq a0_b0 = fq_mul_montgomery(c, a.c0, b.c0);
  Fq a1_b1 = fq_mul_montgomery(c, a.c1, b.c1);
  Fq sum_prod = fq_mul_montgomery(c, a_sum, b_sum);
  Fq c0 = fq_sub(c, a0_b0, a1_b1);
  Fq sum_a0b0_a1b1 = fq_add(c, a0_b0, a1_b1);
  Fq c1 = fq_sub(c, sum_prod, sum_a0b0_a1b1);
  return {c0, c1};
}
// fq2.rs:324-339  (xi = 9 + u)
Fq2 fq2_mul_by_nonresidue(Builder& c, const Fq2& a) {
  Fq a0_3 = fq_triple(c, a.c0);
  Fq a0_9 = fq_triple(c, a0_3);
  Fq a1_3 = fq_triple(c, a.c1);
  Fq a1_9 = fq_triple(c, a1_3);
  Fq c0 = fq_sub(c, a0_9, a.c1);
  Fq c1 = fq_add(c, a1_9, a.c0);
  return {c0, c1};
}

// =================================================================== fq6.rs
// fq6.rs:154-160
Fq6 fq6_add(Builder& c, const Fq6& a, const Fq6& b) {
  Fq2 c0 = fq2_add(c, a.c0, b.c0);
  Fq2 c1 = fq2_add(c, a.c1, b.c1);
  Fq2 c2 = fq2_add(c, a.c2, b.c2);
  return {c0, c1, c2};
}
// fq6.rs:170-176
Fq6 fq6_sub(Builder& c, const Fq6& a, const Fq6& b) {
  Fq2 c0 = fq2_sub(c, a.c0, b.c0);
  Fq2 c1 = fq2_sub(c, a.c1, b.c1);
  Fq2 c2 = fq2_sub(c, a.c2, b.c2);
  return {c0, c1, c2};
}
// fq6.rs:178-184
Fq6 fq6_double(Builder& c, const Fq6& a) {
  Fq2 c0 = fq2_double(c, a.c0);
  Fq2 c1 = fq2_double(c, a.c1);
  Fq2 c2 = fq2_double(c, a.c2);
  return {c0, c1, c2};
}
// fq6.rs:186-192
Fq6 fq6_div6(Builder& c, const Fq6& a) {
  Fq2 c0 = fq2_div6(c, a.c0);
  Fq2 c1 = fq2_div6(c, a.c1);
  Fq2 c2 = fq2_div6(c, a.c2);
  return {c0, c1, c2};
}
// fq6.rs:346-349
Fq6 fq6_mul_by_nonresidue(Builder& c, const Fq6& a) {
  Fq2 u = fq2_mul_by_nonresidue(c, a.c2);
  return {u, a.c0, a.c1};
}
// fq6.rs:194-260 (Toom-Cook-3)
Fq6 fq6_mul_montgomery(Builder& c, const Fq6& a, const Fq6& b) {
  const Fq2 &a_c0 = a.c0, &a_c1 = a.c1, &a_c2 = a.c2, &b_c0 = b.c0, &b_c1 = b.c1, &b_c2 = b.c2;
  Fq2 v0 = fq2_mul_montgomery(c, a_c0, b_c0);

  Fq2 wires_2 = fq2_add(c, a_c0, a_c2);
  Fq2 wires_3 = fq2_add(c, wires_2, a_c1);
  Fq2 wires_4 = fq2_sub(c, wires_2, a_c1);
  Fq2 wires_5 = fq2_double(c, a_c1);
  Fq2 wires_6 = fq2_double(c, a_c2);
  Fq2 wires_7 = fq2_double(c, wires_6);
  Fq2 wires_8 = fq2_add(c, a_c0, wires_5);
  Fq2 wires_9 = fq2_add(c, wires_8, wires_7);

  Fq2 wires_10 = fq2_add(c, b_c0, b_c2);
  Fq2 wires_11 = fq2_add(c, wires_10, b_c1);
  Fq2 wires_12 = fq2_sub(c, wires_10, b_c1);
  Fq2 wires_13 = fq2_double(c, b_c1);
  Fq2 wires_14 = fq2_double(c, b_c2);
  Fq2 wires_15 = fq2_double(c, wires_14);
  Fq2 wires_16 = fq2_add(c, b_c0, wires_13);
  Fq2 wires_17 = fq2_add(c, wires_16, wires_15);

  Fq2 v1 = fq2_mul_montgomery(c, wires_3, wires_11);
  Fq2 v2 = fq2_mul_montgomery(c, wires_4, wires_12);
  Fq2 v3 = fq2_mul_montgomery(c, wires_9, wires_17);
  Fq2 v4 = fq2_mul_montgomery(c, a_c2, b_c2);

  Fq2 v2_2 = fq2_double(c, v2);

  Fq2 v0_3 = fq2_triple(c, v0);
  Fq2 v1_3 = fq2_triple(c, v1);
  Fq2 v2_3 = fq2_triple(c, v2);
  Fq2 v4_3 = fq2_triple(c, v4);

  Fq2 v0_6 = fq2_double(c, v0_3);
  Fq2 v1_6 = fq2_double(c, v1_3);
  Fq2 v4_6 = fq2_double(c, v4_3);

  Fq2 v4_12 = fq2_double(c, v4_6);

  Fq2 wires_18 = fq2_sub(c, v0_3, v1_3);
  Fq2 wires_19 = fq2_sub(c, wires_18, v2);
  Fq2 wires_20 = fq2_add(c, wires_19, v3);
  Fq2 wires_21 = fq2_sub(c, wires_20, v4_12);
  Fq2 wires_22 = fq2_mul_by_nonresidue(c, wires_21);
  Fq2 r0 = fq2_add(c, wires_22, v0_6);

  Fq2 wires_23 = fq2_sub(c, v1_6, v0_3);
  Fq2 wires_24 = fq2_sub(c, wires_23, v2_2);
  Fq2 wires_25 = fq2_sub(c, wires_24, v3);
  Fq2 wires_26 = fq2_add(c, wires_25, v4_12);
  Fq2 wires_27 = fq2_mul_by_nonresidue(c, v4_6);
  Fq2 r1 = fq2_add(c, wires_26, wires_27);

  Fq2 wires_28 = fq2_sub(c, v1_3, v0_6);
  Fq2 wires_29 = fq2_add(c, wires_28, v2_3);
  Fq2 r2 = fq2_sub(c, wires_29, v4_6);

  return fq6_div6(c, Fq6{r0, r1, r2});
}

// =================================================================== fq12.rs
// fq12.rs:198-221
Fq12 fq12_mul_montgomery(Builder& c, const Fq12& a, const Fq12& b) {
  Wires in = concat(to_wires(a), to_wires(b));
  Wires out = c.component("fq12::mul_montgomery", in, 3048, [](Builder& c, const Wires& in) {
    Fq12 a = fq12_from_wires(in.data()), b = fq12_from_wires(in.data() + 3048);
    Fq6 a_sum = fq6_add(c, a.c0, a.c1);
    Fq6 b_sum = fq6_add(c, b.c0, b.c1);
    Fq6 a0_b0 = fq6_mul_montgomery(c, a.c0, b.c0);
    Fq6 a1_b1 = fq6_mul_montgomery(c, a.c1, b.c1);
    Fq6 sum_a0b0_a1b1 = fq6_add(c, a0_b0, a1_b1);
    Fq6 sum_prod = fq6_mul_montgomery(c, a_sum, b_sum);
    Fq6 a1_b1_nonres = fq6_mul_by_nonresidue(c, a1_b1);
    Fq6 c0 = fq6_add(c, a0_b0, a1_b1_nonres);
    Fq6 c1 = fq6_sub(c, sum_prod, sum_a0b0_a1b1);
    return to_wires(Fq12{c0, c1});
  });
  return fq12_from_wires(out.data());
}

// =================================================================== workload roots
uint32_t build_fq12_mul(Builder& b) {
  // tests/fq12_mul_e2e.rs:168-174; inputs a then b (:41-52)
  return b.build_root("fq12_mul", 6096, [](Builder& c, const Wires& in) {
    Fq12 x = fq12_from_wires(in.data()), y = fq12_from_wires(in.data() + 3048);
    return to_wires(fq12_mul_montgomery(c, x, y));
  });
}
uint32_t build_fq_mul(Builder& b) {
  return b.build_root("fq_mul", 508, [](Builder& c, const Wires& in) {
    return fq_mul_montgomery(c, slice(in, 0, 254), slice(in, 254, 508));
  });
}
uint32_t build_fq_add(Builder& b) {
  return b.build_root("fq_add", 508, [](Builder& c, const Wires& in) {
    return fq_add(c, slice(in, 0, 254), slice(in, 254, 508));
  });
}
uint32_t build_fq2_mul(Builder& b) {
  return b.build_root("fq2_mul", 1016, [](Builder& c, const Wires& in) {
    return to_wires(fq2_mul_montgomery(c, fq2_from_wires(in.data()), fq2_from_wires(in.data() + 508)));
  });
}
uint32_t build_fq6_mul(Builder& b) {
  return b.build_root("fq6_mul", 3048, [](Builder& c, const Wires& in) {
    return to_wires(fq6_mul_montgomery(c, fq6_from_wires(in.data()), fq6_from_wires(in.data() + 1524)));
  });
}
uint32_t build_bn_mul(Builder& b, size_t n) {
  return b.build_root("bn_mul" + std::to_string(n), 2 * n, [n](Builder& c, const Wires& in) {
    return bn_mul(c, slice(in, 0, n), slice(in, n, 2 * n));
  });
}
uint32_t build_gate_zoo(Builder& b) {
  // all 11 gate types on two inputs (tests/streaming_evaluate.rs all-gates case), one gate whose
  // output nobody reads (dead: consumes a gate index, emits nothing) and a constant-input gate.
  return b.build_root("gate_zoo", 2, [](Builder& c, const Wires& in) {
    Wires outs;
    for (int t = AND; t <= XNOR; t++) {
      Wire w = c.issue_wire();
      c.add_gate((uint8_t)t, in[0], in[1], w);
      outs.push_back(w);
    }
    Wire dead = c.issue_wire();
    c.add_gate(AND, in[0], in[1], dead);  // never read -> UNREACHABLE
    Wire n = c.issue_wire();
    c.add_gate(NOT, in[0], in[0], n);
    outs.push_back(n);
    Wire k = c.issue_wire();
    c.add_gate(OR, in[1], WIRE_TRUE, k);
    Wire k2 = c.issue_wire();
    c.add_gate(NIMP, k, WIRE_FALSE, k2);
    outs.push_back(k2);
    return outs;
  });
}
uint32_t build_fq_expr(Builder& b) {
  // tests/streaming_evaluate.rs:392-398: ((a^2) * b) + a
  return b.build_root("fq_expr", 508, [](Builder& c, const Wires& in) {
    Fq a = slice(in, 0, 254), bb = slice(in, 254, 508);
    Fq a2 = fq_square_montgomery(c, a);
    Fq a2b = fq_mul_montgomery(c, a2, bb);
    return fq_add(c, a2b, a);
  });
}

}  // namespace gsv

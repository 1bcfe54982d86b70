"""An INDEPENDENT second statement of the reference's gate emission order (TEST INFRASTRUCTURE).

Written from the reference's Rust gadgets (src/gadgets/basic.rs, bigint/{add,cmp,mul}.rs,
bn254/{fp254impl,fq,fq2,fq6,fq12,g1,pairing,final_exponentiation}.rs, groth16.rs) -- the
whole Groth16 verifier: multiplications, the binary Fp inverse and the tower inverses, squares, Frobenius maps,
exponentiation by a constant and the square roots, point decompression, G1 addition, the windowed constant-base
MSM, the G2 / line-evaluation steps, the Miller loop, the final exponentiation -- NOT from the product's C++ generator
(csrc/gadgets*.cpp, csrc/circuit.cpp), and with a different mechanism on purpose:

  * wires are global SSA ids in `issue_wire` order, gates go to one flat (type, a, b, c) stream in `add_gate`
    order -- no templates, no credit stacks;
  * liveness is the reference's rule stated globally: a wire is live iff it is
    read by a gate, passed as an input to a `#[component]` call, or is an output of the root; a gate whose
    output wire is not live keeps its gate index but gets c = UNREACHABLE.  (The product records per-component
    credit templates per output-liveness mask instead.)
  * a component body is recorded once per key and replayed by renumbering (numpy), which is what keeps
    Fq12::mul (20 M gates) fast enough for the CPU suite.

Constants are derived here (powers of the non-residue for Frobenius / twist coefficients, line coefficients of
constant G2 points and the alpha-beta target through the host arithmetic at the end of this file); the multiples of
a constant G1 base in the MSM tables follow arkworks' Jacobian formulas, restated from the published formulas (see
the MSM section).  Circuits too large to flatten are compared through structural_hash.py, which records the same
gadget functions into a component DAG.

`canonical_hash` renumbers wires by first live write, so the stream can be compared with the product's
`Program.flat_stream()` whatever the two sides' wire numbering; tests/golden/stream_hashes.json holds the
SHA-256 of the canonical streams produced here (make_stream_hashes.py).
"""
from __future__ import annotations

import hashlib

import numpy as np

AND, NAND, NIMP, IMP, NCIMP, CIMP, NOR, OR, XOR, XNOR, NOT = range(11)
FALSE, TRUE = 0, 1
DEAD = 0xFFFFFFFF

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583       # fq.rs:57-58
M_INV = 4759646384140481320982610724935209484903937857060724391493050186936685796471    # fq.rs:59-60
N = 254
NOT_MOD = (1 << N) - P


def and_variant(f):  # gate.rs:180-196
    return {(0, 0, 0): AND, (0, 0, 1): NAND, (0, 1, 0): NIMP, (0, 1, 1): IMP,
            (1, 0, 0): NCIMP, (1, 0, 1): CIMP, (1, 1, 0): NOR, (1, 1, 1): OR}[tuple(int(x) for x in f)]


def bits_of(v, n):
    return [(v >> i) & 1 for i in range(n)]


class Block:
    """A recorded component body in block-local numbering: 0 / 1 constants, 2 .. 2 + n_in inputs, then internals."""

    __slots__ = ("n_in", "n_local", "t", "a", "b", "c", "passed", "outs")


class Ctx:
    def __init__(self, n_inputs):
        self.n_in = n_inputs
        self.next = 2 + n_inputs
        self.chunks = []                       # finished numpy chunks (t, a, b, c)
        self.t, self.a, self.b, self.c = [], [], [], []
        self.passed = set()                    # wires passed into some component call
        self.memo = {}
        self.stack = []

    # ---- CircuitContext
    def issue(self):
        w = self.next
        self.next += 1
        return w

    def gate(self, typ, a, b, c):
        self.t.append(typ)
        self.a.append(a)
        self.b.append(b)
        self.c.append(c)

    def _flush(self):
        if self.t:
            self.chunks.append((np.array(self.t, np.uint8), np.array(self.a, np.int64), np.array(self.b, np.int64),
                                np.array(self.c, np.int64)))
            self.t, self.a, self.b, self.c = [], [], [], []

    def component(self, key, inputs, body):
        """with_named_child: every non-constant input wire is "passed" (+1 credit at entry,
        streaming_mode.rs:222-232); the body is recorded once per key and replayed by renumbering."""
        inputs = list(inputs)
        for w in inputs:
            if w >= 2:
                self.passed.add(w)
        key = (key, len(inputs))
        blk = self.memo.get(key)
        if blk is None:
            sub = Ctx(len(inputs))
            sub.memo = self.memo
            outs = body(sub, list(range(2, 2 + len(inputs))))
            sub._flush()
            blk = Block()
            blk.n_in, blk.n_local = len(inputs), sub.next
            if sub.chunks:
                blk.t = np.concatenate([c[0] for c in sub.chunks])
                blk.a = np.concatenate([c[1] for c in sub.chunks])
                blk.b = np.concatenate([c[2] for c in sub.chunks])
                blk.c = np.concatenate([c[3] for c in sub.chunks])
            else:
                blk.t = np.zeros(0, np.uint8)
                blk.a = blk.b = blk.c = np.zeros(0, np.int64)
            blk.passed = np.array(sorted(sub.passed), np.int64)
            blk.outs = list(outs)
            self.memo[key] = blk
        # replay: local id -> caller id
        m = np.empty(blk.n_local, np.int64)
        m[0], m[1] = 0, 1
        m[2:2 + blk.n_in] = inputs
        n_int = blk.n_local - 2 - blk.n_in
        m[2 + blk.n_in:] = np.arange(self.next, self.next + n_int)
        self.next += n_int
        self._flush()
        if blk.t.size:
            self.chunks.append((blk.t, m[blk.a], m[blk.b], m[blk.c]))
        for w in m[blk.passed].tolist():
            if w >= 2:
                self.passed.add(w)
        return [int(m[o]) for o in blk.outs]

    # ---- finish: liveness + canonical numbering
    def finish(self, outputs):
        self._flush()
        t = np.concatenate([c[0] for c in self.chunks])
        a = np.concatenate([c[1] for c in self.chunks])
        b = np.concatenate([c[2] for c in self.chunks])
        c = np.concatenate([c[3] for c in self.chunks])
        live = np.zeros(self.next, bool)
        live[a] = True
        live[b] = True
        live[np.array(sorted(self.passed), np.int64)] = True
        live[np.array([o for o in outputs if o >= 2], np.int64)] = True
        c = np.where(live[c], c, DEAD)
        return t, a, b, c, list(outputs), self.n_in


def canonical_stream(t, a, b, c, outputs, n_inputs):
    """The stream with wires renumbered by first live write (inputs keep 2 .. 2 + n_inputs): a compact SSA
    stream in the layout the CPU oracle walks.  Returns (type, a, b, c, outputs, n_wires)."""
    t = np.asarray(t, np.uint8)
    a, b, c = (np.asarray(x, np.int64) for x in (a, b, c))
    livew = c != DEAD
    hi = int(max(a.max(initial=0), b.max(initial=0), c[livew].max(initial=0), 2 + n_inputs)) + 1
    new = np.full(hi, -1, np.int64)
    new[:2 + n_inputs] = np.arange(2 + n_inputs)
    n_live = int(livew.sum())
    new[c[livew]] = 2 + n_inputs + np.arange(n_live)
    ca, cb = new[a], new[b]
    assert (ca >= 0).all() and (cb >= 0).all(), "stream reads a wire nobody wrote"
    cc = np.where(livew, new[np.where(livew, c, 0)], DEAD)
    outs = np.array([o if o < 2 else int(new[o]) for o in outputs], np.int64)
    return (t, ca.astype(np.uint32), cb.astype(np.uint32), cc.astype(np.uint32), outs.astype(np.uint32),
            2 + n_inputs + n_live)


def canonical_hash(t, a, b, c, outputs, n_inputs):
    """SHA-256 of the stream with wires renumbered by first live write (inputs keep 2 .. 2 + n_inputs)."""
    t = np.asarray(t, np.uint8)
    a, b, c = (np.asarray(x, np.int64) for x in (a, b, c))
    livew = c != DEAD
    hi = int(max(a.max(initial=0), b.max(initial=0), c[livew].max(initial=0), 2 + n_inputs)) + 1
    new = np.full(hi, -1, np.int64)
    new[:2 + n_inputs] = np.arange(2 + n_inputs)
    new[c[livew]] = 2 + n_inputs + np.arange(int(livew.sum()))
    ca, cb = new[a], new[b]
    assert (ca >= 0).all() and (cb >= 0).all(), "stream reads a wire nobody wrote"
    cc = np.where(livew, new[np.where(livew, c, 0)], DEAD)
    h = hashlib.sha256()
    h.update(np.int64(n_inputs).tobytes())
    h.update(t.tobytes())
    for x in (ca, cb, cc):
        h.update(x.astype(np.uint32).tobytes())
    outs = np.array([o if o < 2 else int(new[o]) for o in outputs], np.int64)
    h.update(outs.astype(np.uint32).tobytes())
    return h.hexdigest(), {"n_gates": int(t.size), "n_ciphertexts": int(((t < 8) & livew).sum()), "n_dead": int((~livew).sum())}


# ======================================================================================== basic.rs
def half_adder(x, a, b):
    r, cy = x.issue(), x.issue()
    x.gate(XOR, a, b, r)
    x.gate(AND, a, b, cy)
    return r, cy


def full_adder(x, a, b, c):
    axc, bxc, r, t, cy = (x.issue() for _ in range(5))
    x.gate(XOR, a, c, axc)
    x.gate(XOR, b, c, bxc)
    x.gate(XOR, a, bxc, r)
    x.gate(AND, axc, bxc, t)
    x.gate(XOR, c, t, cy)
    return r, cy


def half_subtracter(x, a, b):
    r, bo = x.issue(), x.issue()
    x.gate(XOR, a, b, r)
    x.gate(and_variant([1, 0, 0]), a, b, bo)
    return r, bo


def full_subtracter(x, a, b, c):
    bxa, bxc, r, t, cy = (x.issue() for _ in range(5))
    x.gate(XOR, a, b, bxa)
    x.gate(XOR, b, c, bxc)
    x.gate(XOR, bxa, c, r)
    x.gate(AND, bxa, bxc, t)
    x.gate(XOR, c, t, cy)
    return r, cy


def selector(x, a, b, c):
    d, f, g = x.issue(), x.issue(), x.issue()
    x.gate(NAND, a, c, d)
    x.gate(and_variant([1, 0, 1]), c, b, f)
    x.gate(NAND, d, f, g)
    return g


# ======================================================================================== bigint/add.rs
def bn_add(x, a, b):
    def body(x, w):
        n = len(w) // 2
        a, b = w[:n], w[n:]
        r, cy = half_adder(x, a[0], b[0])
        bits = [r]
        for i in range(1, n):
            r, cy = full_adder(x, a[i], b[i], cy)
            bits.append(r)
        return bits + [cy]
    assert len(a) == len(b)
    return x.component("bigint::add", a + b, body)


def bn_add_constant(x, a, k):
    assert k != 0
    n = len(a)

    def body(x, a):
        kb = bits_of(k, n)
        first = kb.index(1)
        bits, cy = [], None
        for i in range(n):
            if i < first:
                bits.append(a[i])
            elif i == first:
                w = x.issue()
                x.gate(XOR, a[i], TRUE, w)   # Gate::not_with_xor
                bits.append(w)
                cy = a[i]
            elif kb[i]:
                w1, w2 = x.issue(), x.issue()
                x.gate(XNOR, a[i], cy, w1)
                x.gate(OR, a[i], cy, w2)
                bits.append(w1)
                cy = w2
            else:
                w1, w2 = x.issue(), x.issue()
                x.gate(XOR, a[i], cy, w1)
                x.gate(AND, a[i], cy, w2)
                bits.append(w1)
                cy = w2
        return bits + [cy]
    return x.component(("bigint::add_constant", k), a, body)


def bn_sub(x, a, b):
    def body(x, w):
        n = len(w) // 2
        a, b = w[:n], w[n:]
        r, bo = half_subtracter(x, a[0], b[0])
        bits = [r]
        for i in range(1, n):
            r, bo = full_subtracter(x, a[i], b[i], bo)
            bits.append(r)
        return bits + [bo]
    return x.component("bigint::sub", a + b, body)


def bn_sub_without_borrow(x, a, b):
    n = len(a)
    return x.component("bigint::sub_without_borrow", a + b, lambda x, w: bn_sub(x, w[:n], w[n:])[:-1])


# ======================================================================================== bigint/cmp.rs
def self_or_zero(x, a, s):
    def body(x, w):
        out = []
        for ai in w[:-1]:
            o = x.issue()
            x.gate(AND, ai, w[-1], o)
            out.append(o)
        return out
    return x.component("bigint::self_or_zero", a + [s], body)


def greater_than(x, a, b):
    n = len(a)

    def body(x, w):
        a, b = w[:n], w[n:]
        nb = []
        for bi in b:
            o = x.issue()
            x.gate(XOR, bi, TRUE, o)
            nb.append(o)
        return [bn_add(x, a, nb)[-1]]
    return x.component("bigint::greater_than", a + b, body)[0]


def less_than_constant(x, a, k):
    def body(x, a):
        na = []
        for ai in a:
            o = x.issue()
            x.gate(XOR, ai, TRUE, o)
            na.append(o)
        return [bn_add_constant(x, na, k)[-1]]
    return x.component(("bigint::less_than_constant", k), a, body)[0]


def bn_select(x, a, b, s):
    n = len(a)
    return x.component("bigint::select", a + b + [s],
                       lambda x, w: [selector(x, w[i], w[n + i], w[2 * n]) for i in range(n)])


# ======================================================================================== bigint/mul.rs
def use_karatsuba(n):
    return n != 21 and n > 19


def mul_naive(x, a, b):
    n = len(a)

    def body(x, w):
        a, b = w[:n], w[n:]
        res = [FALSE] * (2 * n)
        for i, bi in enumerate(b):
            add0 = res[i:i + n]
            add1 = []
            for aj in a:
                o = x.issue()
                x.gate(AND, aj, bi, o)
                add1.append(o)
            res[i:i + n + 1] = bn_add(x, add0, add1)
        return res
    return x.component("bigint::mul_naive", a + b, body)


def mul_karatsuba(x, a, b):
    n = len(a)

    def body(x, w):
        a, b = w[:n], w[n:]
        if n < 5:
            return mul_naive(x, a, b)
        res = [FALSE] * (2 * n)
        l0, l1 = n // 2, (n + 1) // 2
        a0, a1, b0, b1 = a[:l0], a[l0:], b[:l0], b[l0:]
        sq0 = mul_karatsuba(x, a0, b0) if use_karatsuba(l0) else mul_naive(x, a0, b0)
        sq1 = mul_karatsuba(x, a1, b1) if use_karatsuba(l1) else mul_naive(x, a1, b1)
        ea0, eb0, esq0 = list(a0), list(b0), list(sq0)
        if l0 < l1:
            ea0.append(FALSE)
            eb0.append(FALSE)
            esq0 += [FALSE, FALSE]
        sum_a = bn_add(x, ea0, a1)
        sum_b = bn_add(x, eb0, b1)
        sq_sum = bn_add(x, esq0, sq1) + [FALSE]
        sum_mul = mul_karatsuba(x, sum_a, sum_b) if use_karatsuba(len(sum_a)) else mul_naive(x, sum_a, sum_b)
        cross = bn_sub_without_borrow(x, sum_mul, sq_sum)[:n + 1]
        res[:2 * l0] = sq0
        res[l0:l0 + n + 2] = bn_add(x, res[l0:l0 + n + 1], cross)
        res[2 * l0:] = bn_add(x, res[2 * l0:], sq1)[:2 * l1]
        return res
    return x.component("bigint::mul_karatsuba", a + b, body)


def bn_mul(x, a, b):
    n = len(a)
    if n < 5:
        return mul_naive(x, a, b)
    return mul_karatsuba(x, a, b) if use_karatsuba(n) else mul_naive(x, a, b)


def mul_by_constant(x, a, k):
    n = len(a)

    def body(x, a):
        acc = [FALSE] * (2 * n)
        for i, bit in enumerate(bits_of(k, n)):
            if bit:
                acc[i:i + n + 1] = bn_add(x, a, acc[i:i + n])
        return acc
    return x.component(("bigint::mul_by_constant", k), a, body)


def mul_by_constant_modulo_power_two(x, a, k, power):
    n = len(a)

    def body(x, a):
        ones = [i for i, bit in enumerate(bits_of(k, n)) if bit and i < power]
        res = [FALSE] * power
        for ci in range(0, len(ones), 8):
            chunk = ones[ci:ci + 8]

            def chunk_body(x, w, chunk=chunk):
                a, r = w[:n], list(w[n:])
                for i in chunk:
                    nb = min(power - i, n)
                    if nb == 0:
                        continue
                    new = bn_add(x, a[:nb], r[i:i + nb])
                    if i + nb < power:
                        r[i:i + nb + 1] = new
                    else:
                        r[i:i + nb] = new[:nb]
                return r
            res = x.component(("mul_by_const_mod_2p", n, power, ci // 8, k), a + res, chunk_body)
        return res
    return x.component(("bigint::mul_by_constant_modulo_power_two", k, power), a, body)


# ======================================================================================== bn254/fp254impl.rs
def _reduce_tail(x, w1, u):
    """add / add_constant / double share this tail: subtract p if the (N+1)-bit value is >= p."""
    w2 = bn_add_constant(x, w1, NOT_MOD)[:-1]
    v = less_than_constant(x, w1, P)
    s = x.issue()
    x.gate(and_variant([1, 0, 0]), u, v, s)
    return bn_select(x, w1, w2, s)


def fq_add(x, a, b):
    def body(x, w):
        w1 = bn_add(x, w[:N], w[N:])
        u = w1.pop()
        return _reduce_tail(x, w1, u)
    return x.component("fq::add", a + b, body)


def fq_add_constant(x, a, k):
    if k == 0:
        return x.component(("fq::add_constant", 0), a, lambda x, a: list(a))

    def body(x, a):
        w1 = bn_add_constant(x, a, k)
        u = w1.pop()
        return _reduce_tail(x, w1, u)
    return x.component(("fq::add_constant", k), a, body)


def fq_neg(x, a):
    def body(x, a):
        na = [x.issue() for _ in a]
        for o, ai in zip(na, a):
            x.gate(XOR, ai, TRUE, o)
        return fq_add_constant(x, na, (1 - NOT_MOD) % P)
    return x.component("fq::neg", a, body)


def fq_sub(x, a, b):
    return x.component("fq::sub", a + b, lambda x, w: fq_add(x, w[:N], fq_neg(x, w[N:])))


def fq_double(x, a):
    def body(x, a):
        sh = [FALSE] + a[:-1]
        return _reduce_tail(x, sh, a[-1])
    return x.component("fq::double", a, body)


def fq_half(x, a):
    def body(x, a):
        w1 = a[1:] + [FALSE]
        w2 = bn_add_constant(x, w1, (P + 1) // 2)[:-1]
        return bn_select(x, w2, w1, a[0])
    return x.component("fq::half", a, body)


def fq_triple(x, a):
    return x.component("fq::triple", a, lambda x, a: fq_add(x, fq_double(x, a), a))


def fq_div6(x, a):
    third, two_third = pow(3, -1, P), (2 * pow(3, -1, P)) % P

    def body(x, a):
        half = fq_half(x, a)
        result = [x.issue() for _ in range(N)]   # BigIntWires::from_ctx: issued, then replaced bit by bit
        r1 = r2 = FALSE
        for i in range(N):
            j = N - 1 - i
            r2h = x.issue()
            x.gate(AND, r2, half[j], r2h)
            rw = x.issue()
            x.gate(OR, r1, r2h, rw)
            result[j] = rw
            nr1 = x.issue()
            x.gate(XOR, r2, rw, nr1)
            r1 = nr1
            nr2 = x.issue()
            x.gate(XOR, half[j], rw, nr2)
            r2 = nr2
            edge = x.issue()
            x.gate(NIMP, rw, half[j], edge)
            nr1 = x.issue()
            x.gate(XOR, r1, edge, nr1)
            r1 = nr1
        plus1 = bn_add_constant(x, result, third)[:-1]
        result = bn_select(x, plus1, result, r2)
        plus2 = bn_add_constant(x, result, two_third)[:-1]
        return bn_select(x, plus2, result, r1)
    return x.component("fq::div6", a, body)


def montgomery_reduce(x, v):
    def body(x, v):
        lo, hi = v[:N], v[N:]
        q = mul_by_constant_modulo_power_two(x, lo, M_INV, N)
        sub = mul_by_constant(x, q, P)[N:2 * N]
        bound = greater_than(x, sub, hi)
        t = self_or_zero(x, bits_of(P, N), bound)     # constant wires of p (0 / 1 ids)
        ns = bn_sub_without_borrow(x, sub, t)
        return bn_sub_without_borrow(x, hi, ns)
    return x.component("fq::montgomery_reduce", v, body)


def fq_mul(x, a, b):
    return montgomery_reduce(x, bn_mul(x, a, b))


# ---- bigint helpers used by Fp::inverse (bigint/add.rs:128-187, bigint/cmp.rs:24-108)
def bn_double_without_overflow(x, a):   # #[bn_component], emits no gate: [FALSE, a0 .. a(n-2)]
    return x.component("bigint::double_without_overflow", a, lambda x, a: [FALSE] + a[:-1])


def bn_half(a):                         # plain function: a >> 1
    return a[1:] + [FALSE]


def self_or_zero_inv(x, a, s):
    def body(x, w):
        out = []
        for ai in w[:-1]:
            o = x.issue()
            x.gate(and_variant([0, 1, 0]), ai, w[-1], o)
            out.append(o)
        return out
    return x.component("bigint::self_or_zero_inv", a + [s], body)


def equal_zero(x, a):
    def body(x, a):
        if len(a) == 1:
            w = x.issue()
            x.gate(XOR, a[0], TRUE, w)
            return [w]
        res = x.issue()
        x.gate(XNOR, a[0], a[1], res)
        for ai in a[1:]:
            nxt = x.issue()
            x.gate(and_variant([1, 0, 0]), ai, res, nxt)
            res = nxt
        return [res]
    return x.component("bigint::equal_zero", a, body)[0]


def equal_constant(x, a, k):
    n = len(a)

    def body(x, a):
        if k == 0:
            return [equal_zero(x, a)]
        kb = bits_of(k, n)
        one = kb.index(1)
        res = a[one]
        for i, ai in enumerate(a):
            if i == one:
                continue
            nxt = x.issue()
            x.gate(and_variant([0 if kb[i] else 1, 0, 0]), ai, res, nxt)
            res = nxt
        return [res]
    return x.component(("bigint::equal_constant", k), a, body)[0]


def odd_part(x, a):                     # plain function (add.rs:155-187)
    n = len(a)
    sel = [a[0]] + [x.issue() for _ in range(n - 1)]
    for i in range(1, n):
        x.gate(OR, sel[i - 1], a[i], sel[i])
    k = [a[0]] + [x.issue() for _ in range(n - 1)]
    for i in range(1, n):
        x.gate(and_variant([1, 0, 0]), sel[i - 1], a[i], k[i])
    acc = list(a)
    for i in range(n):
        acc = bn_select(x, acc, bn_half(acc), sel[i])
    return acc, k


def fq_mul_by_constant(x, a, k):        # fp254impl.rs:252-272, k already in Montgomery form
    def body(x, a):
        if k == 0:
            return [FALSE] * N
        if k == (1 << N) % P:
            return list(a)
        return montgomery_reduce(x, mul_by_constant(x, a, k))
    return x.component(("fq::mul_by_constant_montgomery", k), a, body)


def fq_inverse(x, a):                   # fp254impl.rs:333-660
    PER = 4

    def iteration(cnt):
        def body(x, w):
            u, v, r, s, k = (list(w[j * N:(j + 1) * N]) for j in range(5))
            for _ in range(cnt):
                not_x1, not_x2 = u[0], v[0]
                x3 = greater_than(x, u, v)
                p2 = x.issue()
                x.gate(and_variant([0, 1, 0]), not_x1, not_x2, p2)
                p3, w2 = x.issue(), x.issue()
                x.gate(AND, not_x1, not_x2, w2)
                x.gate(AND, w2, x3, p3)
                p4 = x.issue()
                x.gate(NIMP, w2, x3, p4)
                # the four candidate updates
                u1, v1, r1, s1, k1 = bn_half(u), v, r, bn_double_without_overflow(x, s), bn_add_constant(x, k, 1)[:-1]
                u2, v2, r2, s2, k2 = u, bn_half(v), bn_double_without_overflow(x, r), s, bn_add_constant(x, k, 1)[:-1]
                u3 = bn_sub_without_borrow(x, u1, v2)
                v3 = v
                r3 = bn_add(x, r, s)[:-1]
                s3 = bn_double_without_overflow(x, s)
                k3 = bn_add_constant(x, k, 1)[:-1]
                u4 = u
                v4 = bn_sub_without_borrow(x, v2, u1)
                r4 = bn_double_without_overflow(x, r)
                s4 = bn_add(x, r, s)[:-1]
                k4 = bn_add_constant(x, k, 1)[:-1]
                new = []
                for c1, c2, c3, c4 in ((u1, u2, u3, u4), (v1, v2, v3, v4), (r1, r2, r3, r4), (s1, s2, s3, s4),
                                       (k1, k2, k3, k4)):
                    w1 = self_or_zero_inv(x, c1, not_x1)
                    w2_ = self_or_zero(x, c2, p2)
                    w3 = self_or_zero(x, c3, p3)
                    w4 = self_or_zero(x, c4, p4)
                    acc = bn_add(x, w1, w2_)[:-1]
                    acc = bn_add(x, acc, w3)[:-1]
                    new.append(bn_add(x, acc, w4)[:-1])
                v_is_one = equal_constant(x, v, 1)
                u, v, r, s, k = (bn_select(x, old, nw, v_is_one) for old, nw in zip((u, v, r, s, k), new))
            return u + v + r + s + k
        return body

    def by_even_chunk(cnt):
        def body(x, w):
            s, even = list(w[:N]), list(w[N:])
            for _ in range(cnt):
                hs, he = fq_half(x, s), fq_half(x, even)
                sel = equal_constant(x, even, 1)
                s = bn_select(x, s, hs, sel)
                even = bn_select(x, even, he, sel)
            return s + even
        return body

    def by_even(x, w):
        s, even = list(w[:N]), list(w[N:])
        for ci, start in enumerate(range(0, N, PER)):
            cnt = min(PER, N - start)
            r = x.component(("inverse::divide_result_by_even_part::chunk", ci), s + even, by_even_chunk(cnt))
            s, even = r[:N], r[N:]
        return s

    def by_2k_chunk(cnt):
        def body(x, w):
            s, k = list(w[:N]), list(w[N:])
            for _ in range(cnt):
                hs = fq_half(x, s)
                km1 = fq_add_constant(x, k, P - 1)
                sel = equal_constant(x, k, 0)
                s = bn_select(x, s, hs, sel)
                k = bn_select(x, k, km1, sel)
            return s + k
        return body

    def by_2k(x, w):
        s, k = list(w[:N]), list(w[N:])
        for start in range(0, 2 * N, PER):
            r = x.component("inverse::divide_result_by_2^k::chunk", s + k, by_2k_chunk(min(PER, 2 * N - start)))
            s, k = r[:N], r[N:]
        return s

    def body(x, a):
        odd, even = odd_part(x, a)
        u = bn_half(fq_neg(x, odd))
        state = u + odd + bits_of(1, N) + bits_of(2, N) + bits_of(1, N)     # u, v, r, s, k
        for start in range(0, 2 * N, PER):
            state = x.component("inverse_iteration", state, iteration(min(PER, 2 * N - start)))
        s, k = state[3 * N:4 * N], state[4 * N:]
        s = x.component("inverse::divide_result_by_even_part", s + even, by_even)
        return x.component("inverse::divide_result_by_2^k", s + k, by_2k)
    return x.component("fq::inverse", a, body)


def fq_inverse_montgomery(x, a):        # fp254impl.rs:676-686: inverse, then times R^3
    return fq_mul_by_constant(x, fq_inverse(x, a), pow(1 << N, 3, P))


# ======================================================================================== fq2 / fq6 / fq12
def fq2_map(f):
    return lambda x, a, *r: [f(x, a[0], *[q[0] for q in r]), f(x, a[1], *[q[1] for q in r])]


fq2_add, fq2_sub, fq2_double, fq2_div6 = fq2_map(fq_add), fq2_map(fq_sub), fq2_map(fq_double), fq2_map(fq_div6)


def fq2_triple(x, a):
    return fq2_add(x, a, fq2_double(x, a))


def fq2_mul(x, a, b):
    a_sum = fq_add(x, a[0], a[1])
    b_sum = fq_add(x, b[0], b[1])
    a0b0 = fq_mul(x, a[0], b[0])
    a1b1 = fq_mul(x, a[1], b[1])
    sum_prod = fq_mul(x, a_sum, b_sum)
    c0 = fq_sub(x, a0b0, a1b1)
    t = fq_add(x, a0b0, a1b1)
    return [c0, fq_sub(x, sum_prod, t)]


def fq2_mul_by_nonresidue(x, a):
    a0_9 = fq_triple(x, fq_triple(x, a[0]))
    a1_9 = fq_triple(x, fq_triple(x, a[1]))
    return [fq_sub(x, a0_9, a[1]), fq_add(x, a1_9, a[0])]


def fq6_map(f):
    return lambda x, a, *r: [f(x, a[i], *[q[i] for q in r]) for i in range(3)]


fq6_add, fq6_sub, fq6_div6 = fq6_map(fq2_add), fq6_map(fq2_sub), fq6_map(fq2_div6)


def fq6_mul(x, a, b):   # fq6.rs:194-260, Toom-Cook-3, statement order as in the reference
    a0, a1, a2 = a
    b0, b1, b2 = b
    v0 = fq2_mul(x, a0, b0)
    w2 = fq2_add(x, a0, a2)
    w3 = fq2_add(x, w2, a1)
    w4 = fq2_sub(x, w2, a1)
    w5 = fq2_double(x, a1)
    w6 = fq2_double(x, a2)
    w7 = fq2_double(x, w6)
    w8 = fq2_add(x, a0, w5)
    w9 = fq2_add(x, w8, w7)
    w10 = fq2_add(x, b0, b2)
    w11 = fq2_add(x, w10, b1)
    w12 = fq2_sub(x, w10, b1)
    w13 = fq2_double(x, b1)
    w14 = fq2_double(x, b2)
    w15 = fq2_double(x, w14)
    w16 = fq2_add(x, b0, w13)
    w17 = fq2_add(x, w16, w15)
    v1 = fq2_mul(x, w3, w11)
    v2 = fq2_mul(x, w4, w12)
    v3 = fq2_mul(x, w9, w17)
    v4 = fq2_mul(x, a2, b2)
    v2_2 = fq2_double(x, v2)
    v0_3 = fq2_triple(x, v0)
    v1_3 = fq2_triple(x, v1)
    v2_3 = fq2_triple(x, v2)
    v4_3 = fq2_triple(x, v4)
    v0_6 = fq2_double(x, v0_3)
    v1_6 = fq2_double(x, v1_3)
    v4_6 = fq2_double(x, v4_3)
    v4_12 = fq2_double(x, v4_6)
    w18 = fq2_sub(x, v0_3, v1_3)
    w19 = fq2_sub(x, w18, v2)
    w20 = fq2_add(x, w19, v3)
    w21 = fq2_sub(x, w20, v4_12)
    w22 = fq2_mul_by_nonresidue(x, w21)
    c0 = fq2_add(x, w22, v0_6)
    w23 = fq2_sub(x, v1_6, v0_3)
    w24 = fq2_sub(x, w23, v2_2)
    w25 = fq2_sub(x, w24, v3)
    w26 = fq2_add(x, w25, v4_12)
    w27 = fq2_mul_by_nonresidue(x, v4_6)
    c1 = fq2_add(x, w26, w27)
    w28 = fq2_sub(x, v1_3, v0_6)
    w29 = fq2_add(x, w28, v2_3)
    c2 = fq2_sub(x, w29, v4_6)
    return fq6_div6(x, [c0, c1, c2])


def fq6_mul_by_nonresidue(x, a):
    return [fq2_mul_by_nonresidue(x, a[2]), a[0], a[1]]


def _fq6_of(w):
    return [[w[(2 * i + j) * N:(2 * i + j + 1) * N] for j in range(2)] for i in range(3)]


def _flat6(v):
    return [w for f2 in v for f in f2 for w in f]


def fq12_mul(x, a, b):   # fq12.rs:198-221, #[component]
    def body(x, w):
        a0, a1, b0, b1 = (_fq6_of(w[k * 6 * N:(k + 1) * 6 * N]) for k in range(4))
        a_sum = fq6_add(x, a0, a1)
        b_sum = fq6_add(x, b0, b1)
        a0b0 = fq6_mul(x, a0, b0)
        a1b1 = fq6_mul(x, a1, b1)
        s = fq6_add(x, a0b0, a1b1)
        sum_prod = fq6_mul(x, a_sum, b_sum)
        nr = fq6_mul_by_nonresidue(x, a1b1)
        c0 = fq6_add(x, a0b0, nr)
        c1 = fq6_sub(x, sum_prod, s)
        return _flat6(c0) + _flat6(c1)
    return x.component("fq12::mul_montgomery", a + b, body)


def fq_exp_by_constant(x, a, e):        # fp254impl.rs:692-725, #[bn_component(offcircuit_args = "exp")]
    def body(x, a):
        if e == 0:
            return bits_of(1, N)
        if e == 1:
            return list(a)
        r = list(a)
        for bit in bin(e)[3:]:          # below the leading one, most significant first
            sq = fq_mul(x, r, r)
            r = fq_mul(x, a, sq) if bit == "1" else sq
        return r
    return x.component(("fq::exp_by_constant_montgomery", e), list(a), body)


def fq_sqrt(x, a):                      # fq.rs:291-299: a^((p + 1) / 4)
    return fq_exp_by_constant(x, a, (P + 1) // 4)


# ======================================================================================== squares, inverses, Frobenius maps
R_MONT = (1 << N) % P


def _f2mul(a, b):      # host arithmetic in Fq2 = Fq[u] / (u^2 + 1), for the Frobenius constants only
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _f2mul(r, a)
        a = _f2mul(a, a)
        e >>= 1
    return r


XI = (9, 1)            # the sextic non-residue 9 + u
FROB_FP2_C1 = [1, P - 1]
FROB_FP6_C1 = [_f2pow(XI, (P ** i - 1) // 3) for i in range(6)]
FROB_FP6_C2 = [_f2pow(XI, (2 * P ** i - 2) // 3) for i in range(6)]
FROB_FP12_C1 = [_f2pow(XI, (P ** i - 1) // 6) for i in range(12)]


def _mont2(c):
    return (c[0] * R_MONT % P, c[1] * R_MONT % P)


fq2_half = fq2_map(fq_half)
fq2_neg = fq2_map(fq_neg)


def fq2_square(x, a):                   # fq2.rs:341-354
    s = fq_add(x, a[0], a[1])
    d = fq_sub(x, a[0], a[1])
    a0a1 = fq_mul(x, a[0], a[1])
    c0 = fq_mul(x, s, d)
    return [c0, fq_double(x, a0a1)]


def fq2_inverse(x, a):                  # fq2.rs:356-372, #[component]
    def body(x, w):
        a0, a1 = list(w[:N]), list(w[N:])
        a0s = fq_mul(x, a0, a0)
        a1s = fq_mul(x, a1, a1)
        norm = fq_add(x, a0s, a1s)
        inv = fq_inverse_montgomery(x, norm)
        c0 = fq_mul(x, a0, inv)
        na1 = fq_neg(x, a1)
        return c0 + fq_mul(x, na1, inv)
    r = x.component("fq2::inverse_montgomery", a[0] + a[1], body)
    return [r[:N], r[N:]]


def fq2_mul_by_constant(x, a, k):       # fq2.rs:257-280; k = (c0, c1) as the caller passes it (Montgomery form)
    if k == (1, 0):
        return [list(a[0]), list(a[1])]
    a_sum = fq_add(x, a[0], a[1])
    a0b0 = fq_mul_by_constant(x, a[0], k[0])
    a1b1 = fq_mul_by_constant(x, a[1], k[1])
    sms = fq_mul_by_constant(x, a_sum, (k[0] + k[1]) % P)
    c0 = fq_sub(x, a0b0, a1b1)
    t = fq_add(x, a0b0, a1b1)
    return [c0, fq_sub(x, sms, t)]


def fq2_frobenius(x, a, i):             # fq2.rs:374-384
    return [list(a[0]), fq_mul_by_constant(x, a[1], FROB_FP2_C1[i % 2] * R_MONT % P)]


fq6_double = fq6_map(fq2_double)
fq6_neg = fq6_map(fq2_neg)


def fq6_square(x, a):                   # fq6.rs:421-448
    a0, a1, a2 = a
    s0 = fq2_square(x, a0)
    w1 = fq2_add(x, a0, a2)
    w2 = fq2_add(x, w1, a1)
    w3 = fq2_sub(x, w1, a1)
    s1 = fq2_square(x, w2)
    s2 = fq2_square(x, w3)
    w4 = fq2_mul(x, a1, a2)
    s3 = fq2_double(x, w4)
    s4 = fq2_square(x, a2)
    w5 = fq2_add(x, s1, s2)
    t1 = fq2_half(x, w5)
    w6 = fq2_mul_by_nonresidue(x, s3)
    c0 = fq2_add(x, s0, w6)
    w7 = fq2_mul_by_nonresidue(x, s4)
    w8 = fq2_sub(x, s1, s3)
    w9 = fq2_sub(x, w8, t1)
    c1 = fq2_add(x, w9, w7)
    w10 = fq2_sub(x, t1, s0)
    return [c0, c1, fq2_sub(x, w10, s4)]


def fq6_inverse(x, r):                  # fq6.rs:450-487
    a, b, c = r
    a_sq = fq2_square(x, a)
    b_sq = fq2_square(x, b)
    c_sq = fq2_square(x, c)
    ab = fq2_mul(x, a, b)
    ac = fq2_mul(x, a, c)
    bc = fq2_mul(x, b, c)
    bc_beta = fq2_mul_by_nonresidue(x, bc)
    t0 = fq2_sub(x, a_sq, bc_beta)                  # a^2 - bc beta
    c_sq_beta = fq2_mul_by_nonresidue(x, c_sq)
    t1 = fq2_sub(x, c_sq_beta, ab)                  # c^2 beta - ab
    t2 = fq2_sub(x, b_sq, ac)                       # b^2 - ac
    w1 = fq2_mul(x, t1, c)
    w2 = fq2_mul(x, t2, b)
    w12 = fq2_add(x, w1, w2)
    w3 = fq2_mul_by_nonresidue(x, w12)
    w4 = fq2_mul(x, a, t0)
    norm = fq2_add(x, w4, w3)
    inv = fq2_inverse(x, norm)
    return [fq2_mul(x, t0, inv), fq2_mul(x, t1, inv), fq2_mul(x, t2, inv)]


def fq6_frobenius(x, a, i):             # fq6.rs:489-514
    f0 = fq2_frobenius(x, a[0], i)
    f1 = fq2_frobenius(x, a[1], i)
    f2 = fq2_frobenius(x, a[2], i)
    f1u = fq2_mul_by_constant(x, f1, _mont2(FROB_FP6_C1[i % 6]))
    f2u = fq2_mul_by_constant(x, f2, _mont2(FROB_FP6_C2[i % 6]))
    return [f0, f1u, f2u]


def fq6_mul_by_constant_fq2(x, a, k):   # fq6.rs:334-344
    return [fq2_mul_by_constant(x, a[j], k) for j in range(3)]


def fq12_square(x, a):                  # fq12.rs:311-324, #[component]
    def body(x, w):
        a0, a1 = _fq6_of(w[:6 * N]), _fq6_of(w[6 * N:])
        w1 = fq6_add(x, a0, a1)
        w2 = fq6_mul_by_nonresidue(x, a1)
        w3 = fq6_add(x, a0, w2)
        w4 = fq6_mul(x, a0, a1)
        w5 = fq6_mul(x, w1, w3)
        w6 = fq6_mul_by_nonresidue(x, w4)
        w7 = fq6_add(x, w4, w6)
        c0 = fq6_sub(x, w5, w7)
        return _flat6(c0) + _flat6(fq6_double(x, w4))
    return x.component("fq12::square_montgomery", list(a), body)


def fq12_cyclotomic_square(x, a):       # fq12.rs:326-392, plain function
    a0, a1 = _fq6_of(a[:6 * N]), _fq6_of(a[6 * N:])
    c0, c1, c2, c3, c4, c5 = a0[0], a0[1], a0[2], a1[0], a1[1], a1[2]

    def fp4_square(p, q, yb_of, other):
        xy = fq2_mul(x, p, q)
        x_plus_y = fq2_add(x, p, q)
        y_beta = fq2_mul_by_nonresidue(x, yb_of)
        x_plus_y_beta = fq2_add(x, other, y_beta)
        xy_beta = fq2_mul_by_nonresidue(x, xy)
        w1 = fq2_mul(x, x_plus_y, x_plus_y_beta)
        w2 = fq2_add(x, xy, xy_beta)
        return fq2_sub(x, w1, w2), fq2_double(x, xy)
    t0, t1 = fp4_square(c0, c4, c4, c0)
    t2, t3 = fp4_square(c2, c3, c2, c3)
    t4, t5 = fp4_square(c1, c5, c5, c1)

    def comb_sub(t, c):
        return fq2_add(x, fq2_double(x, fq2_sub(x, t, c)), t)

    def comb_add(t, c):
        return fq2_add(x, fq2_double(x, fq2_add(x, t, c)), t)
    z0 = comb_sub(t0, c0)
    z4 = comb_sub(t2, c1)
    z3 = comb_sub(t4, c2)
    t5_beta = fq2_mul_by_nonresidue(x, t5)
    z2 = comb_add(t5_beta, c3)
    z1 = comb_add(t1, c4)
    z5 = comb_add(t3, c5)
    return _flat6([z0, z4, z3]) + _flat6([z2, z1, z5])


def fq12_inverse(x, a):                 # fq12.rs:413-428, #[component]
    def body(x, w):
        a0, a1 = _fq6_of(w[:6 * N]), _fq6_of(w[6 * N:])
        a0s = fq6_square(x, a0)
        a1s = fq6_square(x, a1)
        a1sb = fq6_mul_by_nonresidue(x, a1s)
        norm = fq6_sub(x, a0s, a1sb)
        inv = fq6_inverse(x, norm)
        c0 = fq6_mul(x, a0, inv)
        na1 = fq6_neg(x, a1)
        return _flat6(c0) + _flat6(fq6_mul(x, inv, na1))
    return x.component("fq12::inverse_montgomery", list(a), body)


def fq12_frobenius(x, a, i):            # fq12.rs:430-442
    a0, a1 = _fq6_of(a[:6 * N]), _fq6_of(a[6 * N:])
    f0 = fq6_frobenius(x, a0, i)
    f1 = fq6_frobenius(x, a1, i)
    return _flat6(f0) + _flat6(fq6_mul_by_constant_fq2(x, f1, _mont2(FROB_FP12_C1[i % 12])))


# ======================================================================================== pairing layer (single steps)
def _f2inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


G2_COEFF_B = _f2mul((3, 0), _f2inv(XI))             # b' = 3 / (9 + u), the twist's coefficient
TWIST_MUL_BY_Q_X = _f2pow(XI, (P - 1) // 3)
TWIST_MUL_BY_Q_Y = _f2pow(XI, (P - 1) // 2)


def fq2_mul_by_fq(x, a, b):             # fq2.rs:282-291
    return [fq_mul(x, a[0], b), fq_mul(x, a[1], b)]


def fq2_mul_constant_by_fq(x, k, b):    # fq2.rs:307-322, #[component(offcircuit_args = "a")]; k in standard form
    def body(x, b):
        return fq_mul_by_constant(x, b, k[0] * R_MONT % P) + fq_mul_by_constant(x, b, k[1] * R_MONT % P)
    r = x.component(("fq2::mul_constant_by_fq_montgomery", k), list(b), body)
    return [r[:N], r[N:]]


def fq2_add_constant(x, a, k):          # fq2.rs:170-177
    return [fq_add_constant(x, a[0], k[0]), fq_add_constant(x, a[1], k[1])]


def fq6_mul_by_fq2(x, a, b):            # fq6.rs:326-332
    return [fq2_mul(x, a[j], b) for j in range(3)]


def _mul_by_01(x, a, c0, mul_c1, c0_plus_c1):
    """fq6.rs:351-410: the two sparse products differ only in how `* c1` and `c0 + c1` are formed."""
    a0, a1, a2 = a
    w1 = fq2_mul(x, a0, c0)
    w2 = mul_c1(a1)
    w3 = fq2_add(x, a1, a2)
    w4 = mul_c1(w3)
    w5 = fq2_sub(x, w4, w2)
    w6 = fq2_mul_by_nonresidue(x, w5)
    w7 = fq2_add(x, w6, w1)
    w8 = fq2_add(x, a0, a1)
    w9 = c0_plus_c1()
    w10 = fq2_mul(x, w8, w9)
    w11 = fq2_sub(x, w10, w1)
    w12 = fq2_sub(x, w11, w2)
    w13 = fq2_add(x, a0, a2)
    w14 = fq2_mul(x, w13, c0)
    w15 = fq2_sub(x, w14, w1)
    return [w7, w12, fq2_add(x, w15, w2)]


def fq6_mul_by_01(x, a, c0, c1):
    return _mul_by_01(x, a, c0, lambda v: fq2_mul(x, v, c1), lambda: fq2_add(x, c0, c1))


def fq6_mul_by_01_constant1(x, a, c0, k1):   # k1 already in Montgomery form
    return _mul_by_01(x, a, c0, lambda v: fq2_mul_by_constant(x, v, k1), lambda: fq2_add_constant(x, c0, k1))


def _mul_by_034(x, key, a, c0, c3, sparse):
    def body(x, w):
        a0, a1 = _fq6_of(w[:6 * N]), _fq6_of(w[6 * N:12 * N])
        c0 = [list(w[12 * N:13 * N]), list(w[13 * N:14 * N])]
        c3 = [list(w[14 * N:15 * N]), list(w[15 * N:16 * N])]
        rest = w[16 * N:]
        w1 = sparse(x, a1, c3, rest)
        w2 = fq6_mul_by_nonresidue(x, w1)
        w3 = fq6_mul_by_fq2(x, a0, c0)
        new_c0 = fq6_add(x, w2, w3)
        w4 = fq6_add(x, a0, a1)
        w5 = fq2_add(x, c3, c0)
        w6 = sparse(x, w4, w5, rest)
        w7 = fq6_add(x, w1, w3)
        return _flat6(new_c0) + _flat6(fq6_sub(x, w6, w7))
    return body


def fq12_mul_by_034(x, a, c0, c3, c4):  # fq12.rs:267-285, #[component]
    body = _mul_by_034(x, None, a, c0, c3, lambda x, v, c, rest: fq6_mul_by_01(x, v, c, [list(rest[:N]), list(rest[N:])]))
    return x.component("fq12::mul_by_034_montgomery", list(a) + c0[0] + c0[1] + c3[0] + c3[1] + c4[0] + c4[1], body)


def fq12_mul_by_034_constant4(x, a, c0, c3, k4):   # fq12.rs:287-309, #[component(offcircuit_args = "c4")]
    body = _mul_by_034(x, None, a, c0, c3, lambda x, v, c, rest: fq6_mul_by_01_constant1(x, v, c, k4))
    return x.component(("fq12::mul_by_034_constant4_montgomery", k4), list(a) + c0[0] + c0[1] + c3[0] + c3[1], body)


def _g2_of(w):
    return [[list(w[(2 * i + j) * N:(2 * i + j + 1) * N]) for j in range(2)] for i in range(3)]


def g2_double_step(x, r):               # pairing.rs:359-407, #[component]: (R, line coefficients)
    def body(x, w):
        rx, ry, rz = _g2_of(w)
        a = fq2_half(x, fq2_mul(x, rx, ry))
        b = fq2_square(x, ry)
        c = fq2_square(x, rz)
        c3 = fq2_triple(x, c)
        e = fq2_mul_by_constant(x, c3, _mont2(G2_COEFF_B))
        f = fq2_triple(x, e)
        g = fq2_half(x, fq2_add(x, b, f))
        ryrz = fq2_add(x, ry, rz)
        ryrzs = fq2_square(x, ryrz)
        bc = fq2_add(x, b, c)
        h = fq2_sub(x, ryrzs, bc)
        i = fq2_sub(x, e, b)
        j = fq2_square(x, rx)
        es = fq2_square(x, e)
        j3 = fq2_triple(x, j)
        bf = fq2_sub(x, b, f)
        new_x = fq2_mul(x, a, bf)
        es3 = fq2_triple(x, es)
        gs = fq2_square(x, g)
        new_y = fq2_sub(x, gs, es3)
        new_z = fq2_mul(x, b, h)
        hn = fq2_neg(x, h)
        return _flat6([new_x, new_y, new_z]) + _flat6([hn, j3, i])
    return x.component("pairing::double_in_place_circuit_montgomery", list(r), body)


def g2_add_step(x, r, q):               # pairing.rs:409-462, #[component]
    def body(x, w):
        rx, ry, rz = _g2_of(w[:6 * N])
        qx, qy, _ = _g2_of(w[6 * N:])
        theta = fq2_sub(x, ry, fq2_mul(x, qy, rz))
        lam = fq2_sub(x, rx, fq2_mul(x, qx, rz))
        c = fq2_square(x, theta)
        d = fq2_square(x, lam)
        e = fq2_mul(x, lam, d)
        f = fq2_mul(x, rz, c)
        g = fq2_mul(x, rx, d)
        w3 = fq2_add(x, e, f)
        w4 = fq2_double(x, g)
        h = fq2_sub(x, w3, w4)
        neg_theta = fq2_neg(x, theta)
        w5 = fq2_mul(x, theta, qx)
        w6 = fq2_mul(x, lam, qy)
        j = fq2_sub(x, w5, w6)
        new_x = fq2_mul(x, lam, h)
        w7 = fq2_sub(x, g, h)
        w8 = fq2_mul(x, theta, w7)
        w9 = fq2_mul(x, e, ry)
        new_y = fq2_sub(x, w8, w9)
        new_z = fq2_mul(x, rz, e)
        return _flat6([new_x, new_y, new_z]) + _flat6([lam, neg_theta, j])
    return x.component("pairing::add_in_place_montgomery", list(r) + list(q), body)


def g2_mul_by_char(x, r):               # pairing.rs:475-498, #[component]
    def body(x, w):
        rx, ry, rz = _g2_of(w)
        sx = fq2_mul_by_constant(x, fq2_frobenius(x, rx, 1), _mont2(TWIST_MUL_BY_Q_X))
        sy = fq2_mul_by_constant(x, fq2_frobenius(x, ry, 1), _mont2(TWIST_MUL_BY_Q_Y))
        return _flat6([sx, sy, rz])
    return x.component("pairing::mul_by_char_montgomery", list(r), body)


def ell(x, f, coeffs, p):               # pairing.rs:160-171, plain function; p = (x, y, z) affine
    co = _fq6_of(coeffs)
    px, py = list(p[:N]), list(p[N:2 * N])
    c0 = fq2_mul_by_fq(x, co[0], py)
    c3 = fq2_mul_by_fq(x, co[1], px)
    return fq12_mul_by_034(x, f, c0, c3, co[2])


def ell_by_constant(x, f, k, p):        # pairing.rs:923-942, #[component(offcircuit_args = "coeffs")]; k standard form
    def body(x, w):
        f, px, py = list(w[:12 * N]), list(w[12 * N:13 * N]), list(w[13 * N:14 * N])
        c0 = fq2_mul_constant_by_fq(x, k[0], py)
        c1 = fq2_mul_constant_by_fq(x, k[1], px)
        return fq12_mul_by_034_constant4(x, f, c0, c1, _mont2(k[2]))
    return x.component(("pairing::ell_by_constant_montgomery", k), list(f) + list(p), body)


def g1_to_affine(x, p):                 # groth16.rs:26-47, #[component]
    def body(x, w):
        px, py, pz = (list(w[j * N:(j + 1) * N]) for j in range(3))
        zi = fq_inverse_montgomery(x, pz)
        zi2 = fq_mul(x, zi, zi)
        zi3 = fq_mul(x, zi, zi2)
        return fq_mul(x, px, zi2) + fq_mul(x, py, zi3) + bits_of(R_MONT, N)
    return x.component("groth16::projective_to_affine_montgomery", list(p), body)


def decompress_g1(x, x_m, y_flag):      # groth16.rs:113-143, #[component]
    def body(x, w):
        xm, flag = list(w[:N]), w[N]
        x2 = fq_mul(x, xm, xm)
        x3 = fq_mul(x, x2, xm)
        rhs = fq_add_constant(x, x3, 3 * R_MONT % P)      # + b, b = 3
        sy = fq_sqrt(x, rhs)
        sy_neg = fq_neg(x, sy)
        return xm + bn_select(x, sy, sy_neg, flag) + bits_of(R_MONT, N)
    return x.component("groth16::decompress_g1_from_compressed", list(x_m) + [y_flag], body)


# ======================================================================================== final exponentiation
BN_X = 4965661367192848881                    # ark_bn254::Config::X


def _naf(v):                                  # ark_ff::biginteger::arithmetic::find_naf: digits, least significant first
    out = []
    while v:
        if v & 1:
            z = 2 - (v % 4)
            v -= z
        else:
            z = 0
        out.append(z)
        v >>= 1
    return out


def fq12_conjugate(x, a):                     # fq12.rs:444-447
    a1 = _fq6_of(a[6 * N:])
    return list(a[:6 * N]) + _flat6(fq6_neg(x, a1))


def fq12_one():                               # Fq12::new_constant(ONE): Montgomery one in c0.c0.c0, zeros elsewhere
    return bits_of(R_MONT, N) + [FALSE] * (11 * N)


def cyclotomic_exp_fast_inverse(x, f):        # final_exponentiation.rs:65-93
    res = fq12_one()
    f_inv = fq12_inverse(x, f)
    found = False
    for d in reversed(_naf(BN_X)):
        if found:
            res = fq12_cyclotomic_square(x, res)
        if d:
            found = True
            res = fq12_mul(x, res, f if d > 0 else f_inv)
    return res


def exp_by_neg_x(x, f):                       # final_exponentiation.rs:95-98
    return fq12_conjugate(x, cyclotomic_exp_fast_inverse(x, f))


def final_exponentiation(x, f):               # final_exponentiation.rs:100-134, #[component]
    def body(x, f):
        f = list(f)
        f_inv = fq12_inverse(x, f)
        f_conj = fq12_conjugate(x, f)
        u = fq12_mul(x, f_inv, f_conj)
        u_frob = fq12_frobenius(x, u, 2)
        r = fq12_mul(x, u_frob, u)
        y0 = exp_by_neg_x(x, r)
        y1 = fq12_square(x, y0)
        y2 = fq12_square(x, y1)
        y3 = fq12_mul(x, y1, y2)
        y4 = exp_by_neg_x(x, y3)
        y5 = fq12_square(x, y4)
        y6 = exp_by_neg_x(x, y5)
        y7 = fq12_conjugate(x, y3)
        y8 = fq12_conjugate(x, y6)
        y9 = fq12_mul(x, y8, y4)
        y10 = fq12_mul(x, y9, y7)
        y11 = fq12_mul(x, y10, y1)
        y12 = fq12_mul(x, y10, y4)
        y13 = fq12_mul(x, y12, r)
        y14 = fq12_frobenius(x, y11, 1)
        y15 = fq12_mul(x, y14, y13)
        y16 = fq12_frobenius(x, y10, 2)
        y17 = fq12_mul(x, y16, y15)
        r2 = fq12_conjugate(x, r)
        y18 = fq12_mul(x, r2, y11)
        y19 = fq12_frobenius(x, y18, 3)
        return fq12_mul(x, y19, y17)
    return x.component("final_exponentiation_montgomery", list(f), body)


# ======================================================================================== Miller loop (Groth16 form)
ATE_LOOP_COUNT = [0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1, 1, 0, 0,
                  -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1]
assert sum(d << i for i, d in enumerate(ATE_LOOP_COUNT)) == 6 * BN_X + 2


def _f2add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def _f2sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def _f2neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def _f2scale(a, k):
    return (a[0] * k % P, a[1] * k % P)


def host_ell_coeffs(q):
    """pairing.rs:30-126 on plain integers: the line coefficients of a CONSTANT G2 point (affine, standard form)."""
    half = (P + 1) // 2
    r = [q[0], q[1], (1, 0)]

    def double():
        rx, ry, rz = r
        a = _f2scale(_f2mul(rx, ry), half)
        b, c = _f2mul(ry, ry), _f2mul(rz, rz)
        e = _f2mul(G2_COEFF_B, _f2add(_f2add(c, c), c))
        f = _f2add(_f2add(e, e), e)
        g = _f2scale(_f2add(b, f), half)
        s = _f2add(ry, rz)
        h = _f2sub(_f2mul(s, s), _f2add(b, c))
        i = _f2sub(e, b)
        j = _f2mul(rx, rx)
        es = _f2mul(e, e)
        r[:] = [_f2mul(a, _f2sub(b, f)), _f2sub(_f2mul(g, g), _f2add(_f2add(es, es), es)), _f2mul(b, h)]
        return (_f2neg(h), _f2add(_f2add(j, j), j), i)

    def add(p):
        rx, ry, rz = r
        theta = _f2sub(ry, _f2mul(p[1], rz))
        lam = _f2sub(rx, _f2mul(p[0], rz))
        c, d = _f2mul(theta, theta), _f2mul(lam, lam)
        e = _f2mul(lam, d)
        f = _f2mul(rz, c)
        g = _f2mul(rx, d)
        h = _f2sub(_f2add(e, f), _f2add(g, g))
        j = _f2sub(_f2mul(theta, p[0]), _f2mul(lam, p[1]))
        r[:] = [_f2mul(lam, h), _f2sub(_f2mul(theta, _f2sub(g, h)), _f2mul(e, ry)), _f2mul(rz, e)]
        return (lam, _f2neg(theta), j)

    def by_char(p):
        conj = lambda v: (v[0], (-v[1]) % P)
        return (_f2mul(conj(p[0]), TWIST_MUL_BY_Q_X), _f2mul(conj(p[1]), TWIST_MUL_BY_Q_Y))
    out = []
    neg_q = (q[0], _f2neg(q[1]))
    for bit in ATE_LOOP_COUNT[::-1][1:]:
        out.append(double())
        if bit == 1:
            out.append(add(q))
        elif bit == -1:
            out.append(add(neg_q))
    q1 = by_char(q)
    q2 = by_char(q1)
    q2 = (q2[0], _f2neg(q2[1]))
    out.append(add(q1))
    out.append(add(q2))
    return out


def ell_coeffs(x, q):                   # pairing.rs:507-545: the same walk on G2 wires
    q = list(q)
    qx, qy, qz = _g2_of(q)
    neg_q = _flat6([qx, fq2_neg(x, qy), qz])
    out, r = [], q
    for bit in ATE_LOOP_COUNT[::-1][1:]:
        res = g2_double_step(x, r)
        r = res[:6 * N]
        out.append(res[6 * N:])
        if bit:
            res = g2_add_step(x, r, q if bit == 1 else neg_q)
            r = res[:6 * N]
            out.append(res[6 * N:])
    q1 = g2_mul_by_char(x, q)
    q2 = g2_mul_by_char(x, q1)
    q2x, q2y, q2z = _g2_of(q2)
    q2 = _flat6([q2x, fq2_neg(x, q2y), q2z])
    res = g2_add_step(x, r, q1)
    r = res[:6 * N]
    out.append(res[6 * N:])
    res = g2_add_step(x, r, q2)
    out.append(res[6 * N:])
    return out


def miller_loop_groth16(x, p1, p2, p3, k1, k2, q3):   # pairing.rs:944-1009, #[component(offcircuit_args = "q1,q2")]
    def body(x, w):
        p1, p2, p3, q3 = list(w[:3 * N]), list(w[3 * N:6 * N]), list(w[6 * N:9 * N]), list(w[9 * N:])
        e1, e2 = iter(host_ell_coeffs(k1)), iter(host_ell_coeffs(k2))
        e3 = iter(ell_coeffs(x, q3))
        f = fq12_one()

        def lines(f):
            f = ell_by_constant(x, f, next(e1), p1)
            f = ell_by_constant(x, f, next(e2), p2)
            return ell(x, f, next(e3), p3)
        n = len(ATE_LOOP_COUNT)
        for i in range(n - 1, 0, -1):
            if i != n - 1:
                f = fq12_square(x, f)
            f = lines(f)
            if ATE_LOOP_COUNT[i - 1]:
                f = lines(f)
        f = lines(f)
        return lines(f)
    return x.component(("pairing::multi_miller_loop_groth16_evaluate_montgomery_fast", k1, k2),
                       list(p1) + list(p2) + list(p3) + list(q3), body)


# ---- the synthetic verification key of the product's named circuits (csrc/bn254_host.cpp synthetic_groth16(7, ..)):
# test data, not reference code -- scalars from a splitmix64 stream, points = scalar * generator
G2_GENERATOR = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
                 11559732032986387107991004021392285783925812861821192530917403151452391805634),
                (8495653923123431417604973247489272438418190587263600148770280649306958101930,
                 4082367875863433681332203403145435568316851327593401208105741076214120093531))


def _g2_affine_add(p, q):
    if p is None:
        return q
    if p == q:
        lam = _f2mul(_f2scale(_f2mul(p[0], p[0]), 3), _f2inv(_f2add(p[1], p[1])))
    else:
        lam = _f2mul(_f2sub(q[1], p[1]), _f2inv(_f2sub(q[0], p[0])))
    x3 = _f2sub(_f2sub(_f2mul(lam, lam), p[0]), q[0])
    return (x3, _f2sub(_f2mul(lam, _f2sub(p[0], x3)), p[1]))


def _g2_mul(p, k):
    acc = None
    while k:
        if k & 1:
            acc = _g2_affine_add(acc, p)
        p = _g2_affine_add(p, p)
        k >>= 1
    return acc


def synthetic_vk_scalars(seed=7):
    m = (1 << 64) - 1
    state = [(seed * 0x2545F4914F6CDD1D + 0x1234567) & m]

    def splitmix():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & m
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        return z ^ (z >> 31)

    def scalar():
        l = [splitmix() for _ in range(4)]
        l[3] &= 0x0FFFFFFFFFFFFFFF
        return (l[0] | l[1] << 64 | l[2] << 128 | l[3] << 192) or 1
    return dict(zip(("alpha", "beta", "gamma", "delta", "ic0", "ic1", "a", "b"), (scalar() for _ in range(8))))


# ======================================================================================== Fq2 square root, G2 decompression
def bn_equal(x, a, b):                  # bigint/cmp.rs:44-59, #[component]
    n = len(a)

    def body(x, w):
        xs = []
        for ai, bi in zip(w[:n], w[n:]):
            o = x.issue()
            x.gate(XOR, ai, bi, o)
            xs.append(o)
        return [equal_constant(x, xs, 0)]
    return x.component("bigint::equal", list(a) + list(b), body)[0]


def fq_is_qnr(x, a):                    # fq.rs:177-192: a^((p - 1) / 2) == -1
    y = fq_exp_by_constant(x, a, (P - 1) // 2)
    return bn_equal(x, y, bits_of((P - 1) * R_MONT % P, N))


def fq2_sqrt_general(x, a):             # fq2.rs:425-447, #[component]
    def body(x, w):
        a0, a1 = list(w[:N]), list(w[N:])
        alpha = fq_add(x, fq_mul(x, a0, a0), fq_mul(x, a1, a1))      # norm_montgomery
        alpha_sqrt = fq_sqrt(x, alpha)
        delta = fq_half(x, fq_add(x, alpha_sqrt, a0))
        is_qnr = fq_is_qnr(x, delta)
        delta_alt = fq_sub(x, delta, alpha_sqrt)
        delta_final = bn_select(x, delta_alt, delta, is_qnr)
        c0 = fq_sqrt(x, delta_final)
        c0_inv = fq_inverse_montgomery(x, c0)
        c1_half = fq_half(x, a1)
        return c0 + fq_mul(x, c0_inv, c1_half)
    r = x.component("fq2::sqrt_general_montgomery", a[0] + a[1], body)
    return [r[:N], r[N:]]


def decompress_g2(x, xw, y_flag):       # groth16.rs:145-182, #[component]
    def body(x, w):
        px, flag = [list(w[:N]), list(w[N:2 * N])], w[2 * N]
        x2 = fq2_square(x, px)
        x3 = fq2_mul(x, x2, px)
        y2 = fq2_add_constant(x, x3, _mont2(G2_COEFF_B))
        y = fq2_sqrt_general(x, y2)
        neg_y = fq2_neg(x, y)
        y0 = bn_select(x, y[0], neg_y[0], flag)
        y1 = bn_select(x, y[1], neg_y[1], flag)
        return px[0] + px[1] + y0 + y1 + bits_of(R_MONT, N) + [FALSE] * N
    return x.component("groth16::decompress_g2_from_compressed", list(xw) + [y_flag], body)


# ======================================================================================== multiplexers, G1
def basic_multiplexer(x, a, s, w):      # basic.rs:73-105, #[component(offcircuit_args = "w")]
    n = len(a)

    def body(x, inp):
        cur, sel = list(inp[:n]), inp[n:]
        for sl in sel:                  # pairs reduced from the LSB selector up: selector(high, low, sel)
            cur = [selector(x, cur[i + 1], cur[i], sl) for i in range(0, len(cur), 2)]
        return [cur[0]]
    assert n == 1 << w and len(s) == w
    return x.component(("basic::multiplexer", w), list(a) + list(s), body)[0]


def bn_multiplexer(x, a, s, w):         # bigint/cmp.rs:171-193, #[bn_component]
    nb, cnt = len(a[0]), len(a)

    def body(x, inp):
        arr, sel = [inp[j * nb:(j + 1) * nb] for j in range(cnt)], inp[cnt * nb:]
        return [basic_multiplexer(x, [ai[i] for ai in arr], sel, w) for i in range(nb)]
    return x.component(("bigint::multiplexer", w), [b for ai in a for b in ai] + list(s), body)


def g1_add(x, p, q):                    # g1.rs:159-235, #[component]
    def body(x, w):
        x1, y1, z1, x2, y2, z2 = (list(w[j * N:(j + 1) * N]) for j in range(6))
        z1s = fq_mul(x, z1, z1)
        z2s = fq_mul(x, z2, z2)
        z1c = fq_mul(x, z1s, z1)
        z2c = fq_mul(x, z2s, z2)
        u1 = fq_mul(x, x1, z2s)
        u2 = fq_mul(x, x2, z1s)
        s1 = fq_mul(x, y1, z2c)
        s2 = fq_mul(x, y2, z1c)
        r = fq_sub(x, s1, s2)
        h = fq_sub(x, u1, u2)
        h2 = fq_mul(x, h, h)
        g = fq_mul(x, h, h2)
        v = fq_mul(x, u1, h2)
        r2 = fq_mul(x, r, r)
        r2g = fq_add(x, r2, g)
        vd = fq_double(x, v)
        x3 = fq_sub(x, r2g, vd)
        vx3 = fq_sub(x, v, x3)
        ww = fq_mul(x, r, vx3)
        s1g = fq_mul(x, s1, g)
        y3 = fq_sub(x, ww, s1g)
        z1z2 = fq_mul(x, z1, z2)
        z3 = fq_mul(x, z1z2, h)
        z1_0 = equal_constant(x, z1, 0)
        z2_0 = equal_constant(x, z2, 0)
        zero = [FALSE] * N
        sel = [z1_0, z2_0]
        return (bn_multiplexer(x, [x3, x2, x1, zero], sel, 2) + bn_multiplexer(x, [y3, y2, y1, zero], sel, 2) +
                bn_multiplexer(x, [z3, z2, z1, zero], sel, 2))
    return x.component("g1::add_montgomery", list(p) + list(q), body)


# ======================================================================================== MSM with constant bases
# The multiplexer tables are constants: multiples of the base in arkworks' Jacobian coordinates (ark-ec 0.5
# short_weierstrass::Projective, absent from /root/reference: `+=` is add-2007-bl, doubling dbl-2009-l with a = 0, the
# identity is (1, 1, 0)).  Their VALUES are restated from the published formulas -- the one shared assumption of this
# section; the gadget structure around them is from g1.rs:275-400.
def _jac_double(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A, B = X * X % P, Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    return (X3, (E * (D - X3) - 8 * C) % P, 2 * Y * Z % P)


def _jac_add(p, q):
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    Z1Z1, Z2Z2 = Z1 * Z1 % P, Z2 * Z2 % P
    U1, U2 = X1 * Z2Z2 % P, X2 * Z1Z1 % P
    S1, S2 = Y1 * Z2 * Z2Z2 % P, Y2 * Z1 * Z1Z1 % P
    if U1 == U2 and S1 == S2:
        return _jac_double(p)
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = (-H * I) % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r + J - 2 * V) % P
    return (X3, (r * (V - X3) + 2 * S1 * J) % P, 2 * Z1 * Z2 * H % P)


def _g1_const(p):                       # G1Projective::new_constant(as_montgomery(p))
    return bits_of(p[0] * R_MONT % P, N) + bits_of(p[1] * R_MONT % P, N) + bits_of(p[2] * R_MONT % P, N)


def g1_multiplexer(x, pts, sel, w):     # g1.rs:275-306, #[component(offcircuit_args = "w")]
    cnt = len(pts)

    def body(x, inp):
        ps, s = [list(inp[j * 3 * N:(j + 1) * 3 * N]) for j in range(cnt)], list(inp[cnt * 3 * N:])
        return sum((bn_multiplexer(x, [q[c * N:(c + 1) * N] for q in ps], s, w) for c in range(3)), [])
    return x.component(("g1::multiplexer", w), [b for q in pts for b in q] + list(sel), body)


def g1_scalar_mul_const(x, scalar, base, W=10):      # g1.rs:308-368, base = Jacobian (x, y, z), standard form
    def body(x, s):
        n = 1 << W
        bases, p = [], (1, 1, 0)
        for _ in range(n):
            bases.append(p)
            p = _jac_add(p, base)
        parts, idx = [], 0
        while idx < N:
            w = min(W, N - idx)
            parts.append(g1_multiplexer(x, [_g1_const(b) for b in bases[:1 << w]], s[idx:idx + w], w))
            idx += W
            nb = []
            for b in bases:
                for _ in range(w):
                    b = _jac_add(b, b)
                nb.append(b)
            bases = nb
        acc = parts[0]
        for q in parts[1:]:
            acc = g1_add(x, acc, q)
        return acc
    return x.component(("g1::scalar_mul_by_constant_base_montgomery", W, base), list(scalar), body)


def g1_msm_const(x, scalars, bases, W=10):           # g1.rs:370-400
    def body(x, w):
        parts = [g1_scalar_mul_const(x, w[i * N:(i + 1) * N], b, W) for i, b in enumerate(bases)]
        acc = parts[0]
        for q in parts[1:]:
            acc = g1_add(x, acc, q)
        return acc
    assert scalars
    return x.component(("g1::msm_with_constant_bases_montgomery", W, tuple(bases)), [b for s_ in scalars for b in s_], body)


# ======================================================================================== host pairing (for one constant)
def _f6mul(a, b):                       # Fq6 = Fq2[v] / (v^3 - xi)
    a0, a1, a2 = a
    b0, b1, b2 = b
    t0, t1, t2 = _f2mul(a0, b0), _f2mul(a1, b1), _f2mul(a2, b2)
    c0 = _f2add(t0, _f2mul(XI, _f2sub(_f2mul(_f2add(a1, a2), _f2add(b1, b2)), _f2add(t1, t2))))
    c1 = _f2add(_f2sub(_f2mul(_f2add(a0, a1), _f2add(b0, b1)), _f2add(t0, t1)), _f2mul(XI, t2))
    c2 = _f2add(_f2sub(_f2mul(_f2add(a0, a2), _f2add(b0, b2)), _f2add(t0, t2)), t1)
    return (c0, c1, c2)


def _f6add(a, b):
    return tuple(_f2add(u, v) for u, v in zip(a, b))


def _f6sub(a, b):
    return tuple(_f2sub(u, v) for u, v in zip(a, b))


def _f6neg(a):
    return tuple(_f2neg(u) for u in a)


def _f6mulv(a):
    return (_f2mul(XI, a[2]), a[0], a[1])


def _f6inv(a):
    a0, a1, a2 = a
    c0 = _f2sub(_f2mul(a0, a0), _f2mul(XI, _f2mul(a1, a2)))
    c1 = _f2sub(_f2mul(XI, _f2mul(a2, a2)), _f2mul(a0, a1))
    c2 = _f2sub(_f2mul(a1, a1), _f2mul(a0, a2))
    t = _f2inv(_f2add(_f2mul(a0, c0), _f2mul(XI, _f2add(_f2mul(a2, c1), _f2mul(a1, c2)))))
    return (_f2mul(c0, t), _f2mul(c1, t), _f2mul(c2, t))


def _f12mul(a, b):                      # Fq12 = Fq6[w] / (w^2 - v)
    t0, t1 = _f6mul(a[0], b[0]), _f6mul(a[1], b[1])
    return (_f6add(t0, _f6mulv(t1)), _f6sub(_f6mul(_f6add(a[0], a[1]), _f6add(b[0], b[1])), _f6add(t0, t1)))


def _f12inv(a):
    t = _f6inv(_f6sub(_f6mul(a[0], a[0]), _f6mulv(_f6mul(a[1], a[1]))))
    return (_f6mul(a[0], t), _f6neg(_f6mul(a[1], t)))


def _f12conj(a):
    return (a[0], _f6neg(a[1]))


def _f12frob(a, i):
    conj = (lambda v: (v[0], (-v[1]) % P)) if i % 2 else (lambda v: v)

    def f6(c):
        return (conj(c[0]), _f2mul(conj(c[1]), FROB_FP6_C1[i % 6]), _f2mul(conj(c[2]), FROB_FP6_C2[i % 6]))
    c1 = f6(a[1])
    return (f6(a[0]), tuple(_f2mul(v, FROB_FP12_C1[i % 12]) for v in c1))


F12_ONE = (((1, 0), (0, 0), (0, 0)), ((0, 0), (0, 0), (0, 0)))


def _f12pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = _f12mul(r, a)
        a = _f12mul(a, a)
        e >>= 1
    return r


def host_miller_loop(p, q):
    """ark-ec's BN multi_miller_loop for one pair, through the line coefficients above (p, q affine, standard form)."""
    co = iter(host_ell_coeffs(q))

    def ell(f):
        c0, c1, c2 = next(co)
        line = ((_f2scale(c0, p[1]), (0, 0), (0, 0)), (_f2scale(c1, p[0]), c2, (0, 0)))
        return _f12mul(f, line)
    f = F12_ONE
    n = len(ATE_LOOP_COUNT)
    for i in range(n - 1, 0, -1):
        if i != n - 1:
            f = _f12mul(f, f)
        f = ell(f)
        if ATE_LOOP_COUNT[i - 1]:
            f = ell(f)
    return ell(ell(f))


def host_final_exponentiation(f):
    """final_exponentiation.rs:38-63 (= ark-ec's BN final exponentiation) on plain integers."""
    neg_x = lambda v: _f12conj(_f12pow(v, BN_X))
    u = _f12mul(_f12inv(f), _f12conj(f))
    r = _f12mul(_f12frob(u, 2), u)
    y0 = neg_x(r)
    y1 = _f12mul(y0, y0)
    y2 = _f12mul(y1, y1)
    y3 = _f12mul(y1, y2)
    y4 = neg_x(y3)
    y5 = _f12mul(y4, y4)
    y6 = neg_x(y5)
    y7, y8 = _f12conj(y3), _f12conj(y6)
    y9 = _f12mul(y8, y4)
    y10 = _f12mul(y9, y7)
    y11 = _f12mul(y10, y1)
    y12 = _f12mul(y10, y4)
    y13 = _f12mul(y12, r)
    y14 = _f12frob(y11, 1)
    y15 = _f12mul(y14, y13)
    y16 = _f12frob(y10, 2)
    y17 = _f12mul(y16, y15)
    y18 = _f12mul(_f12conj(r), y11)
    return _f12mul(_f12frob(y18, 3), y17)


# ======================================================================================== Groth16 verifier
def fq2_equal_constant(x, a, k):        # fq2.rs:148-158 (k in the wires' form)
    u, v = equal_constant(x, a[0], k[0]), equal_constant(x, a[1], k[1])
    w = x.issue()
    x.gate(AND, u, v, w)
    return w


def fq6_equal_constant(x, a, k):        # fq6.rs:139-152
    u, v, w = (fq2_equal_constant(x, a[j], k[j]) for j in range(3))
    t, y = x.issue(), x.issue()
    x.gate(AND, u, v, t)
    x.gate(AND, t, w, y)
    return y


def fq12_equal_constant(x, a, k):       # fq12.rs:158-168
    u = fq6_equal_constant(x, _fq6_of(a[:6 * N]), k[0])
    v = fq6_equal_constant(x, _fq6_of(a[6 * N:]), k[1])
    w = x.issue()
    x.gate(AND, u, v, w)
    return w


def _g1_mul_affine(k):                  # k * (1, 2) as an affine point
    acc, p = (1, 1, 0), (1, 2, 1)
    while k:
        if k & 1:
            acc = _jac_add(acc, p)
        p = _jac_double(p)
        k >>= 1
    zi = pow(acc[2], -1, P)
    return (acc[0] * zi * zi % P, acc[1] * zi * zi * zi % P)


def synthetic_vk(seed=7):
    sc = synthetic_vk_scalars(seed)
    return dict(alpha_g1=_g1_mul_affine(sc["alpha"]), beta_g2=_g2_mul(G2_GENERATOR, sc["beta"]),
                gamma_g2=_g2_mul(G2_GENERATOR, sc["gamma"]), delta_g2=_g2_mul(G2_GENERATOR, sc["delta"]),
                gamma_abc_g1=[_g1_mul_affine(sc["ic0"]), _g1_mul_affine(sc["ic1"])])


def groth16_verify(x, publics, a, b, c, vk):          # groth16.rs:50-105 (plain function)
    neg2 = lambda q: (q[0], _f2neg(q[1]))
    bases = [(pt[0], pt[1], 1) for pt in vk["gamma_abc_g1"][1:1 + len(publics)]]
    msm_temp = g1_msm_const(x, publics, bases)
    g0 = vk["gamma_abc_g1"][0]
    msm = g1_add(x, msm_temp, _g1_const((g0[0], g0[1], 1)))
    msm_affine = g1_to_affine(x, msm)
    f = miller_loop_groth16(x, msm_affine, c, a, neg2(vk["gamma_g2"]), neg2(vk["delta_g2"]), b)
    alpha_beta = _f12inv(host_final_exponentiation(host_miller_loop(vk["alpha_g1"], neg2(vk["beta_g2"]))))
    f = final_exponentiation(x, f)
    ab_m = tuple(tuple(_mont2(c2) for c2 in c6) for c6 in alpha_beta)
    return fq12_equal_constant(x, f, ab_m)


def groth16_verify_compressed(x, w, n_public, vk):     # groth16.rs:250-268; input layout of the product's root
    o = n_public * N
    publics = [list(w[i * N:(i + 1) * N]) for i in range(n_public)]
    a = decompress_g1(x, w[o:o + N], w[o + N])
    o += N + 1
    b = decompress_g2(x, w[o:o + 2 * N], w[o + 2 * N])
    o += 2 * N + 1
    c = decompress_g1(x, w[o:o + N], w[o + N])
    return [groth16_verify(x, publics, a, b, c, vk)]


def groth16_verify_uncompressed(x, w, n_public, vk):   # garbled_groth16.rs:108-176: affine points, z = 1 constants
    o = n_public * N
    publics = [list(w[i * N:(i + 1) * N]) for i in range(n_public)]
    one, zero = bits_of(R_MONT, N), [FALSE] * N
    a = list(w[o:o + 2 * N]) + one
    o += 2 * N
    b = list(w[o:o + 4 * N]) + one + zero
    o += 4 * N
    c = list(w[o:o + 2 * N]) + one
    return [groth16_verify(x, publics, a, b, c, vk)]


# ======================================================================================== roots (the product's named circuits)
def build(circuit):
    """(type, a, b, c, outputs, n_inputs) of a named circuit; names as in gsv_program_build."""
    if circuit == "fq_add":
        x = Ctx(2 * N)
        w = list(range(2, 2 + 2 * N))
        return x.finish(fq_add(x, w[:N], w[N:]))
    if circuit == "fq_mul":
        x = Ctx(2 * N)
        w = list(range(2, 2 + 2 * N))
        return x.finish(fq_mul(x, w[:N], w[N:]))
    if circuit.startswith("bn_mul"):
        n = int(circuit[6:])
        x = Ctx(2 * n)
        w = list(range(2, 2 + 2 * n))
        return x.finish(bn_mul(x, w[:n], w[n:]))
    if circuit == "fq2_mul":
        x = Ctx(4 * N)
        w = list(range(2, 2 + 4 * N))
        r = fq2_mul(x, [w[:N], w[N:2 * N]], [w[2 * N:3 * N], w[3 * N:]])
        return x.finish(r[0] + r[1])
    if circuit == "fq6_mul":
        x = Ctx(12 * N)
        w = list(range(2, 2 + 12 * N))
        return x.finish(_flat6(fq6_mul(x, _fq6_of(w[:6 * N]), _fq6_of(w[6 * N:]))))
    if circuit == "fq12_mul":
        x = Ctx(24 * N)
        w = list(range(2, 2 + 24 * N))
        return x.finish(fq12_mul(x, w[:12 * N], w[12 * N:]))
    if circuit in ("fq12_square", "fq12_cyclotomic_square", "fq12_inverse") or circuit.startswith("fq12_frobenius"):
        x = Ctx(12 * N)
        w = list(range(2, 2 + 12 * N))
        if circuit.startswith("fq12_frobenius"):
            return x.finish(fq12_frobenius(x, w, int(circuit[14:])))
        return x.finish({"fq12_square": fq12_square, "fq12_cyclotomic_square": fq12_cyclotomic_square,
                         "fq12_inverse": fq12_inverse}[circuit](x, w))
    if circuit in ("g2_double_step", "g2_mul_by_char", "g1_to_affine"):
        n = 6 * N if circuit != "g1_to_affine" else 3 * N
        x = Ctx(n)
        w = list(range(2, 2 + n))
        return x.finish({"g2_double_step": g2_double_step, "g2_mul_by_char": g2_mul_by_char, "g1_to_affine": g1_to_affine}[circuit](x, w))
    if circuit == "g2_add_step":
        x = Ctx(12 * N)
        w = list(range(2, 2 + 12 * N))
        return x.finish(g2_add_step(x, w[:6 * N], w[6 * N:]))
    if circuit == "ell":
        x = Ctx(21 * N)
        w = list(range(2, 2 + 21 * N))
        return x.finish(ell(x, w[:12 * N], w[12 * N:18 * N], w[18 * N:]))
    if circuit == "ell_const":
        x = Ctx(15 * N)
        w = list(range(2, 2 + 15 * N))
        return x.finish(ell_by_constant(x, w[:12 * N], ((3, 5), (7, 11), (13, 17)), w[12 * N:]))
    if circuit == "g1_add":
        x = Ctx(6 * N)
        w = list(range(2, 2 + 6 * N))
        return x.finish(g1_add(x, w[:3 * N], w[3 * N:]))
    if circuit == "decompress_g1":
        x = Ctx(N + 1)
        w = list(range(2, 2 + N + 1))
        return x.finish(decompress_g1(x, w[:N], w[N]))
    if circuit == "fq_sqrt":
        x = Ctx(N)
        return x.finish(fq_sqrt(x, list(range(2, 2 + N))))
    if circuit == "fq_inverse":
        x = Ctx(N)
        return x.finish(fq_inverse_montgomery(x, list(range(2, 2 + N))))
    raise ValueError(circuit)

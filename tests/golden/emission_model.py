"""An INDEPENDENT second statement of the reference's gate emission order (TEST INFRASTRUCTURE).

Written from the reference's Rust gadgets (src/gadgets/basic.rs, bigint/{add,cmp,mul}.rs,
bn254/{fp254impl,fq2,fq6,fq12}.rs), NOT from the product's C++ generator
(csrc/gadgets*.cpp, csrc/circuit.cpp), and with a different mechanism on purpose:

  * wires are global SSA ids in `issue_wire` order, gates go to one flat (type, a, b, c) stream in `add_gate`
    order -- no templates, no credit stacks;
  * liveness is the reference's rule stated globally: a wire is live iff it is
    read by a gate, passed as an input to a `#[component]` call, or is an output of the root; a gate whose
    output wire is not live keeps its gate index but gets c = UNREACHABLE.  (The product records per-component
    credit templates per output-liveness mask instead.)
  * a component body is recorded once per key and replayed by renumbering (numpy), which is what keeps
    Fq12::mul (20 M gates) fast enough for the CPU suite.

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
    raise ValueError(circuit)

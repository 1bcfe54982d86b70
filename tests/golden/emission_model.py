"""An INDEPENDENT second statement of the reference's gate emission order (TEST INFRASTRUCTURE).

Written from the reference's Rust gadgets (src/gadgets/basic.rs, bigint/{add,cmp,mul}.rs,
bn254/{fp254impl,fq2,fq6,fq12,g1,pairing}.rs, groth16.rs:26-47) -- multiplications, the
binary Fp inverse and the tower inverses, squares, Frobenius maps, exponentiation by a constant, G1 addition,
the G2 / line-evaluation steps of the pairing -- NOT from the product's C++ generator
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

"""Plain-integer BN254 tower arithmetic used to check the gadget restatement functionally
(the role arkworks plays in the reference's gadget tests, e.g. src/gadgets/bn254/fq12.rs:450+)."""
import random

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 1 << 254
N_BITS = 254


def to_mont(x):
    return x * R % P


def bits_le(x, n=N_BITS):
    return [(x >> i) & 1 for i in range(n)]


def from_bits(bits):
    return sum(int(b) << i for i, b in enumerate(bits))


def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fq2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


XI = (9, 1)


def fq6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = fq2_add(fq2_mul(a0, b0), fq2_mul(XI, fq2_add(fq2_mul(a1, b2), fq2_mul(a2, b1))))
    c1 = fq2_add(fq2_add(fq2_mul(a0, b1), fq2_mul(a1, b0)), fq2_mul(XI, fq2_mul(a2, b2)))
    c2 = fq2_add(fq2_add(fq2_mul(a0, b2), fq2_mul(a1, b1)), fq2_mul(a2, b0))
    return (c0, c1, c2)


def fq6_add(a, b):
    return tuple(fq2_add(x, y) for x, y in zip(a, b))


def fq6_mul_by_v(a):
    return (fq2_mul(XI, a[2]), a[0], a[1])


def fq12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    c0 = fq6_add(fq6_mul(a0, b0), fq6_mul_by_v(fq6_mul(a1, b1)))
    c1 = fq6_add(fq6_mul(a0, b1), fq6_mul(a1, b0))
    return (c0, c1)


def rand_fq(rng):
    return rng.randrange(P)


def rand_fq12(rng):
    return tuple(tuple((rand_fq(rng), rand_fq(rng)) for _ in range(3)) for _ in range(2))


def fq12_flatten(a):
    """Fq12 -> 12 Fq coefficients in wire order: c0.(c0.(c0,c1), c1.., c2..), c1..."""
    return [c for f6 in a for f2 in f6 for c in f2]


def fq12_bits_mont(a):
    out = []
    for c in fq12_flatten(a):
        out += bits_le(to_mont(c))
    return out

"""Host-side logic: the C++ gadget restatement, the credits/liveness pass, the planner, and the
oracle's whole-stream garble/evaluate -- everything that runs without a GPU."""
import random

import numpy as np
import pytest

import bn254_ref as bn

WIRE_DEAD = 0xFFFFFFFF


# gate counts derived at survey time with an independent re-statement
MODEL = {
    "fq_add": (3298, 1523, 1522),
    "bn_mul254": (183326, 55602, 55602),
    "fq_mul": (414284, 102093, 102093),
    "fq12_mul": (20284982, 5439790, 5439206),
}


@pytest.mark.parametrize("name", list(MODEL))
def test_gate_counts_match_model(circuit, name):
    p, st = circuit(name)
    gates, nonfree, cts = MODEL[name]
    assert p.n_gates == gates
    assert sum(p.type_count[:8]) == nonfree
    assert p.n_ciphertexts == cts
    assert st.n_gates == gates
    # #ciphertexts == #live non-free gates (garble_test.rs:82,129-133)
    assert int(((st.type < 8) & (st.c != WIRE_DEAD)).sum()) == cts


def test_fq12_mul_type_mix(circuit):
    p, _ = circuit("fq12_mul")
    # Xor 14.69 M, And 4.83 M, Nand 297 k, Xnor 151 k, Or 156 k, Cimp 148 k
    assert p.type_count == [4831858, 296672, 4572, 0, 2420, 148336, 0, 155932, 14693832, 151360, 0]
    assert p.n_inputs == 6096 and p.n_outputs == 3048


def test_stream_is_well_formed(circuit):
    for name in ("gate_zoo", "fq_add", "fq_mul"):
        p, st = circuit(name)
        defined = np.zeros(st.n_wires, bool)
        defined[: 2 + st.n_inputs] = True
        live = st.c != WIRE_DEAD
        # SSA: every live gate defines a fresh, increasing id
        ids = st.c[live]
        assert np.all(np.diff(ids.astype(np.int64)) == 1) and ids[0] == 2 + st.n_inputs
        # reads only of earlier definitions
        pos = np.full(st.n_wires, -1, np.int64)
        pos[ids] = np.nonzero(live)[0]
        g = np.arange(st.n_gates)
        assert np.all(pos[st.a] < g) and np.all(pos[st.b] < g)
        assert np.all(st.outputs < st.n_wires)


def test_gate_zoo_dead_gate(circuit):
    p, st = circuit("gate_zoo")
    assert p.n_gates == 14 and p.n_live_gates == 13
    assert int((st.c == WIRE_DEAD).sum()) == 1
    assert p.n_ciphertexts == 8 + 2  # 8 AND-family + Or/Nimp on constants; the dead And emits none
    for a in (0, 1):
        for b in (0, 1):
            out = st.execute([a, b])
            want = [a & b, 1 - (a & b), a & (1 - b), (1 - a) | b, (1 - a) & b, (1 - b) | a, 1 - (a | b), a | b,
                    a ^ b, 1 - (a ^ b), 1 - a, 1]
            assert list(out) == want


def _fq_bits(x):
    return bn.bits_le(bn.to_mont(x))


def test_fq_add_functional(circuit):
    _, st = circuit("fq_add")
    rng = random.Random(11)
    for _ in range(10):
        a, b = bn.rand_fq(rng), bn.rand_fq(rng)
        out = st.execute(bn.bits_le(a) + bn.bits_le(b))
        assert bn.from_bits(out) == (a + b) % bn.P


def test_fq_mul_functional(circuit):
    _, st = circuit("fq_mul")
    rng = random.Random(12)
    for a, b in [(0, 0), (1, 1), (bn.P - 1, bn.P - 1)] + [(bn.rand_fq(rng), bn.rand_fq(rng)) for _ in range(5)]:
        out = st.execute(_fq_bits(a) + _fq_bits(b))
        assert bn.from_bits(out) == bn.to_mont(a * b % bn.P)


def test_fq_expr_functional(circuit):
    # tests/streaming_evaluate.rs:392-447: ((a^2) b) + a with a = 13, b = 7
    _, st = circuit("fq_expr")
    for a, b in [(13, 7), (bn.P - 2, 12345)]:
        out = st.execute(_fq_bits(a) + _fq_bits(b))
        assert bn.from_bits(out) == bn.to_mont((a * a * b + a) % bn.P)


def test_fq12_mul_functional(circuit):
    _, st = circuit("fq12_mul")
    rng = random.Random(13)
    for _ in range(2):
        a, b = bn.rand_fq12(rng), bn.rand_fq12(rng)
        out = st.execute(bn.fq12_bits_mont(a) + bn.fq12_bits_mont(b))
        assert list(out) == bn.fq12_bits_mont(bn.fq12_mul(a, b))


def _labels_for(res, bits):
    delta = np.frombuffer(res["delta"], np.uint8)
    act = res["input_label0"].copy()
    act[np.asarray(bits, bool)] ^= delta
    return act


@pytest.mark.parametrize("hasher", [0, 1])
@pytest.mark.parametrize("name", ["gate_zoo", "fq_add", "fq_mul"])
def test_oracle_garble_evaluate_consistency(circuit, orc, name, hasher):
    """tests/streaming_evaluate.rs:125-132 / fq12_mul_e2e.rs:217-235: gw.select(value) == active."""
    p, st = circuit(name)
    rng = random.Random(21)
    for seed in (0, 42):
        g = st.garble(hasher, seed)
        assert g["n_ct"] == p.n_ciphertexts == g["cts"].shape[0]
        assert orc.chain([bytes(c) for c in g["cts"][:64]]) is not None
        bits = [rng.randrange(2) for _ in range(st.n_inputs)]
        delta = np.frombuffer(g["delta"], np.uint8)
        true1 = bytes(np.frombuffer(g["true_label0"], np.uint8) ^ delta)
        ev = st.evaluate(hasher, true1, g["false_label0"], _labels_for(g, bits), bits, g["cts"])
        assert ev["rc"] == 0 and ev["n_ct_used"] == g["n_ct"]
        assert ev["ct_commit"] == g["ct_commit"]  # evaluator's hash == committed hash
        assert list(ev["output_bits"]) == list(st.execute(bits))
        want = g["output_label0"].copy()
        want[ev["output_bits"].astype(bool)] ^= delta
        assert np.array_equal(ev["output_active"], want)


def test_oracle_ciphertext_exhaustion(circuit):
    _, st = circuit("fq_add")
    g = st.garble(0, 1)
    bits = [0] * st.n_inputs
    delta = np.frombuffer(g["delta"], np.uint8)
    true1 = bytes(np.frombuffer(g["true_label0"], np.uint8) ^ delta)
    ev = st.evaluate(0, true1, g["false_label0"], g["input_label0"], bits, g["cts"][:-1])
    assert ev["rc"] == -2  # "Ciphertext source exhausted", evaluate_mode.rs:140-142


def test_seed_draw_order(circuit, orc):
    """garble_mode.rs:80-97,116-118: delta, false.label0, true.label0, then the input label0s."""
    _, st = circuit("fq_add")
    g = st.garble(0, 777)
    r = orc.Rng(777)
    assert r.label() == g["delta"] and r.label() == g["false_label0"] and r.label() == g["true_label0"]
    for i in range(st.n_inputs):
        assert r.label() == bytes(g["input_label0"][i])


def test_planner_invariants(gsv):
    p = gsv.Program("fq12_mul")
    assert p.n_calls == 867 and p.n_tasks == 13
    assert p.max_task_slots <= 1536
    small = gsv.Program("fq12_mul", max_task_slots=4096)
    assert small.n_calls == 507 and small.n_gates == p.n_gates and small.n_ciphertexts == p.n_ciphertexts
    one = gsv.Program("fq_mul", max_task_slots=8192, max_task_gates=10**7)
    assert one.n_calls == 1  # whole circuit as a single task


# ---- pairing / Groth16 gadgets (src/gadgets/bn254/*, src/gadgets/groth16.rs): functional checks through
# the host ExecuteMode walker, the role arkworks plays in the reference's gadget tests
def _mont_bits(x):
    return bn.bits_le(bn.to_mont(x))


def _from_mont(bits):
    return bn.from_bits(bits) * pow(bn.R, -1, bn.P) % bn.P


def test_fq_inverse_and_sqrt_functional(gsv):
    inv = gsv.Program("fq_inverse", lane_only=True)
    assert inv.n_gates == 23200543
    for x in (1, 123456789, bn.P - 2):
        assert _from_mont(inv.execute(_mont_bits(x))) == pow(x, -1, bn.P)
    sq = gsv.Program("fq_sqrt", lane_only=True)
    y = 987654321
    r = _from_mont(sq.execute(_mont_bits(y * y % bn.P)))
    assert r * r % bn.P == y * y % bn.P


def test_fq12_square_functional(gsv):
    p = gsv.Program("fq12_square", lane_only=True)
    rng = random.Random(31)
    a = bn.rand_fq12(rng)
    assert list(p.execute(bn.fq12_bits_mont(a))) == bn.fq12_bits_mont(bn.fq12_mul(a, a))


def test_g1_add_functional(gsv):
    """G1Projective::add_montgomery (g1.rs:159-235) against affine chord addition."""
    p = gsv.Program("g1_add", lane_only=True)

    def aff_add(p1, p2):
        lam = (p2[1] - p1[1]) * pow(p2[0] - p1[0], -1, bn.P) % bn.P
        x = (lam * lam - p1[0] - p2[0]) % bn.P
        return x, (lam * (p1[0] - x) - p1[1]) % bn.P

    g1, g2 = (1, 2), None
    lam = 3 * pow(4, -1, bn.P) % bn.P  # 2G
    x2 = (lam * lam - 2) % bn.P
    g2 = (x2, (lam * (1 - x2) - 2) % bn.P)
    g3 = aff_add(g1, g2)
    bits = []
    for pt in (g2, g3):
        bits += _mont_bits(pt[0]) + _mont_bits(pt[1]) + _mont_bits(1)
    out = p.execute(bits)
    x, y, z = (_from_mont(out[i * 254:(i + 1) * 254]) for i in range(3))
    zi = pow(z, -1, bn.P)
    assert (x * zi * zi % bn.P, y * zi * zi * zi % bn.P) == aff_add(g2, g3)


def test_groth16_verifier_accepts_and_rejects(gsv, lane_program):
    """groth16_verify_compressed (groth16.rs:250-268) over a synthetic key: 11.46 G gates walked in
    ExecuteMode; the valid proof yields 1, a different public input yields 0 (the reference's
    true / bit-flip cases, groth16.rs:510-604).  The README's 11 174 708 821 is 2.5 % lower; the key moves the
    count by a few 1e-5 only, so that difference is between the current sources and the
    published figure, not the key."""
    p = lane_program("groth16_verify_compressed")
    assert p.n_inputs == 1273 and p.n_outputs == 1
    assert 11.0e9 < p.n_gates < 11.8e9 and 0.25 < p.n_ciphertexts / p.n_gates < 0.28
    good, bad = gsv.groth16_synthetic_inputs(424242, False), gsv.groth16_synthetic_inputs(424242, True)
    assert list(p.execute(good)) == [1]
    # the planned program (298 976 calls over recycled global slots) computes the same function
    assert list(p.execute_plan(good, lane_form=True)) == [1]
    assert list(p.execute_plan(bad, lane_form=True)) == [0]


@pytest.mark.parametrize("name", ["gate_zoo", "fq_add", "fq_mul", "fq12_mul", "fq_inverse", "g1_add"])
def test_planned_program_computes_the_recorded_function(gsv, name):
    """Planner self-check without a GPU: tasks, calls, task-local slots and recycled global slots, in
    the levelised and the emission-order form and in a lane-only plan, against ExecuteMode."""
    rng = np.random.default_rng(5)
    both, lane = gsv.Program(name), gsv.Program(name, lane_only=True)
    for _ in range(2):
        bits = rng.integers(0, 2, both.n_inputs, dtype=np.uint8)
        want = both.execute(bits)
        assert np.array_equal(both.execute_plan(bits, lane_form=False), want)
        assert np.array_equal(both.execute_plan(bits, lane_form=True), want)
        assert np.array_equal(lane.execute_plan(bits, lane_form=True), want)
    with pytest.raises(gsv.GsvError):
        lane.execute_plan(bits, lane_form=False)


def test_lane_only_plan_keeps_dependencies(gsv):
    """A lane-only plan (no levelised task form) must carry the same RAW edges between calls: the
    dependency-chain length of a serial circuit (Fq inverse: 2 x 254 dependent rounds) is a large
    fraction of its gates in both plans."""
    both = gsv.Program("fq_inverse")
    lane = gsv.Program("fq_inverse", lane_only=True)
    assert lane.n_gates == both.n_gates and lane.n_ciphertexts == both.n_ciphertexts
    assert both.critical_path_gates > both.n_gates // 8
    assert lane.critical_path_gates > lane.n_gates // 8
    assert lane.critical_path_levels == 0 and both.critical_path_levels > 0


@pytest.mark.parametrize("window", [1, 16, 64])
@pytest.mark.parametrize("name", ["fq_mul", "fq12_mul", "g1_add"])
def test_call_pipelining_plans(gsv, name, window):
    """Call pipelining (PlanOptions.pipeline): the same calls with START / DONE dependencies and window tables.  The
    host walker runs every call window by window -- inputs gathered at the window that first reads them, outputs
    published at the window that completes them -- and audits that every reader of a slot a call overwrites is one
    of its DONE dependencies; the critical path in levels can only shrink."""
    plain = gsv.Program(name, pipeline=False)
    piped = gsv.Program(name, pipeline=True, window_levels=window)
    assert (piped.n_calls, piped.n_gates, piped.n_ciphertexts) == (plain.n_calls, plain.n_gates, plain.n_ciphertexts)
    assert piped.critical_path_levels <= plain.critical_path_levels
    if plain.n_calls > 1:
        assert piped.critical_path_levels < plain.critical_path_levels
    rng = np.random.default_rng(9)
    for _ in range(2):
        bits = rng.integers(0, 2, plain.n_inputs, dtype=np.uint8)
        want = plain.execute(bits)
        assert np.array_equal(piped.execute_plan(bits, lane_form=False), want)
        assert np.array_equal(piped.execute_plan(bits, lane_form=True), want)


# ---- single steps of the pairing layer (pairing.rs, groth16.rs:26-47) against plain Fq2 / Fq12 arithmetic
def _f2(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % bn.P, (a[0] * b[1] + a[1] * b[0]) % bn.P)


def _f2inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % bn.P, -1, bn.P)
    return (a[0] * d % bn.P, (-a[1]) * d % bn.P)


def _f2add(a, b):
    return ((a[0] + b[0]) % bn.P, (a[1] + b[1]) % bn.P)


def _f2sub(a, b):
    return ((a[0] - b[0]) % bn.P, (a[1] - b[1]) % bn.P)


def _f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _f2(r, a)
        a = _f2(a, a)
        e >>= 1
    return r


G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))   # EIP-197
TWIST_B = _f2((3, 0), _f2inv((9, 1)))


def _on_twist(pt):
    return _f2(pt[1], pt[1]) == _f2add(_f2(_f2(pt[0], pt[0]), pt[0]), TWIST_B)


def _g2_affine_add(p, q):
    lam = _f2(_f2sub(q[1], p[1]), _f2inv(_f2sub(q[0], p[0]))) if p != q else \
        _f2(_f2((3, 0), _f2(p[0], p[0])), _f2inv(_f2add(p[1], p[1])))
    x = _f2sub(_f2sub(_f2(lam, lam), p[0]), q[0])
    return x, _f2sub(_f2(lam, _f2sub(p[0], x)), p[1])


def _fq2_bits(a):
    return _mont_bits(a[0]) + _mont_bits(a[1])


def _fq2_out(out, k):
    return (_from_mont(out[2 * k * 254:(2 * k + 1) * 254]), _from_mont(out[(2 * k + 1) * 254:(2 * k + 2) * 254]))


def test_g2_steps_functional(gsv):
    """double_in_place / add_in_place (pairing.rs:359-462) on homogeneous projective points and the twist Frobenius
    mul_by_char (pairing.rs:475-498): the results are the affine chord / tangent sums, respectively stay on the twist."""
    assert _on_twist(G2_GEN)
    z = (5, 7)                                                    # R = 2G in projective form with an arbitrary z
    g2x2 = _g2_affine_add(G2_GEN, G2_GEN)
    r = (_f2(g2x2[0], z), _f2(g2x2[1], z), z)
    dbl = gsv.Program("g2_double_step", lane_only=True)
    out = dbl.execute(_fq2_bits(r[0]) + _fq2_bits(r[1]) + _fq2_bits(r[2]))
    x, y, zz = (_fq2_out(out, k) for k in range(3))
    zi = _f2inv(zz)
    assert (_f2(x, zi), _f2(y, zi)) == _g2_affine_add(g2x2, g2x2)
    add = gsv.Program("g2_add_step", lane_only=True)
    q = G2_GEN
    out = add.execute(_fq2_bits(r[0]) + _fq2_bits(r[1]) + _fq2_bits(r[2]) + _fq2_bits(q[0]) + _fq2_bits(q[1]) + _fq2_bits((1, 0)))
    x, y, zz = (_fq2_out(out, k) for k in range(3))
    zi = _f2inv(zz)
    assert (_f2(x, zi), _f2(y, zi)) == _g2_affine_add(g2x2, G2_GEN)
    # line coefficients of the addition step: (lambda, -theta, theta qx - lambda qy) with theta = ry - qy rz, lambda = rx - qx rz
    theta, lam = _f2sub(r[1], _f2(q[1], r[2])), _f2sub(r[0], _f2(q[0], r[2]))
    assert _fq2_out(out, 3) == lam and _fq2_out(out, 4) == _f2sub((0, 0), theta)
    assert _fq2_out(out, 5) == _f2sub(_f2(theta, q[0]), _f2(lam, q[1]))
    frob = gsv.Program("g2_mul_by_char", lane_only=True)
    out = frob.execute(_fq2_bits(G2_GEN[0]) + _fq2_bits(G2_GEN[1]) + _fq2_bits((1, 0)))
    px, py = _fq2_out(out, 0), _fq2_out(out, 1)
    conj = lambda a: (a[0], (-a[1]) % bn.P)
    assert px == _f2(conj(G2_GEN[0]), _f2pow((9, 1), (bn.P - 1) // 3)) and py == _f2(conj(G2_GEN[1]), _f2pow((9, 1), (bn.P - 1) // 2))
    assert _on_twist((px, py)) and _fq2_out(out, 2) == (1, 0)


def test_g1_to_affine_functional(gsv):
    """projective_to_affine_montgomery (groth16.rs:26-47): Jacobian (X, Y, Z) -> (X / Z^2, Y / Z^3, 1)."""
    p = gsv.Program("g1_to_affine", lane_only=True)
    z = 0x1234567890ABCDEF1234567890ABCDEF % bn.P
    X, Y = 1 * z * z % bn.P, 2 * z * z * z % bn.P                 # the generator (1, 2)
    out = p.execute(_mont_bits(X) + _mont_bits(Y) + _mont_bits(z))
    assert [_from_mont(out[k * 254:(k + 1) * 254]) for k in range(3)] == [1, 2, 1]


def test_line_evaluation_functional(gsv):
    """ell / ell_by_constant (pairing.rs:160-171, 923-942): f * (c0 py + c1 px w^3 + c2 w^4) through the sparse
    034 products, against a dense Fq12 multiplication."""
    import random

    rng = random.Random(21)
    f = bn.rand_fq12(rng)
    px, py = bn.rand_fq(rng), bn.rand_fq(rng)
    co = [(bn.rand_fq(rng), bn.rand_fq(rng)) for _ in range(3)]

    def expect(co):
        c0 = (co[0][0] * py % bn.P, co[0][1] * py % bn.P)
        c3 = (co[1][0] * px % bn.P, co[1][1] * px % bn.P)
        sparse = ((c0, (0, 0), (0, 0)), (c3, co[2], (0, 0)))
        return bn.fq12_flatten(bn.fq12_mul(f, sparse))
    ell = gsv.Program("ell", lane_only=True)
    bits = bn.fq12_bits_mont(f) + sum((_fq2_bits(c) for c in co), []) + _mont_bits(px) + _mont_bits(py) + _mont_bits(1)
    out = ell.execute(bits)
    assert [_from_mont(out[k * 254:(k + 1) * 254]) for k in range(12)] == expect(co)
    const = ((3, 5), (7, 11), (13, 17))                            # the constants of the `ell_const` root
    ellc = gsv.Program("ell_const", lane_only=True)
    out = ellc.execute(bn.fq12_bits_mont(f) + _mont_bits(px) + _mont_bits(py) + _mont_bits(1))
    assert [_from_mont(out[k * 254:(k + 1) * 254]) for k in range(12)] == expect(const)

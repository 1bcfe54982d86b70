"""Row a8 / f1 at FULL size: the product's recorded circuits against the independent model, through a hash of the
component DAG (tests/golden/structural_hash.py) instead of the flattened stream -- equal hashes mean equal
flattened streams (gate order, types, wiring, dead gates) whatever the size.

The model side (tests/golden/emission_model.py, written from the reference's Rust gadgets) now covers the whole
Groth16 verifier: point decompression (Fp / Fp2 square roots), the windowed constant-base MSM, projective -> affine,
the Miller loop with two constant and one variable G2 point (its constant line coefficients computed by the model's own
host arithmetic), the final exponentiation and the comparison with the alpha-beta constant (the model's own host
pairing).  tests/golden/structural_hashes.json holds the model's hashes; the product's generator must reproduce them.
"""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import emission_model as em  # noqa: E402
import structural_hash as sh  # noqa: E402
import structural_roots as sr  # noqa: E402

with open(os.path.join(HERE, "golden", "structural_hashes.json")) as f:
    GOLDEN = json.load(f)["circuits"]

SLOW = pytest.mark.skipif(not os.environ.get("GSV_SLOW_TESTS"), reason="minutes of CPU; set GSV_SLOW_TESTS=1")


def _check_product(gsv, name):
    p = gsv.Program(name, lane_only=True)   # the recorded circuit does not depend on the plan
    assert p.n_inputs == GOLDEN[name]["n_inputs"]
    assert sh.product_hash(p) == GOLDEN[name]["structural_sha256"]


@pytest.mark.parametrize("name", ["fq_mul", "fq12_mul", "fq_inverse", "fq12_inverse", "g1_add", "g1_to_affine", "ell_const",
                                  "g2_add_step", "fq_sqrt", "fq2_sqrt", "decompress_g1", "g1_msm1", "final_exponentiation"])
def test_product_structure_matches_independent_model(gsv, name):
    _check_product(gsv, name)


def test_full_verifier_structure_matches_independent_model(lane_program):
    """groth16_verify_compressed, 11 457 232 209 gates: every gate, wire and dead output of the product's recorded
    circuit equals the independent model's."""
    p = lane_program("groth16_verify_compressed")
    assert p.n_inputs == GOLDEN["groth16_verify_compressed"]["n_inputs"] and p.n_gates == 11457232209
    assert sh.product_hash(p) == GOLDEN["groth16_verify_compressed"]["structural_sha256"]


@SLOW
@pytest.mark.parametrize("name", ["miller_loop_groth16", "groth16_verify"])
def test_large_structures_match_independent_model(gsv, name):
    _check_product(gsv, name)


@pytest.mark.parametrize("name", ["fq_mul", "g1_add", "g2_double_step", "fq_sqrt", "g1_msm1"])
def test_model_reproduces_committed_structural_hashes(name):
    n, fn = sr.ROOTS[name]
    assert sh.model_hash(fn, n) == GOLDEN[name]["structural_sha256"]


@SLOW
def test_model_reproduces_the_verifier_hash():
    n, fn = sr.ROOTS["groth16_verify_compressed"]
    assert sh.model_hash(fn, n) == GOLDEN["groth16_verify_compressed"]["structural_sha256"]


def test_structural_hash_agrees_with_the_flat_hash_and_sees_changes(gsv):
    """The DAG hash is a function of the flattened stream: circuits whose flat canonical hashes agree (test_emission_order)
    agree here, and single-gate / single-wire / liveness changes in a body change the root's hash."""
    n = em.N
    base = sh.model_hash(lambda x, w: em.fq_mul(x, w[:n], w[n:]), 2 * n)
    assert base == GOLDEN["fq_mul"]["structural_sha256"]
    # operands swapped at the root: another wiring of the same gates
    assert sh.model_hash(lambda x, w: em.fq_mul(x, w[n:], w[:n]), 2 * n) != base
    # one output dropped by the root: a dead gate somewhere below
    assert sh.model_hash(lambda x, w: em.fq_mul(x, w[:n], w[n:])[:-1], 2 * n) != base

    def one_gate_flipped(x, w):
        r = em.fq_mul(x, w[:n], w[n:])
        o = x.issue()
        x.gate(em.XOR, r[0], r[1], o)
        return r[:-1] + [o]

    def other_type(x, w):
        r = em.fq_mul(x, w[:n], w[n:])
        o = x.issue()
        x.gate(em.XNOR, r[0], r[1], o)
        return r[:-1] + [o]
    assert sh.model_hash(one_gate_flipped, 2 * n) != sh.model_hash(other_type, 2 * n)

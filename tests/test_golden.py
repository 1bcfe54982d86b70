"""Committed golden vectors (tests/golden/garble_vectors.json, made by tests/golden/make_golden.py).

CPU: the oracle still produces them.  GPU: the CUDA path produces them through the C ABI, with the oracle
not loaded at all -- the fixture, not a live CPU run, is the target."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "garble_vectors.json")) as f:
    VECTORS = json.load(f)["vectors"]
HASH_ID = {"aes": 0, "blake3": 1}
GROUPS = sorted({(v["circuit"], v["hasher"]) for v in VECTORS})


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check(v, delta, fl, tl, il, ol, cts, commit):
    assert bytes(delta).hex() == v["delta"]
    assert bytes(fl).hex() == v["false_label0"] and bytes(tl).hex() == v["true_label0"]
    assert bytes(il[0]).hex() == v["input_label0_first"] and _sha(il) == v["input_label0_sha256"]
    assert bytes(ol[0]).hex() == v["output_label0_first"] and _sha(ol) == v["output_label0_sha256"]
    assert len(cts) == v["n_ciphertexts"] and _sha(cts) == v["ct_stream_sha256"]
    if v["ct_first"]:
        assert bytes(cts[0]).hex() == v["ct_first"]
    assert bytes(commit).hex() == v["ct_commit"]


def test_fixture_covers_reference_seeds_and_hashers():
    assert {v["seed"] for v in VECTORS} == {0, 42, 1234}
    assert {v["hasher"] for v in VECTORS} == {"aes", "blake3"}
    fq12 = [v for v in VECTORS if v["circuit"] == "fq12_mul"]
    assert fq12 and all(v["n_gates"] == 20284982 and v["n_ciphertexts"] == 5439206 for v in fq12)


@pytest.mark.parametrize("name,hasher", GROUPS)
def test_oracle_reproduces_golden(gsv, orc, circuit, name, hasher):
    p, st = circuit(name)
    for v in (v for v in VECTORS if v["circuit"] == name and v["hasher"] == hasher):
        assert p.n_gates == v["n_gates"]
        r = st.garble(HASH_ID[hasher], v["seed"])
        _check(v, r["delta"], r["false_label0"], r["true_label0"], r["input_label0"], r["output_label0"], r["cts"],
               r["ct_commit"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,hasher", GROUPS)
def test_cuda_path_reproduces_golden(gsv, name, hasher):
    vs = [v for v in VECTORS if v["circuit"] == name and v["hasher"] == hasher]
    p = gsv.Program(name)
    for mode in (1, 2):  # levelised, lane
        sess = gsv.Session(p, len(vs), ct_mode=gsv.CT_KEEP, exec_mode=mode)
        res = sess.garble([v["seed"] for v in vs], HASH_ID[hasher])
        for i, v in enumerate(vs):
            _check(v, res.delta[i], res.false_label0[i], res.true_label0[i], res.input_label0[i], res.output_label0[i],
                   sess.read_ciphertexts(i), res.ct_commit[i])


# ---- the full Groth16 verifier --------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gate_zoo", "fq_mul", "fq12_mul", "g1_add"])
def test_template_walk_oracle_equals_flat_stream_oracle(gsv, orc, circuit, name):
    """The oracle's depth-first walk over the exported template DAG (what pins the verifier, too large to
    flatten) is the same garbling as its flat-stream loop: same labels, ciphertext count, commitment."""
    p, st = circuit(name)
    dag = orc.TemplateDag(*p.export_templates())
    assert dag.n_inputs == p.n_inputs and dag.n_outputs == p.n_outputs
    for hasher in (0, 1):
        a, b = dag.garble(hasher, 1234), st.garble(hasher, 1234, want_ct=False)
        assert a["n_gates"] == p.n_gates and a["n_ct"] == p.n_ciphertexts
        assert a["ct_commit"] == b["ct_commit"] and a["delta"] == b["delta"]
        assert np.array_equal(a["input_label0"], b["input_label0"])
        assert np.array_equal(a["output_label0"], b["output_label0"])


def _verifier_vectors():
    with open(os.path.join(HERE, "golden", "verifier_vectors.json")) as f:
        return json.load(f)


def test_verifier_fixture_shape():
    d = _verifier_vectors()
    assert d["n_gates"] == 11457232209 and d["n_ciphertexts"] == 2980239027
    assert [v["seed"] for v in d["vectors"]] == [1234, 1235]


@pytest.mark.gpu
def test_full_verifier_garbling_matches_oracle_fixture(gsv):
    """BASELINE.json config 2 at full size: the 11.46 G-gate Groth16 verifier garbled on the GPU (levelised
    kernel, host-folded commitment) gives the CPU oracle's delta, constants, first input label, output label
    and -- over all 2 980 239 027 ciphertexts -- chain commitment (tests/golden/make_verifier_golden.py)."""
    d = _verifier_vectors()
    p = gsv.Program("groth16_verify_compressed")
    assert p.n_gates == d["n_gates"] and p.n_ciphertexts == d["n_ciphertexts"]
    vs = d["vectors"]
    sess = gsv.Session(p, len(vs), ct_mode=gsv.CT_COMMIT_HOST, exec_mode=1, group=2)
    res = sess.garble([v["seed"] for v in vs], gsv.HASH_AES)
    for i, v in enumerate(vs):
        assert bytes(res.delta[i]).hex() == v["delta"]
        assert bytes(res.false_label0[i]).hex() == v["false_label0"] and bytes(res.true_label0[i]).hex() == v["true_label0"]
        assert bytes(res.input_label0[i, 0]).hex() == v["input_label0_first"]
        assert bytes(res.output_label0[i, 0]).hex() == v["output_label0"]
        assert bytes(res.ct_commit[i]).hex() == v["ct_commit"]
    # ExecuteMode on the GPU over the same planned program (the reference's pre-check,
    # examples/groth16_cut_and_choose.rs:235-254): the synthetic proof verifies, its tampered variant does not
    bits = np.stack([gsv.groth16_synthetic_inputs(), gsv.groth16_synthetic_inputs(flip_public=True)] * 3)
    out, ms = sess.execute(bits)
    assert out[:, 0].tolist() == [1, 0, 1, 0, 1, 0]
    assert ms < 60e3

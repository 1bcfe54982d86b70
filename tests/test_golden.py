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

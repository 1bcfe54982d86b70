"""GPU parity: every ciphertext, label, commitment and decoded bit produced through the C ABI on
the B200 must equal the CPU oracle's on the same seeds (bit-exact; integer/byte work)."""
import random

import os

import numpy as np
import pytest

import bn254_ref as bn

pytestmark = pytest.mark.gpu


def test_native_library_on_gpu(gsv):
    assert gsv.device_count() >= 1


@pytest.mark.parametrize("hasher", [0, 1])
def test_device_gate_hash(gsv, orc, hasher):
    rng = np.random.default_rng(1)
    n = 4096
    x = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    gid = rng.integers(0, 2**63, n, dtype=np.uint64)
    gid[:8] = [0, 1, 7, 2**32 - 1, 2**32, 2**32 + 5, 11_174_708_820, 2**63 + 17]
    got = gsv.hash_blocks(hasher, x, gid)
    for i in range(0, n, 7):
        assert bytes(got[i]) == orc.hash_gate(hasher, bytes(x[i]), int(gid[i])), i


def test_device_label_commit(gsv, orc):
    rng = np.random.default_rng(2)
    lab = rng.integers(0, 256, (1000, 16), dtype=np.uint8)
    got = gsv.commit_labels(lab)
    for i in range(0, 1000, 13):
        assert bytes(got[i]) == orc.commit_label(bytes(lab[i]))


def _check_garble(gsv, orc, circuit, name, hasher, seeds, group, worker_threads=0, every_ct=True):
    p, st = circuit(name)
    sess = gsv.Session(p, len(seeds), group=group, worker_threads=worker_threads, ct_mode=gsv.CT_KEEP)
    res = sess.garble(seeds, hasher)
    assert res.n_ciphertexts == p.n_ciphertexts
    refs = []
    for i, seed in enumerate(seeds):
        ref = st.garble(hasher, seed)
        refs.append(ref)
        assert bytes(res.delta[i]) == ref["delta"]
        assert bytes(res.false_label0[i]) == ref["false_label0"]
        assert bytes(res.true_label0[i]) == ref["true_label0"]
        assert np.array_equal(res.input_label0[i], ref["input_label0"])
        assert np.array_equal(res.output_label0[i], ref["output_label0"]), (name, hasher, seed)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"], (name, hasher, seed)
        if every_ct:
            assert np.array_equal(sess.read_ciphertexts(i), ref["cts"])
    return p, st, sess, res, refs


@pytest.mark.parametrize("hasher", [0, 1])
@pytest.mark.parametrize("name", ["gate_zoo", "fq_add", "fq_mul", "fq_expr"])
def test_garble_matches_oracle(gsv, orc, circuit, name, hasher):
    # the reference tests' seeds
    _check_garble(gsv, orc, circuit, name, hasher, [0, 42, 99, 777], group=2)


@pytest.mark.parametrize("group,wt", [(1, 256), (4, 256), (4, 128), (2, 512), (1, 64)])
def test_garble_group_and_worker_shapes(gsv, orc, circuit, group, wt):
    seeds = [1234, 12345, 2**64 - 1, 5, 6, 7, 8, 9]
    _check_garble(gsv, orc, circuit, "fq_mul", 0, seeds, group=group, worker_threads=wt)


@pytest.mark.parametrize("hasher", [0, 1])
@pytest.mark.parametrize("name,B", [("gate_zoo", 5), ("fq_add", 33), ("fq_mul", 64), ("fq_expr", 40)])
def test_lane_mode_garble_matches_oracle(gsv, orc, circuit, name, B, hasher):
    """Lane mode (one warp = 32 instances, emission order): full, partial and multiple groups."""
    p, st = circuit(name)
    seeds = [0, 42, 99, 777] + list(range(1000, 1000 + B - 4))
    sess = gsv.Session(p, B, ct_mode=gsv.CT_KEEP, exec_mode=2)
    res = sess.garble(seeds, hasher)
    for i in sorted({0, 1, 2, 3, 31 % B, 32 % B, B - 1}):
        ref = st.garble(hasher, seeds[i])
        assert bytes(res.delta[i]) == ref["delta"]
        assert np.array_equal(res.input_label0[i], ref["input_label0"])
        assert np.array_equal(res.output_label0[i], ref["output_label0"]), (name, i)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"], (name, i)
        assert np.array_equal(sess.read_ciphertexts(i), ref["cts"])


def test_lane_mode_evaluate_and_ring(gsv, orc, circuit):
    p, st = circuit("fq_mul")
    B = 48
    seeds = list(range(300, 300 + B))
    sess = gsv.Session(p, B, ct_mode=gsv.CT_KEEP, exec_mode=2)
    res = sess.garble(seeds, gsv.HASH_AES)
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, (B, p.n_inputs), dtype=np.uint8)
    act = _eval_inputs(res, bits)
    ev = sess.evaluate(gsv.HASH_AES, res.true_label1, res.false_label0, act, bits)
    assert np.array_equal(ev.ct_commit, res.ct_commit)
    for i in (0, 31, 32, 47):
        ref = st.garble(orc.HASH_AES, seeds[i])
        o = st.evaluate(orc.HASH_AES, bytes(res.true_label1[i]), bytes(res.false_label0[i]), act[i], bits[i], ref["cts"])
        assert np.array_equal(ev.output_active[i], o["output_active"])
        assert np.array_equal(ev.output_bits[i], o["output_bits"])
    # commitment through a small ring (2^17, not the whole 102 093-ciphertext... use fq12 for a real wrap)
    ring = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT, exec_mode=2, ct_ring_log2=17).garble(seeds, gsv.HASH_AES)
    assert np.array_equal(ring.ct_commit, res.ct_commit)
    assert np.array_equal(ring.output_label0, res.output_label0)


def test_lane_mode_fq12_mul(gsv, orc, circuit):
    p, st = circuit("fq12_mul")
    B = 64
    seeds = [0, 42] + list(range(7000, 7000 + B - 2))
    res = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT, exec_mode=2, ct_ring_log2=18).garble(seeds, gsv.HASH_AES)  # ring wraps 20x
    for i in (0, 1, 63):
        ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"]
        assert np.array_equal(res.output_label0[i], ref["output_label0"])


def test_garble_ragged_batch(gsv, orc, circuit):
    # odd instance count forces group 1; single instance
    _check_garble(gsv, orc, circuit, "fq_add", 0, [3, 4, 5], group=0)
    _check_garble(gsv, orc, circuit, "fq_add", 1, [9], group=0)


@pytest.mark.parametrize("hasher", [0, 1])
def test_fq12_mul_garble_matches_oracle(gsv, orc, circuit, hasher):
    """config 1 (tests/fq12_mul_e2e.rs, seed 0): all 5.44 M ciphertexts, 3048 output labels,
    the chain commitment."""
    _check_garble(gsv, orc, circuit, "fq12_mul", hasher, [0, 42], group=2)


def _eval_inputs(res, bits):
    """EvaluatedWire::new_from_garbled for every input (evaluate_mode.rs:33-38)."""
    act = res.input_label0.copy()
    act[np.asarray(bits, bool)] ^= np.broadcast_to(res.delta[:, None, :], act.shape)[np.asarray(bits, bool)]
    return act


@pytest.mark.parametrize("hasher", [0, 1])
@pytest.mark.parametrize("name", ["gate_zoo", "fq_mul"])
def test_evaluate_matches_oracle(gsv, orc, circuit, name, hasher):
    seeds = [0, 42, 99, 777]
    p, st, sess, res, refs = _check_garble(gsv, orc, circuit, name, hasher, seeds, group=2, every_ct=False)
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, (len(seeds), p.n_inputs), dtype=np.uint8)
    act = _eval_inputs(res, bits)
    ev = sess.evaluate(hasher, res.true_label1, res.false_label0, act, bits)  # the session's own stream
    streams = [sess.read_ciphertexts(i) for i in range(len(seeds))]
    sess2 = gsv.Session(p, len(seeds), group=1, ct_mode=gsv.CT_KEEP)  # FileSource-style host streams
    ev2 = sess2.evaluate(hasher, res.true_label1, res.false_label0, act, bits, ct_streams=streams)
    for i in range(len(seeds)):
        o = st.evaluate(hasher, bytes(res.true_label1[i]), bytes(res.false_label0[i]), act[i], bits[i], refs[i]["cts"])
        assert o["rc"] == 0
        for e in (ev, ev2):
            assert np.array_equal(e.output_active[i], o["output_active"])
            assert np.array_equal(e.output_bits[i], o["output_bits"])
            assert bytes(e.ct_commit[i]) == o["ct_commit"] == refs[i]["ct_commit"]
        # gw.select(value) == active_label (tests/fq12_mul_e2e.rs:217-235)
        want = res.output_label0[i].copy()
        want[ev.output_bits[i].astype(bool)] ^= res.delta[i]
        assert np.array_equal(ev.output_active[i], want)


def test_evaluate_ciphertext_exhaustion(gsv, circuit):
    p, st = circuit("fq_add")
    sess = gsv.Session(p, 2, ct_mode=gsv.CT_KEEP)
    res = sess.garble([1, 2], 0)
    bits = np.zeros((2, p.n_inputs), np.uint8)
    streams = [sess.read_ciphertexts(i)[:-1] for i in range(2)]
    with pytest.raises(gsv.GsvError) as e:
        sess.evaluate(0, res.true_label1, res.false_label0, res.input_label0, bits, ct_streams=streams)
    assert e.value.code == -5  # "Ciphertext source exhausted"


def test_fq12_mul_e2e_batch_properties(gsv, circuit):
    """Full-size, size-independent checks on a 32-instance batch: garble -> evaluate round trip,
    decoded product == plain BN254 arithmetic, evaluator's stream hash == garbler's commitment,
    distinct seeds give distinct commitments, same seed is reproducible."""
    p, _ = circuit("fq12_mul")
    B = 32
    seeds = list(range(100, 100 + B - 1)) + [100]
    sess = gsv.Session(p, B, group=2, ct_mode=gsv.CT_KEEP)
    res = sess.garble(seeds, gsv.HASH_AES)
    assert bytes(res.ct_commit[0]) == bytes(res.ct_commit[B - 1])
    assert len({bytes(c) for c in res.ct_commit}) == B - 1
    rng = random.Random(5)
    bits = np.zeros((B, p.n_inputs), np.uint8)
    want = []
    for i in range(B):
        a, b = bn.rand_fq12(rng), bn.rand_fq12(rng)
        bits[i] = bn.fq12_bits_mont(a) + bn.fq12_bits_mont(b)
        want.append(bn.fq12_bits_mont(bn.fq12_mul(a, b)))
    ev = sess.evaluate(gsv.HASH_AES, res.true_label1, res.false_label0, _eval_inputs(res, bits), bits)
    assert np.array_equal(ev.output_bits, np.array(want, np.uint8))
    sel = res.output_label0.copy()
    sel[ev.output_bits.astype(bool)] ^= np.broadcast_to(res.delta[:, None, :], sel.shape)[ev.output_bits.astype(bool)]
    assert np.array_equal(ev.output_active, sel)
    assert np.array_equal(ev.ct_commit, res.ct_commit)


def test_commit_ring_backpressure(gsv, orc, circuit):
    """GSV_CT_COMMIT with a ring far smaller than the stream (2^17 of 5.44 M ciphertexts): the
    garbling workers must stall on the chain warps' progress and the commitment stay bit-exact."""
    p, st = circuit("fq12_mul")
    seeds = [0, 42, 99, 777, 1, 2, 3, 4]
    sess = gsv.Session(p, len(seeds), group=2, ct_mode=gsv.CT_COMMIT, ct_ring_log2=17)
    res = sess.garble(seeds, gsv.HASH_AES)
    for i in (0, 1, 7):
        ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"]
        assert np.array_equal(res.output_label0[i], ref["output_label0"])
    with pytest.raises(gsv.GsvError):
        sess.read_ciphertexts(0)  # a ring keeps no stream


def test_many_instances_chain_warps(gsv, orc, circuit):
    """More than 32 instances per SM-resident chain warp set: 96 instances, odd chain-warp fill."""
    p, st = circuit("fq_mul")
    B = 96 + 8
    seeds = list(range(5000, 5000 + B))
    res = gsv.Session(p, B, group=4, ct_mode=gsv.CT_COMMIT).garble(seeds, gsv.HASH_AES)
    for i in (0, 31, 32, 95, 96, B - 1):
        assert bytes(res.ct_commit[i]) == st.garble(orc.HASH_AES, seeds[i], want_ct=False)["ct_commit"]


def test_commit_only_mode_matches_keep(gsv, circuit):
    p, _ = circuit("fq_mul")
    seeds = [7, 8, 9, 10]
    a = gsv.Session(p, 4, ct_mode=gsv.CT_KEEP).garble(seeds, 0)
    b = gsv.Session(p, 4, ct_mode=gsv.CT_COMMIT).garble(seeds, 0)
    c = gsv.Session(p, 4, ct_mode=gsv.CT_NONE).garble(seeds, 0)
    assert np.array_equal(a.ct_commit, b.ct_commit)
    assert np.array_equal(a.output_label0, b.output_label0) and np.array_equal(a.output_label0, c.output_label0)


def test_cut_and_choose_commit_records(gsv, orc, circuit):
    """Garbler::create + commit on the GPU == GarbledInstanceCommit::new over the oracle's labels
    (garbler.rs:85-116: ct commit, (c(l0), c(l1)) per input, output label1/label0, constants)."""
    from gsv_b200 import cut_and_choose as cc

    p, st = circuit("fq_add")
    total = 6
    g = cc.Garbler(p, total, master_seed=1234)
    g.create()
    rec = g.commit()
    seeds = cc.instance_seeds(1234, total)
    xor = lambda a, b: bytes(x ^ y for x, y in zip(a, b))
    for i in (0, 3, 5):
        ref = st.garble(orc.HASH_AES, int(seeds[i]), want_ct=False)
        d = ref["delta"]
        assert bytes(rec.ct_commit()[i]) == ref["ct_commit"]
        for j in (0, 1, p.n_inputs - 1):
            l0 = bytes(ref["input_label0"][j])
            assert bytes(rec.input_commits()[i, j, 0]) == orc.commit_label(l0)
            assert bytes(rec.input_commits()[i, j, 1]) == orc.commit_label(xor(l0, d))
        for j in (0, p.n_outputs - 1):
            o0 = bytes(ref["output_label0"][j])
            assert bytes(rec.output_commits()[i, j, 0]) == orc.commit_label(xor(o0, d))  # label1 first
            assert bytes(rec.output_commits()[i, j, 1]) == orc.commit_label(o0)
        assert bytes(rec.constant_commits()[i, 0]) == orc.commit_label(xor(ref["true_label0"], d))
        assert bytes(rec.constant_commits()[i, 1]) == orc.commit_label(ref["false_label0"])


@pytest.mark.parametrize("name,B,mode", [("fq_inverse", 32, 2), ("fq_inverse", 4, 1), ("g1_add", 64, 2), ("fq12_square", 4, 1)])
def test_pairing_gadget_circuits_match_oracle(gsv, orc, circuit, name, B, mode):
    """Sub-circuits of the Groth16 verifier (inverse: 32 k calls, recycled global slots with WAR
    dependencies): commitment and output labels vs the oracle in both execution modes."""
    p, st = circuit(name)
    seeds = [0, 42, 99, 777] + list(range(50, 50 + B - 4))
    res = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT, exec_mode=mode, group=2 if mode == 1 else 0).garble(seeds, gsv.HASH_AES)
    for i in (0, 3, B - 1):
        ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"], (name, i)
        assert np.array_equal(res.output_label0[i], ref["output_label0"])


@pytest.mark.parametrize("mode", [1, 2])
def test_keep_raw_stream_evaluates_without_commit(gsv, orc, circuit, mode):
    """GSV_CT_KEEP_RAW: the stream stays in HBM for the evaluator, no chain is folded on either side
    (the `()` ciphertext handler / evaluator without a commitment check)."""
    p, st = circuit("fq_mul")
    B = 4
    seeds = [5, 6, 7, 8]
    sess = gsv.Session(p, B, ct_mode=gsv.CT_KEEP_RAW, exec_mode=mode, group=2 if mode == 1 else 0)
    res = sess.garble(seeds, gsv.HASH_AES)
    bits = np.random.default_rng(11).integers(0, 2, (B, p.n_inputs), dtype=np.uint8)
    act = _eval_inputs(res, bits)
    ev = sess.evaluate(gsv.HASH_AES, res.true_label1, res.false_label0, act, bits, want_commit=False)
    for i in range(B):
        ref = st.garble(orc.HASH_AES, seeds[i])
        assert np.array_equal(sess.read_ciphertexts(i), ref["cts"])
        o = st.evaluate(orc.HASH_AES, bytes(res.true_label1[i]), bytes(res.false_label0[i]), act[i], bits[i], ref["cts"])
        assert np.array_equal(ev.output_active[i], o["output_active"])
        assert np.array_equal(ev.output_bits[i], o["output_bits"])


def test_lane_only_program_matches_oracle(gsv, orc, circuit):
    """Programs planned with lane_only (how the 11 G-gate verifier is planned for large batches) carry
    only the emission-order task form; same commitment and labels."""
    _, st = circuit("fq_inverse")
    p = gsv.Program("fq_inverse", lane_only=True)
    B = 64
    seeds = list(range(900, 900 + B))
    res = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT, exec_mode=2).garble(seeds, gsv.HASH_AES)
    for i in (0, 33, 63):
        ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
        assert bytes(res.ct_commit[i]) == ref["ct_commit"]
        assert np.array_equal(res.output_label0[i], ref["output_label0"])
    with pytest.raises(gsv.GsvError):
        gsv.Session(p, 2, exec_mode=1)  # no levelised form in this plan


@pytest.mark.parametrize("mode,B", [(1, 4), (1, 3), (2, 40)])
def test_host_folded_commitment_matches_gpu_chain(gsv, orc, circuit, mode, B, monkeypatch):
    """GSV_CT_COMMIT_HOST: gate hashes on the GPU, the serial chain folded by host AES-NI threads
    draining the ring (small ring and 1 MB drain buffers so both wrap many times)."""
    monkeypatch.setenv("GSV_HOST_CHAIN_BUF_MB", "1")
    p, st = circuit("fq12_mul")
    seeds = [0, 42] + list(range(10, 10 + B - 2))
    host = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT_HOST, exec_mode=mode, ct_ring_log2=19).garble(seeds, gsv.HASH_AES)
    dev = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT, exec_mode=mode, ct_ring_log2=19).garble(seeds, gsv.HASH_AES)
    assert np.array_equal(host.ct_commit, dev.ct_commit)
    assert np.array_equal(host.output_label0, dev.output_label0)
    for i in (0, 1):
        ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
        assert bytes(host.ct_commit[i]) == ref["ct_commit"]
    monkeypatch.delenv("GSV_HOST_CHAIN_BUF_MB")
    again = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT_HOST, exec_mode=mode).garble(seeds, gsv.HASH_AES)  # whole stream resident
    assert np.array_equal(again.ct_commit, dev.ct_commit)


def test_cut_and_choose_protocol_round_trip(gsv, circuit):
    """The whole flow of src/cut_and_choose/tests.rs (total 5 / finalize 2) on Fq mul: create + commit,
    evaluator picks, open_commit, run_regarbling, prepare_input_labels, evaluate_from; then the reference's
    tamper cases (corrupted ciphertext, wrong seed, wrong input label, wrong constant)."""
    import importlib

    cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
    import bn254_ref as bn

    p, _ = circuit("fq_mul")
    total, fin = 5, 2
    garbler = cc.Garbler(p, total, master_seed=1234)
    garbler.create()
    commits = garbler.commit()
    ev = cc.Evaluator(p, total, fin, rng_seed=99, commits=commits)
    assert len(ev.to_finalize) == fin
    open_, closed = cc.open_commit(garbler, ev.to_finalize)
    assert sorted(i for i, _ in open_) == [i for i in range(total) if i not in ev.to_finalize]
    assert sorted(closed) == ev.to_finalize and all(c.shape == (p.n_ciphertexts, 16) for c in closed.values())
    ev.run_regarbling(open_, closed)

    a, b = 123456789123456789 % bn.P, (bn.P - 987654321)
    bits = np.array(bn.bits_le(bn.to_mont(a)) + bn.bits_le(bn.to_mont(b)), np.uint8)
    cases = cc.prepare_input_labels(garbler, ev.to_finalize, bits)
    res = ev.evaluate_from(closed, cases)
    assert [i for i, _, _ in res] == ev.to_finalize
    for _, out_bits, _ in res:
        assert bn.from_bits(list(out_bits)) == bn.to_mont(a * b % bn.P)

    # --- tamper cases
    bad = {i: c.copy() for i, c in closed.items()}
    bad[ev.to_finalize[0]][1000, 3] ^= 1
    with pytest.raises(cc.ConsistencyError) as e:
        ev.run_regarbling(open_, bad)
    assert e.value.kind == "CiphertextMismatch" and e.value.index == ev.to_finalize[0]
    with pytest.raises(cc.ConsistencyError) as e:
        ev.evaluate_from(bad, cases)
    assert e.value.kind == "CiphertextMismatch"
    wrong = [(i, s ^ 1) if k == 1 else (i, s) for k, (i, s) in enumerate(open_)]
    with pytest.raises(cc.ConsistencyError) as e:
        ev.run_regarbling(wrong, closed)
    assert e.value.kind == "RegarblingMismatch" and e.value.index == open_[1][0]
    c0 = cases[0]
    flipped = c0.input_active.copy()
    flipped[7, 0] ^= 0x80
    with pytest.raises(cc.ConsistencyError) as e:
        ev.evaluate_from(closed, [cc.EvaluatorCaseInput(c0.index, flipped, c0.input_bits, c0.true_label, c0.false_label)] + cases[1:])
    assert e.value.kind == "InputLabelsMismatch" and "label_index 7" in str(e.value)
    with pytest.raises(cc.ConsistencyError) as e:
        ev.evaluate_from(closed, [cc.EvaluatorCaseInput(c0.index, c0.input_active, c0.input_bits, c0.false_label, c0.false_label)] + cases[1:])
    assert e.value.kind == "TrueConstantMismatch"


def test_cut_and_choose_with_ciphertext_files(gsv, circuit, tmp_path):
    """The same protocol with the reference's file handlers (FileCiphertextHandlerProvider / FileSource,
    ciphertext_repository.rs:59-136): finalized instances are re-garbled straight into gc_{i}.bin by the host drain,
    the evaluator folds and evaluates from the files; a flipped byte in a file is caught."""
    import importlib

    cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
    import bn254_ref as bn

    p, _ = circuit("fq_mul")
    total, fin = 4, 2
    garbler = cc.Garbler(p, total, master_seed=7)
    garbler.create()
    commits = garbler.commit()
    ev = cc.Evaluator(p, total, fin, rng_seed=3, commits=commits)
    open_, closed = cc.open_commit(garbler, ev.to_finalize, ct_dir=str(tmp_path))
    assert all(isinstance(v, str) and os.path.getsize(v) == 16 * p.n_ciphertexts for v in closed.values())
    ev.run_regarbling(open_, closed)
    a, b = 5, bn.P - 2
    bits = np.array(bn.bits_le(bn.to_mont(a)) + bn.bits_le(bn.to_mont(b)), np.uint8)
    res = ev.evaluate_from(closed, cc.prepare_input_labels(garbler, ev.to_finalize, bits))
    for _, out_bits, _ in res:
        assert bn.from_bits(list(out_bits)) == bn.to_mont(a * b % bn.P)
    victim = closed[ev.to_finalize[1]]
    with open(victim, "r+b") as f:
        f.seek(16 * 777 + 5)
        byte = f.read(1)
        f.seek(16 * 777 + 5)
        f.write(bytes([byte[0] ^ 0x10]))
    with pytest.raises(cc.ConsistencyError) as e:
        ev.run_regarbling(open_, closed)
    assert e.value.kind == "CiphertextMismatch" and e.value.index == ev.to_finalize[1]


@pytest.mark.parametrize("mode,B,ring_mb", [(1, 4, 12), (1, 6, 0), (2, 40, 96)])
def test_linked_garbler_evaluator_stream(gsv, orc, circuit, mode, B, ring_mb, monkeypatch):
    """Garbler -> evaluator streaming (examples/groth16_garble.rs:170-267, tests/garbler_evaluator_connection.rs): the
    garbler's kernel fills a ring in the evaluator's memory (the second GPU when there is one, peer stores), the
    evaluator consumes it behind the progress words and hashes what it received.  Checked against the oracle:
    evaluator's chain hash == the garbler's commitment, active output labels == select(value), bits == plaintext."""
    monkeypatch.setenv("GSV_HOST_CHAIN_BUF_MB", "1")
    p, st = circuit("fq12_mul")
    two = gsv.device_count() >= 2
    sm = 0 if two else 64   # one GPU: the two persistent grids share the SMs
    gs = gsv.Session(p, B, device=0, ct_mode=gsv.CT_NONE, exec_mode=mode, group=2 if mode == 1 else 0, sm_limit=sm)
    es = gsv.Session(p, B, device=1 if two else 0, ct_mode=gsv.CT_NONE, exec_mode=mode, group=2 if mode == 1 else 0, sm_limit=sm)
    gsv.link_sessions(gs, es, ring_bytes=ring_mb << 20)
    rng = np.random.default_rng(5)
    for run in range(2):   # a link serves repeated runs
        seeds = [0, 42] + list(range(100 * run + 10, 100 * run + 10 + B - 2))
        bits = rng.integers(0, 2, (B, p.n_inputs), dtype=np.uint8)
        gres, ev = gsv.stream_garble_evaluate(gs, es, seeds, gsv.HASH_AES, bits)
        for i in (0, 1, B - 1):
            ref = st.garble(orc.HASH_AES, seeds[i], want_ct=False)
            assert bytes(ev.ct_commit[i]) == ref["ct_commit"]
            want_bits = st.execute(bits[i])
            assert np.array_equal(ev.output_bits[i], want_bits)
            delta = np.frombuffer(ref["delta"], np.uint8)
            assert np.array_equal(ev.output_active[i], ref["output_label0"] ^ (delta[None, :] * want_bits[:, None]))
            assert np.array_equal(gres.output_label0[i], ref["output_label0"])


def test_ciphertext_files_from_host_drain(gsv, orc, circuit, tmp_path, monkeypatch):
    """FileCiphertextHandler / FileSource (ciphertext_repository.rs:59-136, ciphertext_source.rs:35-106): a
    GSV_CT_COMMIT_HOST run writes gc_{i}.bin straight from the drain buffers (small ring and drain buffers, so the
    files are assembled from many wrapped chunks); the bytes equal the oracle's stream and evaluate correctly."""
    monkeypatch.setenv("GSV_HOST_CHAIN_BUF_MB", "1")
    p, st = circuit("fq12_mul")
    B = 5
    seeds = [0, 42, 7, 8, 9]
    paths = [str(tmp_path / f"gc_{i}.bin") if i != 3 else None for i in range(B)]   # instance 3: not kept
    sess = gsv.Session(p, B, ct_mode=gsv.CT_COMMIT_HOST, exec_mode=1, group=1, ct_ring_log2=19)
    sess.set_ciphertext_files(paths)
    res = sess.garble(seeds, gsv.HASH_AES)
    sess.set_ciphertext_files(None)
    ref = st.garble(orc.HASH_AES, seeds[1])
    assert bytes(res.ct_commit[1]) == ref["ct_commit"]
    got = np.fromfile(paths[1], np.uint8).reshape(-1, 16)
    assert got.shape[0] == p.n_ciphertexts and np.array_equal(got, ref["cts"])
    assert not (tmp_path / "gc_3.bin").exists()
    # evaluate instance 4 from its file (FileSource), alone in a one-instance session
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2, (1, p.n_inputs), dtype=np.uint8)
    active = res.input_label0[4:5] ^ (res.delta[4:5, None, :] * bits[:, :, None])
    ev = gsv.Session(p, 1, ct_mode=gsv.CT_NONE, exec_mode=1, group=1).evaluate(
        gsv.HASH_AES, res.true_label1[4:5], res.false_label0[4:5], active, bits,
        ct_streams=[np.fromfile(paths[4], np.uint8)])
    want = st.execute(bits[0])
    assert np.array_equal(ev.output_bits[0], want)
    assert np.array_equal(ev.output_active[0], res.output_label0[4] ^ (res.delta[4][None, :] * want[:, None]))
    assert np.array_equal(ev.ct_commit[0], res.ct_commit[4])


@pytest.mark.parametrize("name,B,G,n", [("gate_zoo", 1, 1, 128), ("fq_mul", 2, 2, 200), ("fq12_mul", 4, 4, 300)])
def test_execute_mode_on_gpu(gsv, circuit, name, B, G, n):
    """ExecuteMode on the GPU (bit-sliced 128 executions per instance slot) == the host walker of the recorded
    circuit (gsv_program_execute) and, for Fq mul, the BN254 product."""
    p, st = circuit(name)
    rng = np.random.default_rng(21)
    bits = rng.integers(0, 2, (n, p.n_inputs), dtype=np.uint8)
    sess = gsv.Session(p, B, ct_mode=gsv.CT_NONE, exec_mode=1, group=G)
    out, ms = sess.execute(bits)
    for e in (0, 1, 31, 32, 127, n - 1):
        assert np.array_equal(out[e], st.execute(bits[e])), e
    # garbling afterwards still works on the same session (label slots are re-seeded)
    res = sess.garble(list(range(B)), gsv.HASH_AES)
    ref = st.garble(0, 0, want_ct=False)
    assert np.array_equal(res.output_label0[0], ref["output_label0"])
    with pytest.raises(gsv.GsvError):
        sess.execute(np.zeros((128 * B + 1, p.n_inputs), np.uint8))


@pytest.mark.parametrize("mode", [1, 2])
def test_evaluate_fed_from_host_streams_through_a_ring(gsv, orc, circuit, tmp_path, mode):
    """FileSource at scale (ciphertext_source.rs:35-106): gc_{i}.bin images (np.memmap) are fed through a small
    device ring WHILE the evaluate kernel runs and hashed on host threads; same outputs and chain hash as the
    resident path, and a truncated file is reported as "Ciphertext source exhausted"."""
    p, st = circuit("fq12_mul")
    B = 3 if mode == 1 else 33
    seeds = list(range(50, 50 + B))
    g = gsv.Session(p, B, ct_mode=gsv.CT_KEEP, exec_mode=mode, group=1 if mode == 1 else 0)
    res = g.garble(seeds, gsv.HASH_AES)
    files = []
    for i in range(B):
        path = tmp_path / f"gc_{i}.bin"
        g.read_ciphertexts(i).tofile(path)
        files.append(np.memmap(path, dtype=np.uint8, mode="r"))
    g.close()
    rng = np.random.default_rng(13)
    bits = rng.integers(0, 2, (B, p.n_inputs), dtype=np.uint8)
    active = res.input_label0 ^ (res.delta[:, None, :] * bits[:, :, None])
    e = gsv.Session(p, B, ct_mode=gsv.CT_NONE, exec_mode=mode, group=1 if mode == 1 else 0)
    ev = e.evaluate(gsv.HASH_AES, res.true_label1, res.false_label0, active, bits, ct_streams=files, ct_ring_log2=18)
    assert np.array_equal(ev.ct_commit, res.ct_commit)
    for i in (0, B - 1):
        want = st.execute(bits[i])
        assert np.array_equal(ev.output_bits[i], want)
        assert np.array_equal(ev.output_active[i], res.output_label0[i] ^ (res.delta[i][None, :] * want[:, None]))
    short = [f[: 16 * (p.n_ciphertexts - 5)] for f in files]
    with pytest.raises(gsv.GsvError) as ex:
        e.evaluate(gsv.HASH_AES, res.true_label1, res.false_label0, active, bits, ct_streams=short, ct_ring_log2=18)
    assert ex.value.code == -5


@pytest.mark.parametrize("name,B,G,window", [("fq_mul", 4, 2, 16), ("fq12_mul", 8, 4, 64), ("fq12_mul", 3, 1, 1),
                                             ("fq_inverse", 4, 4, 64)])
def test_call_pipelining_matches_the_plain_plan(gsv, name, B, G, window):
    """Call pipelining (gsv_plan_options.pipeline: calls queued when their producers START, inputs gathered
    window by window behind per-slot ready flags, outputs published early) changes the schedule only: labels,
    the whole ciphertext stream, its chain commitment, the evaluation and ExecuteMode are those of the plain plan."""
    seeds = list(range(900, 900 + B))
    plain = gsv.Program(name, pipeline=False)
    piped = gsv.Program(name, pipeline=True, window_levels=window)
    assert piped.n_calls == plain.n_calls and piped.n_ciphertexts == plain.n_ciphertexts
    assert piped.critical_path_levels <= plain.critical_path_levels
    out = []
    for p in (plain, piped):
        s = gsv.Session(p, B, group=G, ct_mode=gsv.CT_KEEP, exec_mode=1)
        for _ in range(2):  # twice: the ready flags of the first call are stale in the second (epochs)
            r = s.garble(seeds, gsv.HASH_AES)
        out.append((s, r))
    (s0, r0), (s1, r1) = out
    for f in ("delta", "input_label0", "output_label0", "ct_commit"):
        assert np.array_equal(getattr(r0, f), getattr(r1, f)), f
    assert np.array_equal(s0.read_ciphertexts(B - 1), s1.read_ciphertexts(B - 1))
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2, (B, piped.n_inputs), dtype=np.uint8)
    ev = s1.evaluate(gsv.HASH_AES, r1.true_label1, r1.false_label0, _eval_inputs(r1, bits), bits)
    want = np.stack([piped.execute(b) for b in bits])
    assert np.array_equal(ev.output_bits, want)
    assert np.array_equal(ev.ct_commit, r0.ct_commit)
    sel = r0.output_label0.copy()
    sel[want.astype(bool)] ^= np.broadcast_to(r0.delta[:, None, :], sel.shape)[want.astype(bool)]
    assert np.array_equal(ev.output_active, sel)
    got, _ = s1.execute(bits)
    assert np.array_equal(got, want)

"""Full Groth16 verifier (groth16_verify_compressed, 11.46 G gates) on the GPU: garble -> evaluate.

usage: verifier_e2e.py MODE B        MODE = levelised | lane | lane_throughput
  levelised / lane : garble B instances keeping the raw ciphertext stream (47.7 GB per instance, so
                     B <= 2 on one B200), evaluate a valid and a tampered proof against it, and check
                     verify bit, gw.select(value) == active label, and (same seeds) reproducibility.
  lane_throughput  : garble B instances, ciphertexts dropped (`()` handler): gates/s of the verifier.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsv_b200 as g

mode, B = sys.argv[1], int(sys.argv[2])
t0 = time.time()
prog = g.Program("groth16_verify_compressed", lane_only=(mode != "levelised"))
rec = {"circuit": "groth16_verify_compressed", "mode": mode, "instances": B, "gates": prog.n_gates,
       "ciphertexts": prog.n_ciphertexts, "calls": prog.n_calls, "tasks": prog.n_tasks,
       "global_slots": prog.n_global_slots, "plan_s": round(time.time() - t0, 1)}
print(rec, flush=True)
seeds = [1234 + i for i in range(B)]
if mode == "lane_throughput":
    s = g.Session(prog, B, ct_mode=g.CT_NONE, exec_mode=2)
    r = s.garble(seeds, g.HASH_AES, want_inputs=False)
    rec.update(garble_ms=r.ms_garble, gates_per_s=prog.n_gates * B / (r.ms_garble * 1e-3))
else:
    s = g.Session(prog, B, ct_mode=g.CT_KEEP_RAW, exec_mode=1 if mode == "levelised" else 2,
                  group=(2 if B % 2 == 0 else 1) if mode == "levelised" else 0)
    r = s.garble(seeds, g.HASH_AES)
    rec.update(garble_ms=r.ms_garble, gates_per_s=prog.n_gates * B / (r.ms_garble * 1e-3))
    print(rec, flush=True)
    for flip in (False, True):
        bits = np.tile(g.groth16_synthetic_inputs(424242, flip), (B, 1))
        act = r.input_label0.copy()
        m = bits.astype(bool)
        act[m] ^= np.broadcast_to(r.delta[:, None, :], act.shape)[m]
        ev = s.evaluate(g.HASH_AES, r.true_label1, r.false_label0, act, bits, want_commit=False)
        want_bit = 0 if flip else 1
        sel = r.output_label0[:, 0, :] ^ (r.delta * want_bit)
        ok = bool(np.all(ev.output_bits[:, 0] == want_bit) and np.array_equal(ev.output_active[:, 0, :], sel))
        rec["evaluate_%s" % ("tampered" if flip else "valid")] = {"verify_bit": ev.output_bits[:, 0].tolist(),
                                                                  "label_matches_garbler": ok, "ms": ev.ms_evaluate}
        assert ok, rec
print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/verifier_%s_B%d.json" % (mode, B), "w").write(json.dumps(rec))

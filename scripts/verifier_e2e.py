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
NT = int(sys.argv[3]) if len(sys.argv) > 3 else 0
GRP = int(sys.argv[4]) if len(sys.argv) > 4 else (2 if B % 2 == 0 else 1)
t0 = time.time()
prog = g.Program("groth16_verify_compressed", lane_only=mode.startswith("lane"))
rec = {"circuit": "groth16_verify_compressed", "mode": mode, "instances": B, "worker_threads": NT, "group": GRP, "gates": prog.n_gates,
       "ciphertexts": prog.n_ciphertexts, "calls": prog.n_calls, "tasks": prog.n_tasks,
       "global_slots": prog.n_global_slots, "plan_s": round(time.time() - t0, 1)}
print(rec, flush=True)
seeds = [1234 + i for i in range(B)]
if mode == "levelised_none":
    s = g.Session(prog, B, ct_mode=g.CT_NONE, exec_mode=1, group=GRP, worker_threads=NT)
    r = s.garble(seeds, g.HASH_AES, want_inputs=False)
    rec.update(garble_ms=r.ms_garble, gates_per_s=prog.n_gates * B / (r.ms_garble * 1e-3))
elif mode == "host_commit":
    # config 2/4 of BASELINE.json: garbling with the ciphertext commitment only, every gate hash on the
    # GPU, the serial chain folded by host AES-NI threads draining the ring (GSV_CT_COMMIT_HOST)
    s = g.Session(prog, B, ct_mode=g.CT_COMMIT_HOST, exec_mode=1, group=GRP, worker_threads=NT)
    for it in range(int(os.environ.get("REPS", "1"))):
        t1 = time.time()
        r = s.garble(seeds, g.HASH_AES, want_inputs=False)
        wall = time.time() - t1
        rec.setdefault("runs", []).append({"kernel_ms": r.ms_garble, "wall_s": round(wall, 2),
                                           "gates_per_s": prog.n_gates * B / wall})
        print(rec["runs"][-1], flush=True)
    rec["commit_seed1234"] = bytes(r.ct_commit[0]).hex()
    rec["commit_seed1235"] = bytes(r.ct_commit[1]).hex() if B > 1 else None
    rec["output_label0_seed1234"] = bytes(r.output_label0[0, 0]).hex()
elif mode == "lane_throughput":
    s = g.Session(prog, B, ct_mode=g.CT_NONE, exec_mode=2)
    r = s.garble(seeds, g.HASH_AES, want_inputs=False)
    rec.update(garble_ms=r.ms_garble, gates_per_s=prog.n_gates * B / (r.ms_garble * 1e-3))
else:
    s = g.Session(prog, B, ct_mode=g.CT_KEEP_RAW, exec_mode=1 if mode == "levelised" else 2,
                  group=GRP if mode == "levelised" else 0, worker_threads=NT)
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
open("gpurun_out/verifier_%s_B%d_nt%d_g%d.json" % (mode, B, NT, GRP), "w").write(json.dumps(rec))

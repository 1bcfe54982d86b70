"""The reference's cut-and-choose flow (examples/groth16_cut_and_choose.rs, src/cut_and_choose/{garbler,evaluator}.rs)
at FULL verifier scale on the GPU(s): total 4 instances / 2 finalized (the example's defaults).

  1. pre-check: ExecuteMode on the GPU -- the synthetic proof verifies (IS_PRE_BOOLEAN_EXEC)
  2. Garbler::create + commit          : 4 instances garbled in one batched GSV_CT_COMMIT_HOST call
  3. Evaluator::create                 : picks 2 instances to finalize
  4. Garbler::open_commit / Evaluator::run_regarbling (opened): the 2 opened seeds are re-garbled in one call
     and their commit records compared
  5. finalized instances               : re-garbled on GPU 0 and STREAMED to the evaluator session (GPU 1 when
     there are two GPUs: peer stores over NVLink) instead of being written to 47.7 GB gc_{i}.bin files; the
     evaluator hashes what it receives (== committed chain hash) and evaluates at the same time
  6. evaluate_from checks              : constant / input / output label commits, verify bit

usage: verifier_cut_and_choose.py [circuit]     (default groth16_verify_compressed; fq12_mul for a quick run)
"""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import gsv_b200 as g

cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
circuit = sys.argv[1] if len(sys.argv) > 1 else "groth16_verify_compressed"
TOTAL, FIN = 4, 2
rec = {"circuit": circuit, "total": TOTAL, "to_finalize": FIN}
t0 = time.perf_counter()
prog = g.Program(circuit)
rec.update(plan_s=round(time.perf_counter() - t0, 1), gates=prog.n_gates, ciphertexts=prog.n_ciphertexts)
two = g.device_count() >= 2
if circuit.startswith("groth16"):
    bits = g.groth16_synthetic_inputs(compressed=circuit.endswith("compressed"))
else:
    bits = np.random.default_rng(1).integers(0, 2, prog.n_inputs, dtype=np.uint8)

# 1. pre-check in ExecuteMode on the GPU
t = time.perf_counter()
pre = g.Session(prog, 1, ct_mode=g.CT_NONE, exec_mode=1, group=1)
out, ms = pre.execute(bits[None, :])
pre.close()
rec["precheck"] = {"verify_bit": int(out[0, 0]), "kernel_ms": round(ms, 1), "wall_s": round(time.perf_counter() - t, 2)}
print(rec["precheck"], flush=True)

# 2. create + commit
t = time.perf_counter()
garbler = cc.Garbler(prog, TOTAL, master_seed=1234, ct_mode=g.CT_COMMIT_HOST, exec_mode=1, group=4)
garbler.create()
commits = garbler.commit()
rec["create_commit_s"] = round(time.perf_counter() - t, 2)
print("create+commit", rec["create_commit_s"], "s", flush=True)

# 3. the evaluator picks
ev = cc.Evaluator(prog, TOTAL, FIN, rng_seed=99, commits=commits)
rec["finalized"] = ev.to_finalize

# 4. opened instances: seeds revealed, re-garbled by the evaluator (one batched call), commit records compared
t = time.perf_counter()
opened = [i for i in range(TOTAL) if i not in ev.to_finalize]
check = cc.Garbler(prog, len(opened), 0, hasher=0, ct_mode=g.CT_COMMIT_HOST, exec_mode=1, group=2,
                   seeds=np.array([int(garbler.seeds_all[i]) for i in opened], dtype=np.uint64))
rec_open = check.commit().records
check.session.close()
for k, i in enumerate(opened):
    assert np.array_equal(rec_open[k], commits.records[i]), f"RegarblingMismatch {i}"
rec["regarble_opened_s"] = round(time.perf_counter() - t, 2)
print("opened instances re-garbled and equal", rec["regarble_opened_s"], "s", flush=True)

# 5./6. finalized instances: streamed garbler -> evaluator, hashed and evaluated on the way
t = time.perf_counter()
garbler.session.close()
sm = 0 if two else 70
gs = g.Session(prog, FIN, device=0, ct_mode=g.CT_NONE, exec_mode=1, group=2, sm_limit=sm)
es = g.Session(prog, FIN, device=1 if two else 0, ct_mode=g.CT_NONE, exec_mode=1, group=2, sm_limit=sm)
g.link_sessions(gs, es)
seeds_fin = [int(garbler.seeds_all[i]) for i in ev.to_finalize]
gres, res = g.stream_garble_evaluate(gs, es, seeds_fin, g.HASH_AES, np.tile(bits, (FIN, 1)))
rec["stream_finalized_s"] = round(time.perf_counter() - t, 2)
for k, i in enumerate(ev.to_finalize):
    assert np.array_equal(res.ct_commit[k], commits.ct_commit()[i]), f"CiphertextMismatch {i}"
    oc = g.commit_labels(res.output_active[k])
    exp = commits.output_commits()[i]                      # [n_out, 2, 16]: (c(label1), c(label0))
    sel = exp[np.arange(prog.n_outputs), 1 - res.output_bits[k].astype(np.int64)]
    assert np.array_equal(sel, oc), f"OutputLabelMismatch {i}"
rec["verify_bits"] = res.output_bits[:, 0].tolist()
rec["nvlink" if two else "same_gpu"] = True
rec["stream_GBps"] = round(prog.n_ciphertexts * 16 * FIN / rec["stream_finalized_s"] / 1e9, 2)
print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/cut_and_choose_{circuit}.json", "w").write(json.dumps(rec))

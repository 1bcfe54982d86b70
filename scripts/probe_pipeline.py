"""Call pipelining probe: the same circuit planned without and with pipelining (PlanOptions.pipeline), garbled with the
same seeds -- results must be identical (labels and chain commitments), times are compared.

usage: probe_pipeline.py [circuits=fq_mul,fq12_mul,fq12_inverse,g1_msm1] [B=16] [G=4] [W=64,16] [ct_modes=0,1] [reps=3] [execute] [noref]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import gsv_b200 as g

kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
flags = set(a for a in sys.argv[1:] if "=" not in a)
B, G = int(kv.get("B", 16)), int(kv.get("G", 4))
rows = []
for circ in kv.get("circuits", "fq_mul,fq12_mul,fq12_inverse,g1_msm1").split(","):
    ref = None
    for W in ([] if "noref" in flags else [0]) + [int(w) for w in kv.get("W", "64,16").split(",")]:
        prog = g.Program(circ, pipeline=bool(W), window_levels=W)
        for ct_mode in [int(m) for m in kv.get("ct_modes", "0,1").split(",")]:
            s = g.Session(prog, B, group=G, ct_mode=ct_mode, exec_mode=1)
            best = None
            for _ in range(int(kv.get("reps", 3))):
                r = s.garble(list(range(100, 100 + B)), g.HASH_AES)
                best = r if best is None or r.ms_garble < best.ms_garble else best
            key = (circ, ct_mode)
            got = (best.output_label0.tobytes(), None if ct_mode == g.CT_NONE else best.ct_commit.tobytes())
            if W == 0:
                ref = ref or {}
                ref[key] = got
            same = ref is None or ref[key] == got
            row = dict(circuit=circ, window=W, ct_mode=ct_mode, ms=round(best.ms_garble, 3), same=same,
                       crit=prog.critical_path_levels, calls=prog.n_calls,
                       us_per_level=round(1e3 * best.ms_garble / max(prog.critical_path_levels, 1), 3),
                       ggps=round(prog.n_gates * B / best.ms_garble / 1e6, 3))
            if "execute" in flags and ct_mode == g.CT_NONE:
                # ExecuteMode on the same (pipelined) plan against the host walk of the recorded topology
                bits = np.random.default_rng(5).integers(0, 2, (8, prog.n_inputs), dtype=np.uint8)
                out, _ = s.execute(bits)
                row["exec_ok"] = bool(np.array_equal(out, np.stack([prog.execute(b) for b in bits])))
            rows.append(row)
            print(json.dumps(row), flush=True)
            s.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/probe_pipeline.json", "w"), indent=1)
bad = [r for r in rows if not r["same"] or r.get("exec_ok") is False]
print("MISMATCH" if bad else "all equal", flush=True)
sys.exit(1 if bad else 0)

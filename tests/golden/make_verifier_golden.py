"""Generates tests/golden/verifier_vectors.json: the CPU oracle's garbling of the FULL Groth16 verifier
(`groth16_verify_compressed`, 11.46 G gates) for the given seeds, walked over the exported template DAG
(oracle/gsv_oracle.c: gsvo_garble_templates).  About 6 minutes of one core per seed.

    python tests/golden/make_verifier_golden.py [seed ...]      # default: 1234 1235
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge
    ge.build()
    import gsv_b200 as g
    from oracle import oracle as o

    seeds = [int(x) for x in sys.argv[1:]] or [1234, 1235]
    p = g.Program("groth16_verify_compressed", lane_only=True)
    dag = o.TemplateDag(*p.export_templates())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "verifier_vectors.json")
    out = {"_doc": "CPU-oracle garbling of groth16_verify_compressed (synthetic vk of gsv_groth16_synthetic_inputs), "
                   "AES hasher; labels are 16 big-endian bytes", "circuit": "groth16_verify_compressed",
           "n_gates": p.n_gates, "n_ciphertexts": p.n_ciphertexts, "vectors": []}
    for seed in seeds:
        t0 = time.time()
        r = dag.garble(o.HASH_AES, seed)
        assert r["n_gates"] == p.n_gates and r["n_ct"] == p.n_ciphertexts
        out["vectors"].append({"seed": seed, "delta": r["delta"].hex(), "false_label0": r["false_label0"].hex(),
                               "true_label0": r["true_label0"].hex(),
                               "input_label0_first": bytes(r["input_label0"][0]).hex(),
                               "output_label0": bytes(r["output_label0"][0]).hex(),
                               "ct_commit": r["ct_commit"].hex(), "oracle_seconds": round(time.time() - t0, 1)})
        print(out["vectors"][-1], flush=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

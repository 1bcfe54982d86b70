"""Writes tests/golden/structural_hashes.json: the structural hash (structural_hash.py) of the named circuits as
produced by the INDEPENDENT model (emission_model.py gadgets recorded by structural_hash.SCtx) -- up to the whole
Groth16 verifier.  tests/test_structural_hash.py holds the product's recorder (Program.export_templates()) to them.

usage: python tests/golden/make_structural_hashes.py        (about 8 minutes, 3 GB)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import structural_hash as sh  # noqa: E402
import structural_roots as sr  # noqa: E402

out = {}
for name, (n, fn) in sr.ROOTS.items():
    t = time.time()
    out[name] = {"n_inputs": n, "structural_sha256": sh.model_hash(fn, n)}
    print(name, out[name], round(time.time() - t, 1), "s", flush=True)
with open(os.path.join(HERE, "structural_hashes.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_structural_hashes.py (emission_model.py gadgets over structural_hash.SCtx)",
               "definition": "structural_hash.py: sha256 over the component DAG, bottom-up, per (body, output-liveness mask)",
               "synthetic_key": {
                   "note": "key of the groth16_* / miller_loop_groth16 circuits: points = scalar * generator (G1 (1, 2); G2 the "
                           "EIP-197 generator); alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1 = [ic0, ic1]; "
                           "csrc/bn254_host.cpp synthetic_groth16(7, ..)",
                   "scalars": {k: hex(v) for k, v in sr.em.synthetic_vk_scalars(7).items()}},
               "circuits": out}, f, indent=1)

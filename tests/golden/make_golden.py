"""Generates tests/golden/garble_vectors.json.

The reference (Rust) cannot run in the build image and ships no golden vectors for this path, so the
fixture is produced by the CPU oracle AFTER it has been pinned on FIPS-197 / OpenSSL / BLAKE3 / ChaCha20
and the SURVEY Appendix-E table (tests/test_oracle_primitives.py).  It freezes today's bytes: the CPU suite
checks the oracle against it (an oracle regression cannot silently move the target) and the GPU suite
checks the CUDA path against it WITHOUT loading the oracle.

    python tests/golden/make_golden.py        # rewrites the fixture
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CIRCUITS = ["gate_zoo", "fq_add", "fq_mul", "fq_expr", "fq2_mul", "fq12_mul", "fq_inverse", "g1_add"]
SEEDS = [0, 42, 1234]   # 0: tests/fq12_mul_e2e.rs, 42 / 1234: the reference examples' seeds
HASHERS = {"aes": 0, "blake3": 1}


def sha(x):
    return hashlib.sha256(bytes(x)).hexdigest()


def main():
    import __graft_entry__ as ge
    ge.build()
    import gsv_b200 as g
    from oracle import oracle as o

    out = {"_doc": "CPU-oracle outputs per (circuit, hasher, seed); labels are 16 big-endian bytes (S::to_bytes), "
                   "*_sha256 are digests of the concatenated bytes", "vectors": []}
    for name in CIRCUITS:
        p = g.Program(name)
        t, a, b, c, outs, nw = p.flat_stream()
        st = o.Stream(t, a, b, c, outs, nw, p.n_inputs)
        for hname, h in HASHERS.items():
            if hname == "blake3" and p.n_gates > 5_000_000:
                continue  # keep generation short: the big circuits are pinned with the AES hasher
            for seed in SEEDS:
                r = st.garble(h, seed)
                out["vectors"].append({
                    "circuit": name, "hasher": hname, "seed": seed,
                    "n_gates": p.n_gates, "n_ciphertexts": int(r["cts"].shape[0]),
                    "delta": r["delta"].hex(), "false_label0": r["false_label0"].hex(),
                    "true_label0": r["true_label0"].hex(),
                    "input_label0_first": bytes(r["input_label0"][0]).hex(),
                    "input_label0_sha256": sha(r["input_label0"].tobytes()),
                    "output_label0_first": bytes(r["output_label0"][0]).hex(),
                    "output_label0_sha256": sha(r["output_label0"].tobytes()),
                    "ct_first": bytes(r["cts"][0]).hex() if len(r["cts"]) else None,
                    "ct_stream_sha256": sha(r["cts"].tobytes()),
                    "ct_commit": r["ct_commit"].hex(),
                })
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "garble_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(out['vectors'])} vectors to {path}")


if __name__ == "__main__":
    main()

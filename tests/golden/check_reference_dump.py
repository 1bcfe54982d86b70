"""Compares the JSON lines printed by gsv-cuda/tests/golden_dump.rs (run against the REAL reference crate, wherever
a Rust toolchain exists) with this repository's committed fixtures: the one step that lifts "parity unpinned" for
the seed -> label derivation (row a2) and the Fq12-mul gate stream (row a8).

usage: cargo test --release golden_dump -- --nocapture | grep '^{' > dump.jsonl
       python tests/golden/check_reference_dump.py dump.jsonl
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main(path):
    with open(os.path.join(HERE, "garble_vectors.json")) as f:
        vec = {(v["circuit"], v["hasher"], v["seed"]): v for v in json.load(f)["vectors"]}
    import importlib

    cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
    bad = 0
    for ln in open(path):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        if d["kind"] == "garble":
            want = vec[(d["circuit"], d["hasher"], d["seed"])]
            pairs = [("ct_commit", "ct_commit"), ("false_label0", "false_label0"), ("true_label0", "true_label0"),
                     ("first_input_label0", "input_label0_first"), ("first_output_label0", "output_label0_first"),
                     ("n_gates", "n_gates")]
            for a, b in pairs:
                ok = d[a] == want[b]
                bad += not ok
                print(("ok  " if ok else "FAIL"), d["circuit"], d["hasher"], d["seed"], a, d[a], "" if ok else f"!= {want[b]}")
        elif d["kind"] == "rng_u128":
            want = vec[("fq_add", "aes", d["seed"])] if ("fq_add", "aes", d["seed"]) in vec else None
            if want:   # draw order: delta, false.label0, true.label0, first input label (garble_mode.rs:80-97)
                got = d["draws"]
                exp = [want["delta"], want["false_label0"], want["true_label0"], want["input_label0_first"]]
                ok = got == exp
                bad += not ok
                print(("ok  " if ok else "FAIL"), "rng_u128 seed", d["seed"], got, "" if ok else f"!= {exp}")
        elif d["kind"] == "instance_seeds":
            exp = [int(x) for x in cc.instance_seeds(d["master"], len(d["seeds"]))]
            ok = d["seeds"] == exp
            bad += not ok
            print(("ok  " if ok else "FAIL"), "instance_seeds", d["master"], "" if ok else f"{d['seeds']} != {exp}")
        elif d["kind"] == "hash":
            exp = {"aes": "8a7289ea9b51aa8cdcbd087a643871fc", "blake3": "2012369da457cf398b10b1aef058df9e"}[d["hasher"]]
            ok = d["value"] == exp
            bad += not ok
            print(("ok  " if ok else "FAIL"), "hash", d["hasher"], d["value"], "" if ok else f"!= {exp}")
    print("ALL EQUAL" if bad == 0 else f"{bad} MISMATCHES")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))

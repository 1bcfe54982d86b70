"""Writes tests/golden/stream_hashes.json: SHA-256 of the canonical (type, a, b, c) gate stream of the named
circuits as produced by the INDEPENDENT emission model (emission_model.py, written from the reference's Rust
gadgets and SURVEY.md Appendix B -- not from the product's generator).  tests/test_emission_order.py holds the
product's generator (`Program.flat_stream()`) to these hashes.

usage: python tests/golden/make_stream_hashes.py [--slow]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import emission_model as em  # noqa: E402

CIRCUITS = ["fq_add", "bn_mul4", "bn_mul19", "bn_mul21", "bn_mul64", "bn_mul254", "fq_mul", "fq2_mul", "fq6_mul", "fq12_mul",
            "fq_inverse", "g1_add", "fq12_square", "fq12_cyclotomic_square", "fq12_frobenius1", "fq12_frobenius2",
            "fq12_frobenius3", "fq12_inverse", "g2_double_step", "g2_add_step", "g2_mul_by_char", "ell", "ell_const",
            "g1_to_affine"]

# fq_sqrt (Fq::sqrt_montgomery = a^((p + 1) / 4), 148.7 M gates) and decompress_g1 (149.6 M) take 70-90 s and 17 GB
# each here: only with --slow
SLOW = ["fq_sqrt", "decompress_g1"]
if "--slow" in sys.argv:
    CIRCUITS += SLOW
else:
    with open(os.path.join(HERE, "stream_hashes.json")) as f:
        KEEP = {k: v for k, v in json.load(f)["circuits"].items() if k in SLOW}

out = {}
for c in CIRCUITS:
    h, info = em.canonical_hash(*em.build(c))
    out[c] = dict(sha256=h, **info)
    print(c, h, info)
if "--slow" not in sys.argv:
    out.update(KEEP)
with open(os.path.join(HERE, "stream_hashes.json"), "w") as f:
    json.dump({"generator": "tests/golden/emission_model.py", "canonical": "wires renumbered by first live write; "
               "sha256(n_inputs:i64 | type:u8[] | a:u32[] | b:u32[] | c:u32[] (0xFFFFFFFF = dead) | outputs:u32[])",
               "circuits": out}, f, indent=1)

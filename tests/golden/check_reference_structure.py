"""Compares template dumps written by gsv-cuda/tests/structure_dump.rs -- the REAL reference gadgets recorded through
the C ABI (GpuRecorder: CircuitContext) wherever a Rust toolchain exists -- with tests/golden/structural_hashes.json,
the hashes of the independent model.  Together with tests/test_structural_hash.py (product's generator == model) this
closes row a8 against the reference itself for every key-independent circuit up to the final exponentiation.

Dump format (little endian): magic "GSVT", u32 root, 6 x u64 sizes (words), then the six u32 arrays
(tmpl, gates, calls, items, call_wires, outs) of gsv_program_export_templates.

usage: cargo test --release structure_dump -- --nocapture        (writes target/structure/<circuit>.gsvt)
       python tests/golden/check_reference_structure.py gsv-cuda/target/structure
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import structural_hash as sh  # noqa: E402


def read_dump(path):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == b"GSVT", path
    root, = struct.unpack_from("<I", raw, 4)
    sizes = struct.unpack_from("<6Q", raw, 8)
    off, arrays = 56, []
    for n in sizes:
        arrays.append(np.frombuffer(raw, np.uint32, n, off))
        off += 4 * n
    return (root, *arrays)


def main(directory):
    with open(os.path.join(HERE, "structural_hashes.json")) as f:
        golden = json.load(f)["circuits"]
    bad = 0
    for fn in sorted(os.listdir(directory)):
        if not fn.endswith(".gsvt"):
            continue
        name = fn[:-5]
        got = sh.templates_hash(*read_dump(os.path.join(directory, fn)))
        ok = name in golden and got == golden[name]["structural_sha256"]
        bad += not ok
        print("ok  " if ok else "FAIL", name, got)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))

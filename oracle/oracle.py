"""ctypes wrapper of the CPU oracle (oracle/gsv_oracle.c).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsv_oracle.so")

HASH_AES, HASH_BLAKE3 = 0, 1
WIRE_DEAD = 0xFFFFFFFF


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB_PATH


class _Stream(C.Structure):
    _fields_ = [
        ("n_gates", C.c_uint64), ("type", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p),
        ("c", C.c_void_p), ("n_wires", C.c_uint32), ("n_inputs", C.c_uint32),
        ("n_outputs", C.c_uint32), ("outputs", C.c_void_p),
    ]


class _Summary(C.Structure):
    _fields_ = [
        ("delta", C.c_uint8 * 16), ("false_label0", C.c_uint8 * 16), ("true_label0", C.c_uint8 * 16),
        ("ct_commit", C.c_uint8 * 16), ("n_ct", C.c_uint64), ("n_gates", C.c_uint64),
    ]


class _Rng(C.Structure):
    _fields_ = [("key", C.c_uint32 * 8), ("block", C.c_uint64), ("buf", C.c_uint32 * 16), ("pos", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.gsvo_tweak.argtypes = [C.c_uint64, C.c_void_p]
        L.gsvo_hash_aes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.gsvo_hash_blake3.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.gsvo_blake3_small.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.gsvo_garble_gate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.gsvo_degarble_gate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.gsvo_seed_key.argtypes = [C.c_uint64, C.c_void_p]
        L.gsvo_rng_init.argtypes = [C.c_void_p, C.c_uint64]
        L.gsvo_rng_u64.restype = C.c_uint64
        L.gsvo_chacha20_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.gsvo_garble_stream.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.gsvo_evaluate_stream.argtypes = [C.c_int, C.c_void_p] + [C.c_void_p] * 5 + [C.c_uint64] + [C.c_void_p] * 4
        L.gsvo_execute_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gsvo_compact_stream.argtypes = [C.c_void_p] * 6
        L.gsvo_garble_templates.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gsvo_garble_templates_prefix.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _b16(x) -> bytes:
    b = bytes(x)
    assert len(b) == 16
    return b


def aes128(key: bytes, block: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().gsvo_aes128_encrypt(_b16(key), _b16(block), out)
    return bytes(out)


def aes_fixed(block: bytes, impl: str = "auto") -> bytes:
    out = (C.c_uint8 * 16)()
    if impl == "portable":
        lib().gsvo_aes_fixed_portable(_b16(block), out)
    elif impl == "aesni":
        if lib().gsvo_aes_fixed_aesni(_b16(block), out) != 0:
            raise RuntimeError("no AES-NI")
    else:
        lib().gsvo_aes_fixed(_b16(block), out)
    return bytes(out)


def have_aesni() -> bool:
    return bool(lib().gsvo_have_aesni())


def tweak(gid: int) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().gsvo_tweak(gid, out)
    return bytes(out)


def hash_gate(hasher: int, x: bytes, gid: int) -> bytes:
    out = (C.c_uint8 * 16)()
    (lib().gsvo_hash_aes if hasher == HASH_AES else lib().gsvo_hash_blake3)(_b16(x), gid, out)
    return bytes(out)


def blake3_small(msg: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    assert lib().gsvo_blake3_small(msg, len(msg), out) == 0
    return bytes(out)


def garble_gate(hasher: int, gate_type: int, a0: bytes, b0: bytes, delta: bytes, gid: int):
    c0 = (C.c_uint8 * 16)()
    ct = (C.c_uint8 * 16)()
    has = lib().gsvo_garble_gate(hasher, gate_type, _b16(a0), _b16(b0), _b16(delta), gid, c0, ct)
    return bytes(c0), (bytes(ct) if has else None)


def degarble_gate(hasher: int, gate_type: int, ct, a_act: bytes, a_val: int, b_act: bytes, gid: int) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().gsvo_degarble_gate(hasher, gate_type, ct, _b16(a_act), int(a_val), _b16(b_act), gid, out)
    return bytes(out)


def gate_eval(gate_type: int, a: int, b: int) -> int:
    return lib().gsvo_gate_eval(gate_type, a, b)


def chain(cts) -> bytes:
    h = (C.c_uint8 * 16)()
    for ct in cts:
        lib().gsvo_chain_update(h, _b16(ct))
    return bytes(h)


def commit_label(label: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().gsvo_commit_label(_b16(label), out)
    return bytes(out)


def seed_key(seed: int) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().gsvo_seed_key(seed, out)
    return bytes(out)


def chacha20_block(key_words, counter: int):
    k = (C.c_uint32 * 8)(*key_words)
    out = (C.c_uint32 * 16)()
    lib().gsvo_chacha20_block(k, counter, out)
    return list(out)


class Rng:
    """ChaCha20Rng::seed_from_u64(seed)."""

    def __init__(self, seed: int):
        self._r = _Rng()
        lib().gsvo_rng_init(C.byref(self._r), seed)

    def label(self) -> bytes:
        out = (C.c_uint8 * 16)()
        lib().gsvo_rng_label(C.byref(self._r), out)
        return bytes(out)

    def u64(self) -> int:
        return lib().gsvo_rng_u64(C.byref(self._r))


class Stream:
    """A flat emission-order gate stream (arrays are kept alive by this object)."""

    def __init__(self, type_, a, b, c, outputs, n_wires: int, n_inputs: int):
        self.type = np.ascontiguousarray(type_, np.uint8)
        self.a = np.ascontiguousarray(a, np.uint32)
        self.b = np.ascontiguousarray(b, np.uint32)
        self.c = np.ascontiguousarray(c, np.uint32)
        self.outputs = np.ascontiguousarray(outputs, np.uint32)
        self.n_wires, self.n_inputs = int(n_wires), int(n_inputs)
        self.n_gates = int(self.type.shape[0])
        self.n_outputs = int(self.outputs.shape[0])
        self._s = _Stream(self.n_gates, self.type.ctypes.data, self.a.ctypes.data, self.b.ctypes.data,
                          self.c.ctypes.data, self.n_wires, self.n_inputs, self.n_outputs,
                          self.outputs.ctypes.data)

    def compact(self) -> "Stream":
        """Same stream with recycled slot ids (cache-resident live set, like the reference's slab)."""
        a2, b2, c2 = (np.zeros(self.n_gates, np.uint32) for _ in range(3))
        o2 = np.zeros(max(self.n_outputs, 1), np.uint32)
        ns = C.c_uint32(0)
        rc = lib().gsvo_compact_stream(C.byref(self._s), a2.ctypes.data, b2.ctypes.data, c2.ctypes.data,
                                       o2.ctypes.data, C.byref(ns))
        assert rc == 0
        return Stream(self.type, a2, b2, c2, o2[: self.n_outputs], ns.value, self.n_inputs)

    def garble(self, hasher: int, seed: int, want_ct: bool = True):
        n_nonfree = int(((self.type < 8) & (self.c != WIRE_DEAD)).sum())
        inl = np.zeros((self.n_inputs, 16), np.uint8)
        outl = np.zeros((self.n_outputs, 16), np.uint8)
        cts = np.zeros((n_nonfree if want_ct else 0, 16), np.uint8)
        sm = _Summary()
        rc = lib().gsvo_garble_stream(hasher, seed, C.byref(self._s), inl.ctypes.data, outl.ctypes.data,
                                      cts.ctypes.data if want_ct else None, cts.shape[0], C.byref(sm))
        if rc != 0:
            raise RuntimeError(f"oracle garble failed: {rc}")
        return {
            "delta": bytes(sm.delta), "false_label0": bytes(sm.false_label0),
            "true_label0": bytes(sm.true_label0), "ct_commit": bytes(sm.ct_commit),
            "n_ct": int(sm.n_ct), "input_label0": inl, "output_label0": outl, "cts": cts,
        }

    def evaluate(self, hasher: int, true_label: bytes, false_label: bytes, input_active, input_bits, cts):
        ia = np.ascontiguousarray(input_active, np.uint8).reshape(self.n_inputs, 16)
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(self.n_inputs)
        ct = np.ascontiguousarray(cts, np.uint8).reshape(-1, 16)
        oa = np.zeros((self.n_outputs, 16), np.uint8)
        ob = np.zeros(self.n_outputs, np.uint8)
        cc = (C.c_uint8 * 16)()
        used = C.c_uint64(0)
        rc = lib().gsvo_evaluate_stream(hasher, C.byref(self._s), _b16(true_label), _b16(false_label),
                                        ia.ctypes.data, ib.ctypes.data, ct.ctypes.data, ct.shape[0],
                                        oa.ctypes.data, ob.ctypes.data, cc, C.byref(used))
        return {"rc": rc, "output_active": oa, "output_bits": ob, "ct_commit": bytes(cc), "n_ct_used": used.value}

    def execute(self, input_bits):
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(self.n_inputs)
        ob = np.zeros(self.n_outputs, np.uint8)
        rc = lib().gsvo_execute_stream(C.byref(self._s), ib.ctypes.data, ob.ctypes.data)
        assert rc == 0
        return ob


class _Templates(C.Structure):
    _fields_ = [("n_templates", C.c_uint32), ("root", C.c_uint32), ("tmpl", C.c_void_p), ("gates", C.c_void_p),
                ("calls", C.c_void_p), ("items", C.c_void_p), ("call_wires", C.c_void_p), ("outs", C.c_void_p)]


class TemplateDag:
    """A circuit as the product's exported template DAG (`Program.export_templates()`): the oracle walks
    it depth-first in emission order, so the 11 G-gate verifier garbles on the CPU in a few MB."""

    def __init__(self, root: int, tmpl, gates, calls, items, call_wires, outs):
        self.arrays = [np.ascontiguousarray(x, np.uint32) for x in (tmpl, gates, calls, items, call_wires, outs)]
        self.root = int(root)
        self.n_templates = self.arrays[0].shape[0] // 12
        r = self.arrays[0][12 * self.root:12 * self.root + 12]
        self.n_inputs, self.n_outputs = int(r[0]), int(r[11])
        self._t = _Templates(self.n_templates, self.root, *[a.ctypes.data for a in self.arrays])

    def garble(self, hasher: int, seed: int, max_gates: int = 0):
        """max_gates > 0: only the first max_gates gates of the emission order (CPU-baseline samples)."""
        inl = np.zeros((self.n_inputs, 16), np.uint8)
        outl = np.zeros((self.n_outputs, 16), np.uint8)
        sm = _Summary()
        rc = lib().gsvo_garble_templates_prefix(hasher, C.c_uint64(seed), C.byref(self._t), C.c_uint64(max_gates),
                                                inl.ctypes.data, outl.ctypes.data, C.byref(sm))
        if rc != 0 and not (rc == 1 and max_gates):
            raise RuntimeError(f"oracle garble failed: {rc}")
        return {"delta": bytes(sm.delta), "false_label0": bytes(sm.false_label0), "true_label0": bytes(sm.true_label0),
                "ct_commit": bytes(sm.ct_commit), "n_ct": int(sm.n_ct), "n_gates": int(sm.n_gates),
                "input_label0": inl, "output_label0": outl}


GEN_LIB_PATH = os.path.join(_HERE, "libgsv_circuitgen.so")
_gen = None


def _genlib() -> C.CDLL:
    global _gen
    if _gen is None:
        if not os.path.exists(GEN_LIB_PATH):
            build()
        L = C.CDLL(GEN_LIB_PATH)
        L.gsvgen_build.restype = C.c_void_p
        L.gsvgen_build.argtypes = [C.c_char_p]
        L.gsvgen_destroy.argtypes = [C.c_void_p]
        L.gsvgen_last_error.restype = C.c_char_p
        L.gsvgen_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 6
        L.gsvgen_totals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _gen = L
    return _gen


class CircuitGen:
    """A named workload circuit recorded by the host-only generator library (oracle/circuitgen.cpp): the
    template DAG the oracle walks, without loading the CUDA engine (bench.py --impl reference)."""

    def __init__(self, circuit: str):
        L = _genlib()
        h = L.gsvgen_build(circuit.encode())
        if not h:
            raise RuntimeError(f"circuit generator: {L.gsvgen_last_error().decode()}")
        try:
            sizes = (C.c_uint64 * 6)()
            root = C.c_uint32(0)
            assert L.gsvgen_export(h, sizes, C.byref(root), None, None, None, None, None, None) == 0
            self.arrays = [np.zeros(max(int(n), 1), np.uint32) for n in sizes]
            assert L.gsvgen_export(h, sizes, C.byref(root), *[a.ctypes.data for a in self.arrays]) == 0
            self.arrays = [a[: int(n)] for a, n in zip(self.arrays, sizes)]
            self.root = int(root.value)
            self.n_templates = self.arrays[0].shape[0] // 12
            self.total_gates = np.zeros(self.n_templates, np.uint64)
            self.total_ct = np.zeros(self.n_templates, np.uint64)
            assert L.gsvgen_totals(h, self.total_gates.ctypes.data, self.total_ct.ctypes.data) == 0
        finally:
            L.gsvgen_destroy(h)
        self.circuit = circuit
        self.n_gates = int(self.total_gates[self.root])
        self.n_ciphertexts = int(self.total_ct[self.root])

    def dag(self, root: int | None = None) -> "TemplateDag":
        return TemplateDag(self.root if root is None else root, *self.arrays)

    def children(self, tmpl: int | None = None):
        """Callee template index of every call of `tmpl` (default: the circuit's root), in emission order."""
        t = self.root if tmpl is None else tmpl
        r = self.arrays[0][12 * t:12 * t + 12]
        calls = self.arrays[2][3 * int(r[4]):3 * (int(r[4]) + int(r[5]))].reshape(-1, 3)
        return [int(c[0]) for c in calls]

"""gsv_b200 -- host-side mirror of the reference's garbling API on top of libgsv_cuda.so.

The reference's entry points for this path are `CircuitBuilder::streaming_garbling` /
`streaming_evaluation` (src/circuit/mod.rs:185-250) driven per instance by the cut-and-choose
loops (src/cut_and_choose/garbler.rs:191-242).  Here the same calls take a *batch* of seeds and
run on one B200 through the C ABI in include/gsv_cuda.h.  There is no CPU fallback: if the CUDA
library or a device is missing the calls raise.

Labels are numpy uint8 arrays whose last axis is the 16 bytes of `S::to_bytes()`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsv_cuda.so")

HASH_AES = 0
HASH_BLAKE3 = 1
CT_NONE = 0
CT_COMMIT = 1
CT_KEEP = 2
CT_KEEP_RAW = 3
CT_COMMIT_HOST = 4
WIRE_UNREACHABLE = 0xFFFFFFFF

GATE_NAMES = ["And", "Nand", "Nimp", "Imp", "Ncimp", "Cimp", "Nor", "Or", "Xor", "Xnor", "Not"]


class GsvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gsv error {code}: {msg}")
        self.code = code


class _PlanOptions(C.Structure):
    _fields_ = [("max_task_gates", C.c_uint64), ("max_task_slots", C.c_uint32), ("lane_only", C.c_uint32),
                ("pipeline", C.c_uint32), ("window_levels", C.c_uint32)]


class _ProgramInfo(C.Structure):
    _fields_ = [
        ("n_gates", C.c_uint64),
        ("n_live_gates", C.c_uint64),
        ("n_ciphertexts", C.c_uint64),
        ("type_count", C.c_uint64 * 11),
        ("n_inputs", C.c_uint32),
        ("n_outputs", C.c_uint32),
        ("n_tasks", C.c_uint32),
        ("n_calls", C.c_uint32),
        ("n_global_slots", C.c_uint32),
        ("max_task_slots", C.c_uint32),
        ("max_task_levels", C.c_uint32),
        ("max_call_deps", C.c_uint32),
        ("sum_call_levels", C.c_uint64),
        ("critical_path_gates", C.c_uint64),
        ("critical_path_levels", C.c_uint64),
    ]


class _SessionOptions(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("n_instances", C.c_uint32),
        ("group", C.c_uint32),
        ("worker_threads", C.c_uint32),
        ("ct_mode", C.c_uint32),
        ("ct_ring_log2", C.c_uint32),
        ("exec_mode", C.c_uint32),
        ("sm_limit", C.c_uint32),
        ("ct_buffer_bytes", C.c_uint64),
        ("host_threads", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class _GarbleResult(C.Structure):
    _fields_ = [
        ("delta", C.c_void_p),
        ("false_label0", C.c_void_p),
        ("true_label0", C.c_void_p),
        ("input_label0", C.c_void_p),
        ("output_label0", C.c_void_p),
        ("ct_commit", C.c_void_p),
        ("n_ciphertexts", C.c_uint64),
        ("ms_seed", C.c_float),
        ("ms_garble", C.c_float),
        ("ms_commit", C.c_float),
        ("ms_total", C.c_float),
        ("n_launches", C.c_uint32),
        ("reserved", C.c_uint32),
        ("host_fold_busy", C.c_float),
        ("host_drain_wait_kernel", C.c_float),
        ("host_drain_wait_fold", C.c_float),
        ("reserved2", C.c_float),
    ]


class _EvaluateIO(C.Structure):
    _fields_ = [
        ("true_label", C.c_void_p),
        ("false_label", C.c_void_p),
        ("input_active", C.c_void_p),
        ("input_bits", C.c_void_p),
        ("ct_streams", C.POINTER(C.c_void_p)),
        ("ct_stream_len", C.c_uint64),
        ("output_active", C.c_void_p),
        ("output_bits", C.c_void_p),
        ("ct_commit", C.c_void_p),
        ("ms_evaluate", C.c_float),
        ("ms_commit", C.c_float),
        ("ms_total", C.c_float),
        ("n_launches", C.c_uint32),
        ("ct_ring_log2", C.c_uint32),
    ]


_lib = None


def load_library() -> C.CDLL:
    """Loads libgsv_cuda.so (built by `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA engine has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    lib.gsv_last_error.restype = C.c_char_p
    lib.gsv_version.restype = C.c_char_p
    lib.gsv_device_count.restype = C.c_int
    lib.gsv_program_build.restype = C.c_void_p
    lib.gsv_program_build.argtypes = [C.c_char_p, C.POINTER(_PlanOptions)]
    lib.gsv_program_destroy.argtypes = [C.c_void_p]
    lib.gsv_program_get_info.argtypes = [C.c_void_p, C.POINTER(_ProgramInfo)]
    lib.gsv_program_flat_stream.restype = C.c_int64
    lib.gsv_program_flat_stream.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_void_p]
    lib.gsv_program_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.gsv_program_execute_plan.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.gsv_groth16_synthetic_inputs.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_uint32]
    lib.gsv_session_create.restype = C.c_void_p
    lib.gsv_session_create.argtypes = [C.c_void_p, C.POINTER(_SessionOptions)]
    lib.gsv_session_destroy.argtypes = [C.c_void_p]
    lib.gsv_garble_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(_GarbleResult)]
    lib.gsv_session_read_ciphertexts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.gsv_evaluate_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(_EvaluateIO)]
    lib.gsv_commit_labels.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.gsv_hash_blocks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.gsv_bench_hash.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise GsvError(rc, load_library().gsv_last_error().decode())


def device_count() -> int:
    return load_library().gsv_device_count()


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Program:
    """A recorded + planned circuit (the once-per-topology flatten step)."""

    def __init__(self, circuit: str, max_task_slots: int = 0, max_task_gates: int = 0, lane_only: bool = False,
                 pipeline: Optional[bool] = None, window_levels: int = 0):
        lib = load_library()
        opt = _PlanOptions(max_task_gates, max_task_slots, 1 if lane_only else 0,
                           0 if pipeline is None else (1 if pipeline else 2), window_levels)
        self._h = lib.gsv_program_build(circuit.encode(), C.byref(opt))
        if not self._h:
            raise GsvError(-1, lib.gsv_last_error().decode())
        self.circuit = circuit
        info = _ProgramInfo()
        _check(lib.gsv_program_get_info(self._h, C.byref(info)))
        self.n_gates = info.n_gates
        self.n_live_gates = info.n_live_gates
        self.n_ciphertexts = info.n_ciphertexts
        self.type_count = list(info.type_count)
        self.n_inputs = info.n_inputs
        self.n_outputs = info.n_outputs
        self.n_tasks = info.n_tasks
        self.n_calls = info.n_calls
        self.n_global_slots = info.n_global_slots
        self.max_task_slots = info.max_task_slots
        self.max_task_levels = info.max_task_levels
        self.critical_path_gates = info.critical_path_gates
        self.critical_path_levels = info.critical_path_levels
        self.max_call_deps = info.max_call_deps
        self.sum_call_levels = info.sum_call_levels

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.gsv_program_destroy(self._h)
            self._h = None

    def execute(self, input_bits) -> np.ndarray:
        """ExecuteMode: boolean evaluation of the recorded topology on the host (self-check)."""
        lib = load_library()
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(self.n_inputs)
        ob = np.zeros(self.n_outputs, np.uint8)
        n = C.c_uint64(0)
        _check(lib.gsv_program_execute(self._h, _ptr(ib), _ptr(ob), C.byref(n)))
        return ob

    def depth(self):
        """(all gates, non-free gates only): gate-level dependency depth of the outputs."""
        lib = load_library()
        a, n = C.c_uint64(0), C.c_uint64(0)
        lib.gsv_program_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _check(lib.gsv_program_depth(self._h, C.byref(a), C.byref(n)))
        return int(a.value), int(n.value)

    def execute_plan(self, input_bits, lane_form: bool = False) -> np.ndarray:
        """Boolean evaluation of the PLANNED program (tasks / calls / recycled slots): planner self-check."""
        lib = load_library()
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(self.n_inputs)
        ob = np.zeros(self.n_outputs, np.uint8)
        _check(lib.gsv_program_execute_plan(self._h, 1 if lane_form else 0, _ptr(ib), _ptr(ob)))
        return ob

    def export_templates(self):
        """(root, tmpl, gates, calls, items, call_wires, outs): the recorded template DAG, for checkers that
        walk circuits too large to flatten (see gsv_program_export_templates)."""
        lib = load_library()
        sizes = (C.c_uint64 * 6)()
        root = C.c_uint32(0)
        lib.gsv_program_export_templates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 6
        _check(lib.gsv_program_export_templates(self._h, sizes, C.byref(root), None, None, None, None, None, None))
        arrs = [np.zeros(max(int(n), 1), np.uint32) for n in sizes]
        _check(lib.gsv_program_export_templates(self._h, sizes, C.byref(root), *[_ptr(a) for a in arrs]))
        return (int(root.value),) + tuple(a[: int(n)] for a, n in zip(arrs, sizes))

    def flat_stream(self):
        """Emission-order gate stream (type, a, b, c, outputs, n_wires) for checkers."""
        lib = load_library()
        n = self.n_gates
        t = np.zeros(n, np.uint8)
        a = np.zeros(n, np.uint32)
        b = np.zeros(n, np.uint32)
        c = np.zeros(n, np.uint32)
        outs = np.zeros(max(self.n_outputs, 1), np.uint32)
        nw = C.c_uint32(0)
        got = lib.gsv_program_flat_stream(self._h, _ptr(t), _ptr(a), _ptr(b), _ptr(c), n, _ptr(outs), C.byref(nw))
        if got < 0:
            _check(int(got))
        return t, a, b, c, outs[: self.n_outputs], nw.value


@dataclass
class GarbleResult:
    """Per-instance results of streaming_garbling (StreamingResult, src/circuit/mod.rs:82-107)."""

    delta: np.ndarray          # [B,16]
    false_label0: np.ndarray   # [B,16]
    true_label0: np.ndarray    # [B,16]
    input_label0: Optional[np.ndarray]   # [B,n_inputs,16]
    output_label0: Optional[np.ndarray]  # [B,n_outputs,16]
    ct_commit: np.ndarray      # [B,16]
    n_ciphertexts: int
    ms_seed: float
    ms_garble: float
    ms_commit: float
    ms_total: float
    n_launches: int
    host_fold_busy: float = 0.0          # CT_COMMIT_HOST: busiest fold thread's share of the call inside the fold
    host_drain_wait_kernel: float = 0.0  # drain loop waiting for the kernel
    host_drain_wait_fold: float = 0.0    # drain loop waiting for the fold threads

    @property
    def true_label1(self) -> np.ndarray:
        return self.true_label0 ^ self.delta


@dataclass
class EvalResult:
    output_active: np.ndarray  # [B,n_outputs,16]
    output_bits: np.ndarray    # [B,n_outputs]
    ct_commit: np.ndarray      # [B,16]
    ms_evaluate: float
    ms_commit: float
    ms_total: float
    n_launches: int


class Session:
    """Device state for a batch of instances of one program on one GPU."""

    def __init__(self, program: Program, n_instances: int, device: int = 0, group: int = 0,
                 worker_threads: int = 0, ct_mode: int = CT_KEEP, ct_ring_log2: int = 0,
                 exec_mode: int = 0, sm_limit: int = 0, ct_buffer_bytes: int = 0, host_threads: int = 0):
        lib = load_library()
        self.program = program
        self.n_instances = n_instances
        self.device = device
        self.ct_mode = ct_mode
        self.group = group
        opt = _SessionOptions(device, n_instances, group, worker_threads, ct_mode, ct_ring_log2, exec_mode,
                              sm_limit, ct_buffer_bytes, host_threads, 0)
        self._h = lib.gsv_session_create(program._h, C.byref(opt))
        if not self._h:
            msg = lib.gsv_last_error().decode()
            raise GsvError(-2 if "no CUDA device" in msg else -3, msg)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.gsv_session_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def garble(self, seeds: Sequence[int], hasher: int = HASH_AES, want_inputs: bool = True,
               want_outputs: bool = True) -> GarbleResult:
        lib = load_library()
        B, p = self.n_instances, self.program
        seeds_a = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        if seeds_a.shape != (B,):
            raise ValueError("need one seed per instance")
        delta = np.zeros((B, 16), np.uint8)
        fl = np.zeros((B, 16), np.uint8)
        tl = np.zeros((B, 16), np.uint8)
        il = np.zeros((B, p.n_inputs, 16), np.uint8) if want_inputs else None
        ol = np.zeros((B, p.n_outputs, 16), np.uint8) if want_outputs else None
        cc = np.zeros((B, 16), np.uint8)
        r = _GarbleResult()
        r.delta, r.false_label0, r.true_label0 = _ptr(delta), _ptr(fl), _ptr(tl)
        r.input_label0, r.output_label0, r.ct_commit = _ptr(il), _ptr(ol), _ptr(cc)
        _check(lib.gsv_garble_batch(self._h, hasher, _ptr(seeds_a), C.byref(r)))
        return GarbleResult(delta, fl, tl, il, ol, cc, int(r.n_ciphertexts), r.ms_seed, r.ms_garble,
                            r.ms_commit, r.ms_total, r.n_launches, r.host_fold_busy, r.host_drain_wait_kernel,
                            r.host_drain_wait_fold)

    def execute(self, input_bits: np.ndarray):
        """ExecuteMode on the GPU: boolean evaluation of the planned circuit for input_bits[n, n_inputs]
        (n <= 128 * n_instances), bit-sliced.  Returns (output_bits[n, n_outputs], kernel milliseconds)."""
        lib = load_library()
        p = self.program
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(-1, p.n_inputs)
        ob = np.zeros((ib.shape[0], p.n_outputs), np.uint8)
        ms = C.c_float(0)
        lib.gsv_execute_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_float)]
        _check(lib.gsv_execute_batch(self._h, _ptr(ib), ib.shape[0], _ptr(ob), C.byref(ms)))
        return ob, ms.value

    def set_ciphertext_files(self, paths: Optional[Sequence[Optional[str]]]) -> None:
        """gc_{i}.bin writers (ciphertext_repository.rs:94-106) for the next garbling runs of a CT_COMMIT_HOST
        session (or the runs a linked evaluator receives): instance i's stream goes to paths[i] (None = skip).
        Pass None to close the files and clear the sink."""
        lib = load_library()
        lib.gsv_session_set_ciphertext_files.argtypes = [C.c_void_p, C.c_void_p]
        for fd in getattr(self, "_ct_fds", []):
            if fd >= 0:
                os.close(fd)
        self._ct_fds = []
        if paths is None:
            _check(lib.gsv_session_set_ciphertext_files(self._h, None))
            return
        if len(paths) != self.n_instances:
            raise ValueError("need one path (or None) per instance")
        self._ct_fds = [os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) if p else -1 for p in paths]
        arr = (C.c_int * len(paths))(*self._ct_fds)
        _check(lib.gsv_session_set_ciphertext_files(self._h, arr))

    def expand_seeds(self, seeds: Sequence[int]) -> GarbleResult:
        """Delta, constants and input label0s of `seeds` (the seed expansion alone, no garbling)."""
        lib = load_library()
        B, p = self.n_instances, self.program
        seeds_a = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        if seeds_a.shape != (B,):
            raise ValueError("need one seed per instance")
        delta, fl, tl = (np.zeros((B, 16), np.uint8) for _ in range(3))
        il = np.zeros((B, p.n_inputs, 16), np.uint8)
        r = _GarbleResult()
        r.delta, r.false_label0, r.true_label0, r.input_label0 = _ptr(delta), _ptr(fl), _ptr(tl), _ptr(il)
        lib.gsv_session_expand_seeds.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_GarbleResult)]
        _check(lib.gsv_session_expand_seeds(self._h, _ptr(seeds_a), C.byref(r)))
        return GarbleResult(delta, fl, tl, il, None, np.zeros((B, 16), np.uint8), int(r.n_ciphertexts), 0.0, 0.0, 0.0, 0.0,
                            r.n_launches)

    def read_ciphertexts(self, instance: int, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        lib = load_library()
        if count is None:
            count = self.program.n_ciphertexts - first
        out = np.zeros((count, 16), np.uint8)
        _check(lib.gsv_session_read_ciphertexts(self._h, instance, first, count, _ptr(out)))
        return out

    def evaluate(self, hasher: int, true_label: np.ndarray, false_label: np.ndarray,
                 input_active: np.ndarray, input_bits: np.ndarray,
                 ct_streams: Optional[Sequence[np.ndarray]] = None, want_commit: bool = True,
                 ct_ring_log2: int = 0) -> EvalResult:
        lib = load_library()
        B, p = self.n_instances, self.program
        tl = np.ascontiguousarray(true_label, np.uint8).reshape(B, 16)
        fl = np.ascontiguousarray(false_label, np.uint8).reshape(B, 16)
        ia = np.ascontiguousarray(input_active, np.uint8).reshape(B, p.n_inputs, 16)
        ib = np.ascontiguousarray(input_bits, np.uint8).reshape(B, p.n_inputs)
        oa = np.zeros((B, p.n_outputs, 16), np.uint8)
        ob = np.zeros((B, p.n_outputs), np.uint8)
        cc = np.zeros((B, 16), np.uint8)
        io = _EvaluateIO()
        io.true_label, io.false_label, io.input_active, io.input_bits = _ptr(tl), _ptr(fl), _ptr(ia), _ptr(ib)
        keep = None
        if ct_streams is not None:
            # np.memmap'ed gc_{i}.bin files are passed through as they are (no copy): the library reads them
            keep = [s if isinstance(s, np.memmap) and s.dtype == np.uint8 else np.ascontiguousarray(s, np.uint8) for s in ct_streams]
            lens = {k.size // 16 for k in keep}
            if len(keep) != B or len(lens) != 1:
                raise ValueError("need B ciphertext streams of equal length")
            arr = (C.c_void_p * B)(*[k.ctypes.data for k in keep])
            io.ct_streams = C.cast(arr, C.POINTER(C.c_void_p))
            io.ct_stream_len = lens.pop()
            io.ct_ring_log2 = ct_ring_log2
        io.output_active, io.output_bits, io.ct_commit = _ptr(oa), _ptr(ob), _ptr(cc if want_commit else None)
        _check(lib.gsv_evaluate_batch(self._h, hasher, C.byref(io)))
        return EvalResult(oa, ob, cc, io.ms_evaluate, io.ms_commit, io.ms_total, io.n_launches)


def link_sessions(garbler: "Session", evaluator: "Session", ring_bytes: int = 0) -> None:
    """Garbler -> evaluator streaming (gsv_session_link): afterwards call garbler.garble(..) and
    evaluator.evaluate(.., ct_streams=None) concurrently from two threads (see `stream_garble_evaluate`)."""
    lib = load_library()
    lib.gsv_session_link.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    _check(lib.gsv_session_link(garbler._h, evaluator._h, ring_bytes))
    garbler._linked = evaluator._linked = True


def stream_garble_evaluate(garbler: "Session", evaluator: "Session", seeds, hasher: int, input_bits: np.ndarray):
    """One run over a linked pair (examples/groth16_garble.rs:170-267): the garbler regarbles from `seeds`
    while the evaluator consumes the ciphertext ring.  The evaluator's inputs are built like the reference's
    G2EMsg::Commit: active input labels = select(bit), constants = (true.label1, false.label0), all pure
    functions of the seed.  Returns (GarbleResult, EvalResult)."""
    import threading

    p = garbler.program
    B = garbler.n_instances
    bits = np.ascontiguousarray(input_bits, np.uint8).reshape(B, p.n_inputs)
    lab = garbler.expand_seeds(seeds)
    active = lab.input_label0 ^ (lab.delta[:, None, :] * bits[:, :, None])
    out = {}

    def run_g():
        try:
            out["g"] = garbler.garble(seeds, hasher)
        except Exception as e:  # pragma: no cover
            out["g_err"] = e

    tg = threading.Thread(target=run_g)
    tg.start()
    try:
        ev = evaluator.evaluate(hasher, lab.true_label0 ^ lab.delta, lab.false_label0, active, bits)
    finally:
        tg.join()
    if "g_err" in out:
        raise out["g_err"]
    return out["g"], ev


def host_chain_fold(h: np.ndarray, blocks: np.ndarray, instance_major: bool = False) -> np.ndarray:
    """Host half of CT_COMMIT_HOST: blocks[n_pos, n_inst, 16] (or [n_inst, n_pos, 16] when
    instance_major) folded into h[n_inst, 16] (returns a copy)."""
    lib = load_library()
    blocks = np.ascontiguousarray(blocks, np.uint8)
    n_inst, n_pos = (blocks.shape[0], blocks.shape[1]) if instance_major else (blocks.shape[1], blocks.shape[0])
    out = np.ascontiguousarray(h, np.uint8).reshape(n_inst, 16).copy()
    lib.gsv_host_chain_fold.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    ps, is_ = (1, n_pos) if instance_major else (n_inst, 1)
    _check(lib.gsv_host_chain_fold(_ptr(out), _ptr(blocks), ps, is_, n_pos, n_inst))
    return out


def host_chain_fold_quads(h: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """The fold over CT_COMMIT_HOST's drain layout: rows[n_quads, n_pos, 4, 16] folded into
    h[4 * n_quads, 16] (returns a copy)."""
    lib = load_library()
    rows = np.ascontiguousarray(rows, np.uint8)
    n_quads, n_pos = rows.shape[0], rows.shape[1]
    out = np.ascontiguousarray(h, np.uint8).reshape(4 * n_quads, 16).copy()
    lib.gsv_host_chain_fold_quads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
    _check(lib.gsv_host_chain_fold_quads(_ptr(out), _ptr(rows), n_pos * 64, n_pos, n_quads))
    return out


def groth16_synthetic_inputs(public_x: int = 424242, flip_public: bool = False, compressed: bool = True) -> np.ndarray:
    """Input bits of a synthetic proof: 1273 for "groth16_verify_compressed", 2286 (affine x, y per point,
    src/garbled_groth16.rs:141-176) for "groth16_verify"."""
    lib = load_library()
    n = 1273 if compressed else 2286
    bits = np.zeros(n, np.uint8)
    _check(lib.gsv_groth16_synthetic_inputs(public_x, 1 if flip_public else 0, _ptr(bits), n))
    return bits


def commit_labels(labels: np.ndarray, device: int = 0) -> np.ndarray:
    """commit_label for many labels (src/cut_and_choose/mod.rs:41-48)."""
    lib = load_library()
    a = np.ascontiguousarray(labels, np.uint8).reshape(-1, 16)
    out = np.zeros_like(a)
    _check(lib.gsv_commit_labels(device, _ptr(a), a.shape[0], _ptr(out)))
    return out.reshape(np.asarray(labels).shape)


def hash_blocks(hasher: int, x: np.ndarray, gid: np.ndarray, device: int = 0) -> np.ndarray:
    lib = load_library()
    a = np.ascontiguousarray(x, np.uint8).reshape(-1, 16)
    g = np.ascontiguousarray(gid, np.uint64).reshape(-1)
    out = np.zeros_like(a)
    _check(lib.gsv_hash_blocks(device, hasher, _ptr(a), _ptr(g), a.shape[0], _ptr(out)))
    return out


def bench_hash(hasher: int, n_blocks: int = 1 << 30, iters: int = 3, device: int = 0) -> float:
    lib = load_library()
    r = C.c_double(0)
    _check(lib.gsv_bench_hash(device, hasher, n_blocks, iters, C.byref(r)))
    return r.value


def bench_hash_latency(hasher: int, warps_per_sm: int = 1, n: int = 20000, device: int = 0) -> float:
    """SM cycles per dependent gate hash with `warps_per_sm` warps chaining hashes on every SM."""
    lib = load_library()
    r = C.c_double(0)
    lib.gsv_bench_hash_latency.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)]
    _check(lib.gsv_bench_hash_latency(device, hasher, warps_per_sm, n, C.byref(r)))
    return r.value


def streaming_garbling(program: Program, seeds: Sequence[int], hasher: int = HASH_AES,
                       ct_mode: int = CT_COMMIT, device: int = 0, **session_kw) -> GarbleResult:
    """CircuitBuilder::streaming_garbling for a batch of seeds (one-shot session)."""
    s = Session(program, len(seeds), device=device, ct_mode=ct_mode, **session_kw)
    try:
        return s.garble(seeds, hasher)
    finally:
        s.close()

"""Cut-and-choose garbling stage on GPUs -- host-side mirror of `src/cut_and_choose/garbler.rs`.

Reference flow (`Garbler::create` :191-242, `Garbler::commit` :244-257, `GarbledInstanceCommit::new`
:85-99): draw one u64 seed per instance from the caller's RNG, garble every instance with
`AesNiHasher` + `AESAccumulatingHash` (one instance per core), then commit to the ciphertext stream,
to both labels of every input wire, to the output labels and to the constants with
`commit(label) = AES_K(label)` (`src/cut_and_choose/mod.rs:41-65`).

Here the instances of one rank are ONE batched GPU call; ranks (one per GPU) own contiguous shards
of the instance range, and the only collective is the all-gather of the fixed-size commit records
(SURVEY.md section 8e).  No CPU fallback: garbling and label commitment run through libgsv_cuda.so.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

_M64 = (1 << 64) - 1


def _seed_key(seed: int) -> List[int]:
    """rand_core 0.6.4 `SeedableRng::seed_from_u64`: PCG32 (XSH-RR) -> 8 little-endian key words."""
    mul, inc, st, out = 6364136223846793005, 11634580027462260723, seed & _M64, []
    for _ in range(8):
        st = (st * mul + inc) & _M64
        xs = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
        rot = st >> 59
        out.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return out


def _chacha20_block(key: List[int], counter: int) -> List[int]:
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + key + [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, 0, 0]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, s)]


def instance_seeds(master_seed: int, total: int) -> np.ndarray:
    """`seeds[i] = rng.gen::<u64>()` for `rng = ChaCha20Rng::seed_from_u64(master_seed)`
    (garbler.rs:201-203; `next_u64` = two consecutive little-endian words, low first)."""
    key = _seed_key(master_seed)
    words: List[int] = []
    blk = 0
    while len(words) < 2 * total:
        words += _chacha20_block(key, blk)
        blk += 1
    return np.array([words[2 * i] | (words[2 * i + 1] << 32) for i in range(total)], dtype=np.uint64)


def shard(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of instance indices owned by `rank`: [first, first + count)."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


@dataclass
class CommitRecords:
    """Per-instance `GarbledInstanceCommit` as one fixed-size byte record:
    ct_commit | (c(l0), c(l1)) per input | (c(out.label1), c(out.label0)) per output |
    c(true.label1) | c(false.label0); every field 16 bytes (garbler.rs:85-99)."""

    records: np.ndarray  # [n_instances, record_len] uint8
    n_inputs: int
    n_outputs: int

    @staticmethod
    def record_len(n_inputs: int, n_outputs: int) -> int:
        return 16 * (1 + 2 * n_inputs + 2 * n_outputs + 2)

    def ct_commit(self) -> np.ndarray:
        return self.records[:, :16]

    def input_commits(self) -> np.ndarray:
        return self.records[:, 16:16 + 32 * self.n_inputs].reshape(-1, self.n_inputs, 2, 16)

    def output_commits(self) -> np.ndarray:
        o = 16 + 32 * self.n_inputs
        return self.records[:, o:o + 32 * self.n_outputs].reshape(-1, self.n_outputs, 2, 16)

    def constant_commits(self) -> np.ndarray:
        return self.records[:, -32:].reshape(-1, 2, 16)


class Garbler:
    """`Garbler::create` + `commit` for the instances of one rank on one GPU."""

    def __init__(self, program, total: int, master_seed: int, device: int = 0, rank: int = 0, world: int = 1,
                 hasher: int = 0, ct_mode: Optional[int] = None, seeds: Optional[np.ndarray] = None, **session_kw):
        from . import CT_COMMIT, Session

        self.program, self.total, self.rank, self.world, self.device = program, total, rank, world, device
        self.hasher = hasher
        # `seeds` overrides the master-seed draw (the evaluator re-garbles from revealed seeds)
        self.seeds_all = instance_seeds(master_seed, total) if seeds is None else np.asarray(seeds, dtype=np.uint64)
        if self.seeds_all.shape != (total,):
            raise ValueError("need one seed per instance")
        self.first, self.count = shard(total, world, rank)
        self.seeds = self.seeds_all[self.first:self.first + self.count]
        self.session = Session(program, max(self.count, 1), device=device,
                               ct_mode=CT_COMMIT if ct_mode is None else ct_mode, **session_kw) if self.count else None
        self.result = None

    def create(self):
        """Garble the local shard (the hot loop of garbler.rs:206-234) in one batched call."""
        if self.count:
            self.result = self.session.garble(self.seeds, self.hasher)
        return self.result

    def commit(self) -> CommitRecords:
        """`GarbledInstanceCommit::new` for every local instance; label commits run on the GPU."""
        from . import commit_labels

        p = self.program
        n = self.count
        rec = np.zeros((n, CommitRecords.record_len(p.n_inputs, p.n_outputs)), np.uint8)
        if n == 0:
            return CommitRecords(rec, p.n_inputs, p.n_outputs)
        r = self.result if self.result is not None else self.create()
        d = r.delta[:, None, :]
        labels = np.concatenate([
            np.stack([r.input_label0, r.input_label0 ^ d], axis=2).reshape(n, -1, 16),      # (l0, l1) per input
            np.stack([r.output_label0 ^ d, r.output_label0], axis=2).reshape(n, -1, 16),     # (label1, label0) per output
            (r.true_label0 ^ r.delta)[:, None, :],                                           # true.select(true)
            r.false_label0[:, None, :],                                                      # false.select(false)
        ], axis=1)
        rec[:, :16] = r.ct_commit
        rec[:, 16:] = commit_labels(labels, device=self.device).reshape(n, -1)
        return CommitRecords(rec, p.n_inputs, p.n_outputs)


def gather_commits(local: CommitRecords, total: int, group=None) -> CommitRecords:
    """All-gather of the per-rank commit records into instance order (the only collective of the
    path).  Works on any torch.distributed backend: NCCL for GPU ranks, gloo in the CPU tests."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rec_len = local.records.shape[1]
    per = max(shard(total, world, r)[1] for r in range(world))
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((per, rec_len), dtype=torch.uint8, device=dev)
    if local.records.shape[0]:
        buf[: local.records.shape[0]] = torch.from_numpy(local.records).to(dev)
    out = torch.empty((world * per, rec_len), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, per, rec_len)
    parts = [out[r, : shard(total, world, r)[1]] for r in range(world)]
    return CommitRecords(np.concatenate(parts, axis=0), local.n_inputs, local.n_outputs)


# ======================================================================================================
# The rest of the protocol surface: open / re-garble / evaluate_from (garbler.rs:259-319,
# evaluator.rs:45-181, 338-476).  Host logic only; every garbling, evaluation and label commitment below is
# a batched call into libgsv_cuda.so.
# ======================================================================================================
class ConsistencyError(Exception):
    """evaluator.rs:222-336 `ConsistencyError`: `kind` is the variant name, `index` the instance."""

    def __init__(self, kind: str, index: int, detail: str = ""):
        super().__init__(f"{kind} (instance {index}) {detail}".strip())
        self.kind, self.index = kind, index


class _ChaChaU64:
    """`ChaCha20Rng::seed_from_u64(seed)` as a stream of `next_u64` draws."""

    def __init__(self, seed: int):
        self.key, self.blk, self.words = _seed_key(seed), 0, []

    def next_u64(self) -> int:
        if len(self.words) < 2:
            self.words += _chacha20_block(self.key, self.blk)
            self.blk += 1
        lo, hi = self.words[0], self.words[1]
        del self.words[:2]
        return lo | (hi << 32)

    def gen_range_inclusive(self, high: int) -> int:
        """`rng.gen_range(0..=high)` for usize: rand 0.8.5 `UniformInt::sample_single_inclusive`
        (widening multiply, rejection zone `(range << lz) - 1`).  Parity unpinned (crate not vendored)."""
        rng_range = (high + 1) & _M64
        if rng_range == 0:
            return self.next_u64()
        lz = 64 - rng_range.bit_length()
        zone = ((rng_range << lz) & _M64) - 1
        while True:
            m = self.next_u64() * rng_range
            if (m & _M64) <= zone:
                return m >> 64


def choose_to_finalize(rng_seed: int, total: int, to_finalize: int) -> List[int]:
    """`Evaluator::create` (evaluator.rs:45-70): Fisher-Yates over 0..total with `gen_range(0..=i)`,
    first `to_finalize` entries, sorted."""
    if to_finalize > total:
        raise ValueError("to_finalize must be <= total")
    rng, idx = _ChaChaU64(rng_seed), list(range(total))
    for i in range(total - 1, 0, -1):
        j = rng.gen_range_inclusive(i)
        idx[i], idx[j] = idx[j], idx[i]
    return sorted(idx[:to_finalize])


@dataclass
class EvaluatorCaseInput:
    """`EvaluatorCaseInput` (cut_and_choose/mod.rs): what the garbler hands over for one finalized
    instance -- the active input labels with their bits, and the two constant labels."""

    index: int
    input_active: np.ndarray  # [n_inputs, 16]
    input_bits: np.ndarray    # [n_inputs]
    true_label: np.ndarray    # [16]  true.select(true)
    false_label: np.ndarray   # [16]  false.select(false)


def open_commit(garbler: "Garbler", to_finalize: List[int], ct_dir: Optional[str] = None):
    """`Garbler::open_commit` for the local shard (garbler.rs:259-319): seeds of the opened instances, and for the
    finalized ones a re-garbling that hands the ciphertext stream over.  Returns (open: [(index, seed)], closed).

    ct_dir given (the reference's `FileCiphertextHandlerProvider`, ciphertext_repository.rs:59-136): the finalized
    instances are re-garbled in ONE GSV_CT_COMMIT_HOST run whose host drain writes `ct_dir/gc_{index}.bin` directly
    from the pinned buffers -- nothing is kept in HBM, so verifier instances (47.7 GB each) work; closed maps
    index -> file path.  Without ct_dir the streams are kept on the GPU (`Sender<S>` handler; must fit HBM) and
    closed maps index -> [n_ct, 16] array."""
    import os

    from . import CT_COMMIT_HOST, CT_KEEP_RAW, Session

    mine = [i for i in to_finalize if garbler.first <= i < garbler.first + garbler.count]
    open_ = [(garbler.first + k, int(garbler.seeds[k])) for k in range(garbler.count)
             if garbler.first + k not in mine]
    closed = {}
    if mine and ct_dir is not None:
        os.makedirs(ct_dir, exist_ok=True)
        paths = [os.path.join(ct_dir, f"gc_{i}.bin") for i in mine]
        sess = Session(garbler.program, len(mine), device=garbler.device, ct_mode=CT_COMMIT_HOST, exec_mode=1)
        sess.set_ciphertext_files(paths)
        sess.garble([int(garbler.seeds_all[i]) for i in mine], garbler.hasher, want_inputs=False, want_outputs=False)
        sess.set_ciphertext_files(None)
        sess.close()
        closed = dict(zip(mine, paths))
    elif mine:
        sess = Session(garbler.program, len(mine), device=garbler.device, ct_mode=CT_KEEP_RAW)
        sess.garble([int(garbler.seeds_all[i]) for i in mine], garbler.hasher, want_inputs=False, want_outputs=False)
        for k, i in enumerate(mine):
            closed[i] = sess.read_ciphertexts(k)   # the bytes of gc_{i}.bin (ciphertext_repository.rs:94-106)
        sess.close()
    return open_, closed


def _as_stream(x) -> np.ndarray:
    """A closed ciphertext stream as a flat uint8 array; file paths are memory-mapped (FileSource)."""
    if isinstance(x, (str, bytes)) or hasattr(x, "__fspath__"):
        return np.memmap(x, dtype=np.uint8, mode="r")
    return np.ascontiguousarray(x, np.uint8).reshape(-1)


def prepare_input_labels(garbler: "Garbler", to_finalize: List[int], input_bits: np.ndarray) -> List[EvaluatorCaseInput]:
    """`Garbler::prepare_input_labels` (cut_and_choose/groth16.rs:71-101): active labels of the real input
    for every finalized local instance."""
    r = garbler.result if garbler.result is not None else garbler.create()
    bits = np.ascontiguousarray(input_bits, np.uint8).reshape(garbler.program.n_inputs)
    cases = []
    for i in to_finalize:
        k = i - garbler.first
        if not 0 <= k < garbler.count:
            continue
        act = r.input_label0[k].copy()
        act[bits.astype(bool)] ^= r.delta[k]
        cases.append(EvaluatorCaseInput(i, act, bits.copy(), r.true_label0[k] ^ r.delta[k], r.false_label0[k].copy()))
    return cases


class Evaluator:
    """`Evaluator` (src/cut_and_choose/evaluator.rs): holds the garbler's commit records, picks the
    instances to finalize, checks the opened ones by re-garbling and evaluates the finalized ones."""

    def __init__(self, program, total: int, to_finalize: int, rng_seed: int, commits: CommitRecords,
                 device: int = 0, hasher: int = 0):
        if commits.records.shape[0] != total:
            raise ValueError("need one commit record per instance")
        self.program, self.total, self.commits, self.device, self.hasher = program, total, commits, device, hasher
        self.to_finalize = choose_to_finalize(rng_seed, total, to_finalize)

    def run_regarbling(self, open_seeds: List[Tuple[int, int]], closed_streams) -> None:
        """evaluator.rs:83-181.  Opened instances: full re-garble from the revealed seed (ONE batched GPU
        call for all of them), rebuild the commit record, compare.  Finalized instances: fold the received
        ciphertext stream and compare with the committed chain hash."""
        from . import host_chain_fold

        seeds = dict(open_seeds)
        opened = [i for i in range(self.total) if i not in self.to_finalize]
        for i in opened:
            if i not in seeds:
                raise ConsistencyError("MissingSeed", i)
        if opened:
            g = Garbler(self.program, len(opened), 0, device=self.device, hasher=self.hasher,
                        seeds=np.array([seeds[i] for i in opened], dtype=np.uint64))
            rec = g.commit().records
            g.session.close()
            for k, i in enumerate(opened):
                if not np.array_equal(rec[k], self.commits.records[i]):
                    raise ConsistencyError("RegarblingMismatch", i)
        for i in self.to_finalize:
            if i not in closed_streams:
                raise ConsistencyError("MissingCiphertextHash", i)
        # fold the received streams (files are read chunk by chunk), all finalized instances interleaved
        fin = list(self.to_finalize)
        streams = [_as_stream(closed_streams[i]) for i in fin]
        h = np.zeros((len(fin), 16), np.uint8)
        n_pos = min(s.size for s in streams) // 16 if streams else 0
        if any(s.size != streams[0].size for s in streams):
            n_pos = 0   # ragged: fold one by one below
        step = 1 << 22
        for a in range(0, n_pos, step):
            b = min(n_pos, a + step)
            h = host_chain_fold(h, np.stack([s[16 * a:16 * b].reshape(-1, 16) for s in streams]), instance_major=True)
        if n_pos == 0:
            for k, s_ in enumerate(streams):
                for a in range(0, s_.size // 16, step):
                    b = min(s_.size // 16, a + step)
                    h[k:k + 1] = host_chain_fold(h[k:k + 1], s_[16 * a:16 * b].reshape(1, -1, 16), instance_major=True)
        for k, i in enumerate(fin):
            if not np.array_equal(h[k], self.commits.ct_commit()[i]):
                raise ConsistencyError("CiphertextMismatch", i, "ciphertext corrupted")

    def evaluate_from(self, closed_streams, cases: List[EvaluatorCaseInput]):
        """evaluator.rs:338-476: for every finalized instance check the constant and input-label commits,
        evaluate from its ciphertext stream (one batched GPU call), re-check the chain hash and the output
        label commit.  Returns [(index, output bits [n_out], active output labels [n_out, 16])]."""
        from . import CT_NONE, Session, commit_labels

        p, n = self.program, len(cases)
        if n == 0:
            return []
        for c in cases:
            if c.index not in self.to_finalize:
                raise ConsistencyError("NotFinalized", c.index)
            if c.index not in closed_streams:
                raise ConsistencyError("MissingCiphertextHash", c.index)
            if c.input_active.shape[0] != p.n_inputs:
                raise ConsistencyError("InputLabelsCountMismatch", c.index)
        idx = [c.index for c in cases]
        consts = commit_labels(np.stack([np.stack([c.true_label, c.false_label]) for c in cases]), device=self.device)
        in_c = commit_labels(np.stack([c.input_active for c in cases]), device=self.device).reshape(n, p.n_inputs, 16)
        for k, c in enumerate(cases):
            want = self.commits.constant_commits()[c.index]
            if not np.array_equal(consts.reshape(n, 2, 16)[k, 0], want[0]):
                raise ConsistencyError("TrueConstantMismatch", c.index)
            if not np.array_equal(consts.reshape(n, 2, 16)[k, 1], want[1]):
                raise ConsistencyError("FalseConstantMismatch", c.index)
            exp = self.commits.input_commits()[c.index]            # [n_in, 2, 16]: (c(l0), c(l1))
            sel = exp[np.arange(p.n_inputs), c.input_bits.astype(np.int64)]
            bad = np.nonzero((sel != in_c[k]).any(axis=1))[0]
            if bad.size:
                raise ConsistencyError("InputLabelsMismatch", c.index, f"label_index {int(bad[0])}")
        # streams may be arrays or gc_{i}.bin paths; the library uploads them whole when they fit and otherwise feeds
        # them through a ring while the kernel runs (FileSource: hashed on host threads while being consumed)
        sess = Session(p, n, device=self.device, ct_mode=CT_NONE)
        ev = sess.evaluate(self.hasher, np.stack([c.true_label for c in cases]), np.stack([c.false_label for c in cases]),
                           np.stack([c.input_active for c in cases]), np.stack([c.input_bits for c in cases]),
                           ct_streams=[_as_stream(closed_streams[i]) for i in idx])
        sess.close()
        out_c = commit_labels(ev.output_active, device=self.device).reshape(n, p.n_outputs, 16)
        res = []
        for k, c in enumerate(cases):
            if not np.array_equal(ev.ct_commit[k], self.commits.ct_commit()[c.index]):
                raise ConsistencyError("CiphertextMismatch", c.index)
            exp = self.commits.output_commits()[c.index]           # [n_out, 2, 16]: (c(label1), c(label0))
            sel = exp[np.arange(p.n_outputs), 1 - ev.output_bits[k].astype(np.int64)]
            if (sel != out_c[k]).any():
                raise ConsistencyError("OutputLabelMismatch", c.index)
            res.append((c.index, ev.output_bits[k].copy(), ev.output_active[k].copy()))
        return res

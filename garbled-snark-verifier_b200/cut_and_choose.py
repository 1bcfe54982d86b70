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
                 hasher: int = 0, ct_mode: Optional[int] = None, **session_kw):
        from . import CT_COMMIT, Session

        self.program, self.total, self.rank, self.world, self.device = program, total, rank, world, device
        self.hasher = hasher
        self.seeds_all = instance_seeds(master_seed, total)
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

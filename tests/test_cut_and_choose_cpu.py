"""Host logic of the cut-and-choose garbling stage: seed derivation, sharding, and the N>1
commit gather over torch.distributed (gloo, world_size 2, CPU)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_instance_seeds_match_oracle_rng(gsv, orc):
    """seeds[i] = rng.gen::<u64>() of ChaCha20Rng::seed_from_u64(master) (garbler.rs:201-203;
    SURVEY.md section 8d config 4 uses master seed 1234)."""
    from gsv_b200 import cut_and_choose as cc

    for master in (0, 1234, 2**64 - 1):
        r = orc.Rng(master)
        want = [r.u64() for _ in range(40)]
        assert list(cc.instance_seeds(master, 40)) == want


def test_shard_partition():
    from gsv_b200 import cut_and_choose as cc

    for total in (0, 1, 5, 16, 17, 128):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                first, count = cc.shard(total, world, r)
                cover += list(range(first, first + count))
            assert cover == list(range(total))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import gsv_b200  # noqa: F401
    from gsv_b200 import cut_and_choose as cc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    first, count = cc.shard(total, world, rank)
    n_in, n_out = 3, 2
    rec = np.zeros((count, cc.CommitRecords.record_len(n_in, n_out)), np.uint8)
    for i in range(count):
        rec[i] = (first + i + 1) % 251  # instance-identifying filler
    got = cc.gather_commits(cc.CommitRecords(rec, n_in, n_out), total)
    q.put((rank, got.records[:, 0].tolist(), got.records.shape))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_gather_commits_gloo_world2(built, total):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, col, shape in res:
        assert shape[0] == total
        assert col == [(i + 1) % 251 for i in range(total)]  # instance order, ragged shards handled


def test_choose_to_finalize_is_a_seeded_sample_without_replacement():
    """Evaluator::create (evaluator.rs:45-70): sorted, distinct, in range, deterministic per seed."""
    import importlib

    cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
    for total, k in ((5, 2), (16, 7), (181, 7), (1, 1), (4, 0)):
        a = cc.choose_to_finalize(1234, total, k)
        assert a == sorted(set(a)) and len(a) == k and all(0 <= i < total for i in a)
        assert a == cc.choose_to_finalize(1234, total, k)
    assert cc.choose_to_finalize(1, 181, 7) != cc.choose_to_finalize(2, 181, 7)
    with pytest.raises(ValueError):
        cc.choose_to_finalize(0, 3, 4)
    # every index is reachable and roughly uniform (Fisher-Yates with an unbiased gen_range)
    hits = np.zeros(10, int)
    for s in range(400):
        for i in cc.choose_to_finalize(s, 10, 3):
            hits[i] += 1
    assert hits.min() > 80 and hits.max() < 160


def test_rng_u64_stream_matches_instance_seeds_and_gen_range_zone():
    import importlib

    cc = importlib.import_module("garbled-snark-verifier_b200.cut_and_choose")
    r = cc._ChaChaU64(1234)
    assert [r.next_u64() for _ in range(40)] == [int(x) for x in cc.instance_seeds(1234, 40)]
    # rand 0.8.5 sample_single_inclusive: v*range >> 64, rejected when the low half exceeds the zone
    r = cc._ChaChaU64(7)
    draws = [r.gen_range_inclusive(5) for _ in range(2000)]
    assert set(draws) == set(range(6))
    r2 = cc._ChaChaU64(7)
    v = r2.next_u64()
    assert draws[0] == (v * 6) >> 64 or ((v * 6) & ((1 << 64) - 1)) > ((6 << 61) - 1)

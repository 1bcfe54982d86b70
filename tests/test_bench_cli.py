"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys,
and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "batch",
                           "--circuit", "fq_mul", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line(built):
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "garbled_gates_per_s" and d["unit"] == "gates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the reference arm runs the CPU oracle and the host-only circuit generator, never the CUDA engine
    assert d["native_libs"] == ["oracle/libgsv_circuitgen.so", "oracle/libgsv_oracle.so"]


def test_reference_arm_is_silent_on_other_ranks(built):
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_host_plan_for_the_drivers_hosts():
    """bench.host_plan: steps in flight, instances per step and fold threads per step from a rank's CPU share.
    (The committed step folds its serial chains on host AES-NI threads; this is the host-dependent part of the bench.)"""
    sys.path.insert(0, ROOT)
    import bench

    n_ct = 2_980_239_027  # the verifier
    # 1 GPU on a 16-CPU host, the driver's --steps 20 --warmup 5: 5 waves of 4 beat 7 waves of 3; a thread per quad
    assert bench.host_plan(3, 16, 20, 5, 16, n_ct) == (4, 16, 4, None)
    # the default run (6 steps): 2 waves of 3
    assert bench.host_plan(3, 16, 6, 3, 16, n_ct) == (3, 16, 4, None)
    # 2 GPUs on a 24-CPU host: 12 CPUs per rank, 3 steps in flight, a fold thread per quad
    assert bench.host_plan(3, 16, 20, 5, 12, n_ct) == (3, 16, 4, None)
    # 8 CPUs per rank: one CPU is left to the drain threads, 2 fold threads per step (two quads each)
    assert bench.host_plan(3, 16, 20, 5, 8, n_ct) == (3, 16, 2, None)
    # 4 GPUs on a 16-CPU host: a session per CPU, one fold thread each, the step still fits
    assert bench.host_plan(3, 16, 20, 5, 4, n_ct) == (4, 16, 1, None)
    # 8 GPUs on a 16-CPU host: two CPUs per rank cannot fold 16 chains per step in time -> smaller steps, said so
    S, B, T, note = bench.host_plan(3, 16, 20, 5, 2, n_ct)
    assert (S, B, T) == (3, 8, 1) and "8 instead of 16" in note
    # 8 GPUs on a large host (16 CPUs per rank): the folds have the CPUs, but eight drains + folds share the host's
    # memory bandwidth -> smaller steps; 4 GPUs still fit
    S, B, T, note = bench.host_plan(3, 16, 20, 5, 16, n_ct, ranks_on_host=8)
    assert (S, B, T) == (4, 8, 2) and "memory bandwidth" in note
    assert bench.host_plan(3, 16, 20, 5, 16, n_ct, ranks_on_host=4) == (4, 16, 4, None)
    assert bench.host_plan(3, 16, 20, 5, 12, n_ct, ranks_on_host=2) == (3, 16, 4, None)
    # explicit flags are left alone
    assert bench.host_plan(2, 32, 20, 5, 2, n_ct, sessions_auto=False, instances_auto=False, host_threads=5) == (2, 32, 5, None)
    # no host fold without the host-folded commitment
    assert bench.host_plan(3, 16, 20, 5, 2, n_ct, commit_host=False)[:2] == (3, 16)

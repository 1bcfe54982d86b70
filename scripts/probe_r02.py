"""Round-2 exploration probe (not the bench): PCIe copy rates and per-phase worker cycles (GSV_PROFILE)
of the levelised kernel on circuits that resemble the verifier's critical path.

usage: probe_r02.py [pcie] [latency] [small] [verifier] [verifier_shapes=4x256,2x256] [circuits=a,b]
Set GSV_PROFILE=1 to get the per-phase cycle split (costs a few percent).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsv_b200 as g

what = set(a for a in sys.argv[1:] if "=" not in a) or {"pcie", "small"}
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)


def run(prog, B, shapes, ct_mode=g.CT_NONE, reps=2):
    for sh in shapes:
        G, NT = (int(v) for v in sh.split("x"))
        if B % G:
            continue
        try:
            s = g.Session(prog, B, group=G, worker_threads=NT, ct_mode=ct_mode, exec_mode=1)
        except g.GsvError as e:
            print(f"  B={B} {sh}: {e}", flush=True)
            continue
        best = None
        for _ in range(reps):
            sys.stderr.flush()
            r = s.garble(list(range(B)), g.HASH_AES, want_inputs=False, want_outputs=False)
            best = r if best is None or r.ms_garble < best.ms_garble else best
        s.close()
        print(f"  B={B} {sh}: garble {best.ms_garble:.2f} ms = {prog.n_gates * B / best.ms_garble / 1e6:.3f} G gates/s, "
              f"{1e3 * best.ms_garble / max(prog.critical_path_levels, 1):.3f} us per critical-path level", flush=True)


if "pcie" in what:
    import torch
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, dst, src in (("D2H", h, d), ("H2D", d, h)):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            t = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        print(f"pcie {name} 1 GiB pinned: {n / best / 1e9:.1f} GB/s", flush=True)
    # both directions at once (two streams)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"pcie D2H + H2D concurrently: {n / dt / 1e9:.1f} GB/s each direction", flush=True)
    del d, h, h2, d2

if "latency" in what:
    for w in (1, 2, 4, 8, 16):
        print(f"dependent gate hash, {w} warp(s) per SM: AES {g.bench_hash_latency(g.HASH_AES, w):.0f} cycles, "
              f"BLAKE3 {g.bench_hash_latency(g.HASH_BLAKE3, w):.0f} cycles", flush=True)

if "small" in what:
    for circ in kv.get("circuits", "fq12_mul,fq12_inverse").split(","):
        p = g.Program(circ)
        print(f"{circ}: gates {p.n_gates} calls {p.n_calls} critical path {p.critical_path_levels} levels, "
              f"sum of call levels {p.sum_call_levels}", flush=True)
        run(p, 32, kv.get("shapes", "4x256,2x256,1x256,4x128,2x128,1x128,4x512").split(","))
        run(p, 4, ["4x256", "1x256"])

if "verifier" in what:
    t = time.perf_counter()
    p = g.Program("groth16_verify_compressed", max_task_slots=int(kv.get("slots", "0")))
    print(f"verifier: planned in {time.perf_counter() - t:.1f} s, gates {p.n_gates} calls {p.n_calls} "
          f"critical path {p.critical_path_levels} levels, sum of call levels {p.sum_call_levels}", flush=True)
    run(p, int(kv.get("B", "32")), kv.get("verifier_shapes", "4x256").split(","), reps=1)

"""Exploration sweep (not the bench): garble-kernel time vs batch / group / worker shape."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsv_b200 as g

circ = sys.argv[1] if len(sys.argv) > 1 else "fq12_mul"
Bs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["16", "64", "256"])]
shapes = [(2, 256), (1, 256), (2, 128), (1, 128), (4, 256), (2, 512)]
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[3].split(",")]
ct_mode = int(sys.argv[4]) if len(sys.argv) > 4 else g.CT_NONE
slots = int(sys.argv[5]) if len(sys.argv) > 5 else 0
p = g.Program(circ, max_task_slots=slots)
print(f"{circ}: gates={p.n_gates} ct={p.n_ciphertexts} calls={p.n_calls} tasks={p.n_tasks} slots={p.max_task_slots} sum_levels={p.sum_call_levels}")
for B in Bs:
    for G, NT in shapes:
        if B % G: continue
        try:
            s = g.Session(p, B, group=G, worker_threads=NT, ct_mode=ct_mode)
        except g.GsvError as e:
            print(f"B={B} G={G} NT={NT}: {e}"); continue
        seeds = list(range(B))
        best = None
        for it in range(3):
            r = s.garble(seeds, g.HASH_AES, want_inputs=False, want_outputs=False)
            if best is None or r.ms_garble < best.ms_garble: best = r
        gps = p.n_gates * B / (best.ms_garble * 1e-3)
        print(f"B={B:5d} G={G} NT={NT:4d}: garble {best.ms_garble:9.3f} ms  {gps/1e9:7.3f} Ggates/s  commit {best.ms_commit:9.3f} ms seed {best.ms_seed:.3f} ms")
        s.close()

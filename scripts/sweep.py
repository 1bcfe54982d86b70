"""Exploration sweep (not the bench): garble-kernel time vs batch / mode / group / worker shape.

usage: sweep.py CIRCUIT B1,B2,.. SHAPES CT_MODE [MAX_TASK_SLOTS] [MAX_TASK_GATES]
  SHAPES: comma list of `lane` or `GxNT` (levelised mode, G instances per item, NT threads/worker)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsv_b200 as g

circ = sys.argv[1] if len(sys.argv) > 1 else "fq12_mul"
Bs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["16", "64", "256"])]
shapes = (sys.argv[3] if len(sys.argv) > 3 else "2x256,lane").split(",")
ct_mode = int(sys.argv[4]) if len(sys.argv) > 4 else g.CT_NONE
slots = int(sys.argv[5]) if len(sys.argv) > 5 else 0
mgates = int(sys.argv[6]) if len(sys.argv) > 6 else 0
p = g.Program(circ, max_task_slots=slots, max_task_gates=mgates)
print(f"{circ}: gates={p.n_gates} ct={p.n_ciphertexts} calls={p.n_calls} tasks={p.n_tasks} "
      f"slots={p.max_task_slots} sum_levels={p.sum_call_levels}", flush=True)
for B in Bs:
    for sh in shapes:
        try:
            if sh == "lane":
                s = g.Session(p, B, ct_mode=ct_mode, exec_mode=2)
            else:
                G, NT = (int(v) for v in sh.split("x"))
                if B % G:
                    continue
                s = g.Session(p, B, group=G, worker_threads=NT, ct_mode=ct_mode, exec_mode=1)
        except g.GsvError as e:
            print(f"B={B} {sh}: {e}", flush=True)
            continue
        seeds = list(range(B))
        best = None
        for it in range(2):
            r = s.garble(seeds, g.HASH_AES, want_inputs=False, want_outputs=False)
            if best is None or r.ms_garble < best.ms_garble:
                best = r
        gps = p.n_gates * B / (best.ms_garble * 1e-3)
        print(f"B={B:5d} {sh:>6s}: garble {best.ms_garble:9.3f} ms  {gps/1e9:7.3f} Ggates/s", flush=True)
        s.close()

"""Run one garble batch (for ncu captures): one.py CIRCUIT B SHAPE CT_MODE"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsv_b200 as g
circ, B, sh, ct = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
p = g.Program(circ)
if sh == "lane":
    s = g.Session(p, B, ct_mode=ct, exec_mode=2)
else:
    G, NT = (int(v) for v in sh.split("x"))
    s = g.Session(p, B, group=G, worker_threads=NT, ct_mode=ct, exec_mode=1)
for it in range(int(sys.argv[5]) if len(sys.argv) > 5 else 2):
    r = s.garble(list(range(B)), g.HASH_AES, want_inputs=False, want_outputs=False)
    print(f"garble {r.ms_garble:.3f} ms", flush=True)

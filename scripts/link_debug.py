"""Stage-by-stage run of a linked garbler -> evaluator pair (debug aid)."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "25")), exit=True)
import numpy as np
import gsv_b200 as g
mode, B, ring_mb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
circ = sys.argv[4] if len(sys.argv) > 4 else "fq12_mul"
def say(*a):
    print(f"[{time.perf_counter():.2f}]", *a, flush=True)
p = g.Program(circ); say("program", p.n_gates)
two = g.device_count() >= 2
sm = 0 if two else 64
gs = g.Session(p, B, device=0, ct_mode=g.CT_NONE, exec_mode=mode, group=2 if mode == 1 else 0, sm_limit=sm); say("garbler session")
es = g.Session(p, B, device=1 if two else 0, ct_mode=g.CT_NONE, exec_mode=mode, group=2 if mode == 1 else 0, sm_limit=sm); say("evaluator session")
g.link_sessions(gs, es, ring_bytes=ring_mb << 20); say("linked")
seeds = list(range(B)); bits = np.random.default_rng(1).integers(0, 2, (B, p.n_inputs), dtype=np.uint8)
lab = gs.expand_seeds(seeds); say("seeds expanded")
t = time.perf_counter()
gres, ev = g.stream_garble_evaluate(gs, es, seeds, g.HASH_AES, bits); dt = time.perf_counter() - t
say(f"streamed in {dt:.3f} s: {p.n_gates * B / dt / 1e9:.2f} G gates/s, {p.n_ciphertexts * B * 16 / dt / 1e9:.1f} GB/s of ciphertexts", ev.ct_commit[0].tobytes().hex(), ev.output_bits[0][:8])

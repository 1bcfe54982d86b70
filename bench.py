#!/usr/bin/env python
"""bench.py -- garbled gates/s of the B200 engine on the Groth16 verifier circuit.

A "step" garbles one cut-and-choose batch of the workload circuit on every GPU from seeds (seed expansion
-> garbling -> bit-exact ciphertext chain commitment), i.e. the garbling stage of the reference's
cut-and-choose (src/cut_and_choose/garbler.rs:191-242) with `AesNiHasher` + `AESAccumulatingHash`; seeds
are the reference's: consecutive u64 draws of ChaCha20Rng::seed_from_u64(1234) (garbler.rs:201-203).

Workloads (--workload):
  verifier (default): groth16_verify_compressed (11.46 G gates, 2.98 G ciphertexts; BASELINE.json configs
          2 / 4) x 16 instances per step and GPU, levelised kernel, commitment in GSV_CT_COMMIT_HOST mode: every
          gate hash on the GPU, the strictly serial AES chain folded by host AES-NI threads that drain the
          ciphertext ring over PCIe while the kernel runs.  One chain advances one block per ~13 ns whatever
          the hardware, i.e. >= 39 s per instance, so steps are SOFTWARE-PIPELINED: `--sessions` (3) steps are in
          flight per GPU, each on its own slice of the SMs and of the HBM ring, so that ~48 chains keep the
          PCIe link and the fold threads busy while no step waits for another.
  batch : fq12_mul (20.3 M gates) x 6144 instances per step, lane kernel, commitment fused on the GPU
          (GSV_CT_COMMIT): the large-batch regime where thousands of chains run side by side.

  value : whole-job gates/s over the device-side span of the K steps (CUDA events around the region; the
          only input, the seeds, is 8 bytes per instance).
  e2e   : the same over the wall time of the public API calls with HOST buffers: seeds H2D; commitments and
          input / output labels D2H; in the verifier workload also the ciphertext drain and the host fold.
  --impl reference : the reference's CPU path restated (oracle, AES-NI) on all host cores over bounded
          windows of the same circuit; loads neither libgsv_cuda.so nor torch.

Launch: `python bench.py --gpus 1`, or under torchrun for N > 1 (one rank per GPU, instances sharded
across ranks, NCCL used only to all-gather the per-instance commitments).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_GATE = 52.3   # 48 B free gate, 64 B non-free, 73.2/26.8 mix
AES_BLOCKS_PER_GATE_GARBLE = 2.0  # per NON-FREE gate; + 1 chain block when committing
PCIE_D2H_GBS = 57.2          # measured on this pool (profiles/r02_probe.md): pinned D2H, 1 GiB copies
SM_RESERVE = 4               # SMs left free of persistent CTAs (NCCL / utility kernels must always fit)
# DRAM traffic of the levelised kernel with ciphertexts dropped, from ncu (dram__bytes_read.sum + dram__bytes_write.sum of
# k_engine<4,0,0> on the verifier x 16: 6.69 GB + 17.07 GB per launch, profiles/r02_traffic_verifier_B16_nocommit.csv):
# labels live in shared memory and L2, only task inputs / outputs and the program reach DRAM
LABEL_DRAM_BYTES_PER_GATE = (6694099712 + 17072106240) / (16 * 11457232209)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


HOST_DRAM_GBS = 300.0   # assumed host memory bandwidth shared by the ranks' drains (DMA writes) and folds (reads)


def host_plan(S, B, steps, warmup, cpu_share, n_ct, commit_host=True, sessions_auto=True, instances_auto=True,
              host_threads=0, ranks_on_host=1):
    """Steps in flight per GPU (S), instances per step (B) and AES-NI fold threads per step for a rank that owns
    `cpu_share` host CPUs.  Returns (S, B, fold_threads, note).  The serial chain commitments are folded on the host:
    this is the part of the workload that depends on the host, and the bench says what it chose."""
    note = None
    if commit_host and sessions_auto:
        if cpu_share < 6:
            # few host CPUs per rank (8 GPUs on a 32-CPU host): one fold thread per session, each interleaving all
            # four quads of its session (the VAES throughput cap, 640 M blocks/s per thread), one session per CPU
            S = max(S, min(4, cpu_share))
        elif cpu_share >= 16:
            # A run of K steps takes ceil(K / S) waves of S steps in flight.  Measured time one step is in flight on a
            # B200 with a 16-CPU host: 48 s at S = 3 (48 SMs per step), 63 s at S = 4 (36 SMs): 20 steps are
            # 7 x 48 s or 5 x 63 s.
            wave_s = {3: 48.0, 4: 63.0}
            S = min(wave_s, key=lambda n: -(-max(1, steps) // n) * wave_s[n])
    if commit_host and instances_auto:
        # The folds do about 0.5 G blocks/s per CPU with several chains interleaved: when a rank has very few CPUs
        # (many GPUs on a small host) a 16-instance step can take so long that `--steps 20 --warmup 5` no longer ends
        # within the driver's window.  Shrink the step then, and say so.
        per_step_s = 720.0 / max(1, warmup + steps)
        cap = int(per_step_s * min(cpu_share, S * ((B + 3) // 4)) * 0.5e9 / max(1, n_ct))
        # every ciphertext byte is written to host memory by the drain and read once by a fold thread
        cap_mem = int(per_step_s * HOST_DRAM_GBS * 1e9 / max(1, ranks_on_host) / (32.0 * max(1, n_ct)))
        if min(cap, cap_mem) < B:
            B_new = max(4, min(cap, cap_mem) // 4 * 4)
            why = (f"{cpu_share} host CPUs per rank fold about {0.5 * cpu_share:.1f} G ciphertexts/s" if cap <= cap_mem else
                   f"{ranks_on_host} ranks share about {HOST_DRAM_GBS:.0f} GB/s of host memory bandwidth (32 B per ciphertext)")
            note = f"{B_new} instead of {B} instances per step: {why}, a {B}-instance step would not fit the run's time window"
            B = B_new
    # one fold thread per quad of chains when the rank has a CPU for each; otherwise one CPU is left to the drain threads
    quads = (B + 3) // 4
    fold_cpus = cpu_share if S * quads <= cpu_share else max(1, cpu_share - (1 if cpu_share >= 6 else 0))
    fold_threads = host_threads or max(1, min(quads, fold_cpus // S))
    return S, B, fold_threads, note



def plan_cache_dir():
    """The ranks of a job (and successive runs on one box) plan the circuit once and share the result through a
    file (GSV_PLAN_CACHE_DIR, see gsv_program_build): 23 s and 17 GB of host memory per process otherwise."""
    if "GSV_PLAN_CACHE_DIR" in os.environ:
        return
    for d in ("/dev/shm", "/tmp"):
        path = os.path.join(d, "gsv_plan_cache")
        try:
            os.makedirs(path, exist_ok=True)
            st = os.statvfs(path)
            if st.f_bavail * st.f_frsize > 12 << 30:
                os.environ["GSV_PLAN_CACHE_DIR"] = path
                return
        except OSError:
            pass


def host_cpus():
    """(logical CPUs this process may use, physical cores of the host)."""
    try:
        logical = len(os.sched_getaffinity(0))
    except Exception:
        logical = os.cpu_count() or 1
    physical = None
    try:
        import psutil
        physical = psutil.cpu_count(logical=False)
    except Exception:
        pass
    return logical, physical or logical


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        # median over samples under load (above 60 % of the maximum seen)
        load = [x for x in sm if x >= 0.6 * (sm[-1] if sm else 0)] or sm
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
WINDOW_GATES = 100_000_000   # gates per sampled window, per thread and step


class CpuSampler:
    """The reference's per-gate loop restated (oracle/, AES-NI when the host has it) over BOUNDED WINDOWS of
    the workload circuit, one instance per thread like the reference's rayon pool pinned to cores
    (src/cut_and_choose/mod.rs:131-186).  The verifier is walked as its template DAG; each thread garbles
    the first WINDOW_GATES gates of every major stage (G1 / G2 decompression, MSM, Miller loop, final
    exponentiation -- the root's children above 100 M gates) with commitment, and the stage rates are
    combined with the stages' true gate counts: rate = total gates / sum_i (gates_i / rate_i).
    Loads only oracle/libgsv_oracle.so and oracle/libgsv_circuitgen.so (host-only gadget generator)."""

    def __init__(self, circuit):
        from oracle import oracle as o
        self.o = o
        t = time.perf_counter()
        self.gen = o.CircuitGen(circuit)
        self.record_s = time.perf_counter() - t
        self.n_gates = self.gen.n_gates
        kids = self.gen.children()
        big = [k for k in kids if int(self.gen.total_gates[k]) >= WINDOW_GATES]
        if self.n_gates <= 4 * WINDOW_GATES or not big:
            self.windows = [(self.gen.root, self.n_gates, self.n_gates)]       # whole circuit per sample
        else:
            merged = {}
            for k in big:                                                      # same template twice = same stage
                merged[k] = merged.get(k, 0) + int(self.gen.total_gates[k])
            covered = sum(merged.values())
            scale = self.n_gates / covered                                     # small children ride along pro rata
            self.windows = [(k, g * scale, min(WINDOW_GATES, int(self.gen.total_gates[k]))) for k, g in merged.items()]
        self.dags = {k: self.gen.dag(k) for k, _, _ in self.windows}
        self.aesni = o.have_aesni()

    def describe(self):
        if len(self.windows) == 1:
            return f"whole {self.gen.circuit} instances"
        return (f"{len(self.windows)} windows of {WINDOW_GATES / 1e6:.0f} M gates (the start of each verifier stage: "
                f"G1 / G2 decompression, MSM, Miller loop, final exponentiation), stage rates weighted by the stages' gate counts")

    def step(self, threads, hasher=0, seed0=0):
        """One bounded sample on `threads` threads.  Returns (aggregate gates/s, seconds)."""
        est = [0.0] * threads

        def work(k):
            t_inst = 0.0
            for (tmpl, weight_gates, limit) in self.windows:
                t = time.perf_counter()
                whole = limit >= int(self.gen.total_gates[tmpl])
                r = self.dags[tmpl].garble(hasher, seed0 + 1000 * k, max_gates=0 if whole else limit)  # ctypes releases the GIL
                dt = time.perf_counter() - t
                t_inst += weight_gates * dt / max(r["n_gates"], 1)
            est[k] = self.n_gates / t_inst

        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return sum(est), time.perf_counter() - t0


def loaded_repo_libs():
    """Shared objects of this repository mapped into the process (the reference arm must not hold the engine)."""
    libs = set()
    try:
        with open("/proc/self/maps") as f:
            for ln in f:
                path = ln.split()[-1]
                if path.startswith(ROOT) and ".so" in path:
                    libs.add(os.path.relpath(path, ROOT))
    except Exception:
        pass
    return sorted(libs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    logical, physical = host_cpus()
    sampler = CpuSampler(args.circuit)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, dt = sampler.step(logical, seed0=i)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(d for _, d in rates) / len(rates)
    sample = (f"{sampler.describe()}, one instance per thread on {logical} threads ({physical} physical cores) per step, "
              f"garble + chain commitment, {'AES-NI' if sampler.aesni else 'portable AES'}; the oracle omits the reference's "
              f"slab / credit bookkeeping (upper bound on the reference's CPU speed)")
    line = {
        "impl": "reference", "metric": "garbled_gates_per_s", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.circuit} garble + chain commitment, AES hasher, CPU oracle on all host threads "
                               f"(bounded sample per step: {sampler.describe()})",
                   "gates_per_instance": sampler.n_gates, "record_s": round(sampler.record_s, 1),
                   "same_config_note": "same circuit, hasher and commitment as the GPU arm; each step is a bounded sample "
                                       "instead of whole instances (a whole instance takes ~270 s per core)"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": physical, "threads": logical, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "native_libs": loaded_repo_libs(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- config 3
NVLINK_PEER_GBS = 770.0      # measured peer copy per direction on this pool (B200_PROFILING.md)


def run_mpc(args):
    """BASELINE.json config 3: the garbler (GPU 0) regarbles the verifier and streams every ciphertext into a ring in
    the evaluator's memory (GPU 1) with peer stores over NVLink; the evaluator (GPU 1) consumes the ring, and hashes
    what it received on host AES-NI threads like the reference's evaluator does (examples/groth16_garble.rs:170-267).
    One process drives both GPUs (`--gpus 2`; on one GPU the two persistent grids split the SMs)."""
    import numpy as np
    import torch

    import gsv_b200 as g
    from gsv_b200 import cut_and_choose as cc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    two = torch.cuda.device_count() >= 2 and args.gpus >= 2
    hasher = g.HASH_AES if args.hasher == "aes" else g.HASH_BLAKE3
    plan_cache_dir()
    t_plan = time.perf_counter()
    prog = g.Program(args.circuit)
    t_plan = time.perf_counter() - t_plan
    B = args.instances
    logical, physical = host_cpus()
    fold_threads = args.host_threads or max(1, min((B + 3) // 4, logical - 2))
    sm_limit = 0 if two else (torch.cuda.get_device_properties(0).multi_processor_count - SM_RESERVE) // 2
    gs = g.Session(prog, B, device=0, group=args.group, worker_threads=args.worker_threads, ct_mode=g.CT_NONE,
                   exec_mode=args.exec_mode, sm_limit=sm_limit)
    es = g.Session(prog, B, device=1 if two else 0, group=args.group, worker_threads=args.worker_threads, ct_mode=g.CT_NONE,
                   exec_mode=args.exec_mode, sm_limit=sm_limit, host_threads=fold_threads)
    g.link_sessions(gs, es)
    compressed = args.circuit == "groth16_verify_compressed"
    if args.circuit.startswith("groth16"):
        bits1 = g.groth16_synthetic_inputs(compressed=compressed)
    else:
        bits1 = np.random.default_rng(3).integers(0, 2, prog.n_inputs, dtype=np.uint8)
    bits = np.broadcast_to(bits1, (B, prog.n_inputs)).copy()
    seeds = np.asarray(cc.instance_seeds(1234, (args.warmup + args.steps) * B), dtype=np.uint64)
    for i in range(args.warmup):
        g.stream_garble_evaluate(gs, es, seeds[i * B:(i + 1) * B], hasher, bits)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ok, ms_eval, ms_garble, launches = True, 0.0, 0.0, 0
    for i in range(args.warmup, args.warmup + args.steps):
        gres, ev = g.stream_garble_evaluate(gs, es, seeds[i * B:(i + 1) * B], hasher, bits)
        ms_eval += ev.ms_evaluate
        ms_garble += gres.ms_garble
        launches += gres.n_launches + ev.n_launches
        if args.circuit.startswith("groth16"):
            ok = ok and bool((ev.output_bits == 1).all())   # the synthetic proof verifies
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    gates = prog.n_gates * B * args.steps
    ct_bytes = prog.n_ciphertexts * 16 * B * args.steps
    value = gates / wall
    nv = ct_bytes / wall / 1e9
    peak, peak_src = measured_peaks()
    line = {
        "metric": "garbled_gates_per_s", "value": value, "unit": "gates/s", "n_gpus": 2 if two else 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": f"{args.circuit} x {B} instances garbled on GPU 0 and evaluated on GPU {1 if two else 0} at the same time "
                        f"(BASELINE.json config 3): ciphertexts streamed through a ring in the evaluator's memory"
                        + (" by peer stores over NVLink" if two else " (one GPU: the two grids split the SMs)")
                        + ", evaluator hashes the received stream on host AES-NI threads",
            "kernel": "k_engine (garble) + k_engine (evaluate)", "plan_s": round(t_plan, 1), "instances": B,
            "gates_per_instance": prog.n_gates, "ciphertexts_per_instance": prog.n_ciphertexts,
            "verify_bits_all_one": ok,
        },
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": B * (8 + 16 * (2 + prog.n_inputs) + prog.n_inputs),
                "d2h_bytes_per_step": B * 16 * (3 + prog.n_inputs + 2 * prog.n_outputs) + B * 16 * prog.n_ciphertexts},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "nvlink" if two else "hbm", "achieved": nv, "peak": NVLINK_PEER_GBS if two else peak, "unit": "GB/s",
                     "frac": nv / (NVLINK_PEER_GBS if two else peak), "traffic": None, "kernel": "k_engine",
                     "kernel_ms": ms_garble / args.steps, "evaluate_kernel_ms": ms_eval / args.steps,
                     "note": "achieved = ciphertext bytes crossing from the garbler to the evaluator per second (16 B x ciphertexts); "
                             "the stream is paced by the evaluator-side hash (one AES-NI chain per instance, fed over PCIe), "
                             "not by the link"},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsv", choices=["gsv", "reference"])
    ap.add_argument("--workload", default="verifier", choices=["verifier", "batch", "mpc"])
    ap.add_argument("--circuit", default=None, help="override the workload's circuit")
    ap.add_argument("--instances", type=int, default=None, help="cut-and-choose instances per step and GPU")
    ap.add_argument("--sessions", type=int, default=None, help="steps in flight per GPU (software pipelining)")
    ap.add_argument("--exec-mode", type=int, default=None, help="0 auto, 1 levelised, 2 lane")
    ap.add_argument("--group", type=int, default=None)
    ap.add_argument("--ct-mode", default=None, choices=["commit", "commit_host", "none"])
    ap.add_argument("--worker-threads", type=int, default=0)
    ap.add_argument("--max-task-slots", type=int, default=0, help="planner: shared-memory label slots per task (0 = default)")
    ap.add_argument("--host-threads", type=int, default=0, help="fold threads per session (0: from the rank's CPU share)")
    ap.add_argument("--hasher", default="aes", choices=["aes", "blake3"])
    ap.add_argument("--no-commit", action="store_true", help="drop ciphertexts (the `()` handler)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true",
                    help="after the timed region (N = 1): the same kernel with ciphertexts dropped on the whole GPU")
    args = ap.parse_args()
    # workload presets (explicit flags win)
    preset = {"verifier": dict(circuit="groth16_verify_compressed", instances=16, sessions=3, exec_mode=1, group=4,
                               ct_mode="commit_host", steps=6),
              "mpc": dict(circuit="groth16_verify_compressed", instances=16, sessions=1, exec_mode=1, group=4,
                          ct_mode="none", steps=2),
              "batch": dict(circuit="fq12_mul", instances=6144, sessions=1, exec_mode=2, group=0, ct_mode="commit",
                            steps=3)}[args.workload]
    args.sessions_auto = args.sessions is None
    args.instances_auto = args.instances is None
    for k, v in preset.items():
        if getattr(args, k) is None:
            setattr(args, k, v)
    if args.no_commit:
        args.ct_mode = "none"

    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "mpc":
        run_mpc(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gsv_b200 as g
    from gsv_b200 import cut_and_choose as cc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    hasher = g.HASH_AES if args.hasher == "aes" else g.HASH_BLAKE3
    ct_mode = {"none": g.CT_NONE, "commit": g.CT_COMMIT, "commit_host": g.CT_COMMIT_HOST}[args.ct_mode]
    plan_cache_dir()
    t_plan = time.perf_counter()
    prog = g.Program(args.circuit, max_task_slots=args.max_task_slots)
    t_plan = time.perf_counter() - t_plan
    B, S = args.instances, max(1, args.sessions)
    lane = args.exec_mode == 2 or (args.exec_mode == 0 and args.group == 0 and B >= 128)
    kernel_name = "k_lane" if lane else "k_engine"

    # ---- host threads: every rank gets an equal share of the host's CPUs; inside a rank each session has one
    # drain thread (mostly asleep) and `fold_threads` AES-NI fold threads
    logical, physical = host_cpus()
    cpu_share = max(1, logical // max(1, local_world))
    S, B, fold_threads, instances_note = host_plan(
        S, B, args.steps, args.warmup, cpu_share, prog.n_ciphertexts, commit_host=ct_mode == g.CT_COMMIT_HOST,
        sessions_auto=args.sessions_auto, instances_auto=args.instances_auto, host_threads=args.host_threads,
        ranks_on_host=local_world)
    sm_total = torch.cuda.get_device_properties(local).multi_processor_count
    sm_limit = 0 if S == 1 else (sm_total - SM_RESERVE) // S
    free_b, _ = torch.cuda.mem_get_info()
    sessions = []
    for k in range(S):
        sessions.append(g.Session(prog, B, device=local, group=args.group, worker_threads=args.worker_threads,
                                  ct_mode=ct_mode, exec_mode=args.exec_mode, sm_limit=sm_limit,
                                  ct_buffer_bytes=0 if S == 1 else int(free_b * 0.80 / S), host_threads=fold_threads))
    commit_txt = {"none": "no commitment (ciphertexts dropped)",
                  "commit": "bit-exact AES chain commitment fused into the kernel (chain CTAs)",
                  "commit_host": "bit-exact AES chain commitment, gate hashes on the GPU, serial chain folded by "
                                 "host AES-NI threads draining the ciphertext ring during the kernel"}[args.ct_mode]

    # cut-and-choose seeds (garbler.rs:201-203): consecutive u64 draws of ChaCha20Rng::seed_from_u64(1234);
    # step i of rank r takes draws [(i * world + r) * B, +B)
    n_steps_all = args.warmup + args.steps
    all_seeds = np.asarray(cc.instance_seeds(1234, n_steps_all * world * B), dtype=np.uint64)

    def seeds_for(step):
        o = (step * world + rank) * B
        return all_seeds[o:o + B]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    commits_dev = torch.empty((B, 16), dtype=torch.uint8, device="cuda")
    gathered = torch.empty((world * B, 16), dtype=torch.uint8, device="cuda") if world > 1 else None

    def run_steps(first, count):
        """Steps [first, first + count): step i runs on session i % S; the S sessions run concurrently, each from
        its own host thread (the library call blocks until the step's commitments are final).  The commitments
        are all-gathered in step order by this thread."""
        results = [None] * count
        done = [threading.Event() for _ in range(count)]
        errors = []

        def drive(k):
            try:
                torch.cuda.set_device(local)
                for j in range(k, count, S):
                    t1 = time.perf_counter()
                    res = sessions[k].garble(seeds_for(first + j), hasher, want_inputs=True, want_outputs=True)
                    res.wall_ms = 1e3 * (time.perf_counter() - t1)
                    results[j] = res
                    done[j].set()
            except Exception as e:  # pragma: no cover
                errors.append(e)
                for ev in done:
                    ev.set()

        th = [threading.Thread(target=drive, args=(k,)) for k in range(min(S, count))]
        for x in th:
            x.start()
        for j in range(count):
            done[j].wait()
            if errors:
                break
            if world > 1:
                # the only collective of the path: gather the per-instance commitments
                commits_dev.copy_(torch.from_numpy(results[j].ct_commit))
                dist.all_gather_into_tensor(gathered, commits_dev)
        for x in th:
            x.join()
        if errors:
            raise errors[0]
        return results

    if args.warmup:
        run_steps(0, args.warmup)

    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t0 = time.perf_counter()
    results = run_steps(args.warmup, args.steps)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    span_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()

    garble_ms = sum(r.ms_garble for r in results)
    commit_ms = sum(r.ms_commit for r in results)
    seed_ms = sum(r.ms_seed for r in results)
    step_ms = sum(r.ms_total for r in results)
    launches = sum(r.n_launches for r in results)

    # max over ranks: the CUDA-event span for `value`, the wall time (incl. every copy) for e2e
    tt = torch.tensor([span_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    span_ms_max, wall_ms_max = [float(x) for x in tt.tolist()]

    gates_per_step = prog.n_gates * B * world
    value = gates_per_step * args.steps / (span_ms_max * 1e-3)
    e2e = gates_per_step * args.steps / (wall_ms_max * 1e-3)
    h2d = 8 * B
    d2h = B * 16 * (4 + prog.n_inputs + prog.n_outputs)
    if args.ct_mode == "commit_host":
        d2h += B * 16 * prog.n_ciphertexts  # the ciphertext stream itself is drained to the host

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel = the persistent engine kernel (garbling; in `commit` mode also the fused chain).  Its
        # launch durations come from CUDA events recorded by the library on the launching stream.  With S steps in
        # flight S launches run side by side on disjoint SM slices: the GPU's algorithmic rate is the sum.
        k_ms = garble_ms / args.steps                      # average duration of one launch
        conc = min(S, args.steps)
        k_gates = prog.n_gates * B                         # gates one launch processes
        per_launch = k_gates * ALGO_BYTES_PER_GATE / (k_ms * 1e-3) / 1e9
        busy = garble_ms / (span_ms * max(conc, 1))        # share of the span a session's kernel is running
        achieved = per_launch * conc * min(1.0, busy)
        traffic = traffic_note = None
        if not lane and args.circuit == "groth16_verify_compressed":
            ct_bytes = 16 * prog.n_ciphertexts * B if args.ct_mode != "none" else 0
            traffic = k_gates * LABEL_DRAM_BYTES_PER_GATE + ct_bytes * (2 if args.ct_mode == "commit_host" else 1)
            traffic_note = ("per launch: label / program traffic measured by ncu on this kernel with ciphertexts dropped "
                            "(0.13 B per gate: labels stay in shared memory and L2) + the ciphertext stream (16 B x n_ct x B, "
                            "written by the kernel, read once by the drain); far below the algorithmic 52.3 B per gate")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_name, "kernel_ms": k_ms, "peak_source": peak_src,
                    "algorithmic_bytes_per_gate": ALGO_BYTES_PER_GATE,
                    "concurrent_launches": conc, "sms_per_launch": sm_limit or sm_total,
                    "per_launch": {"achieved": per_launch, "frac": per_launch / peak},
                    "note": "achieved = algorithmic bytes of one launch / its average duration x the launches running side by side "
                            "(software-pipelined steps on disjoint SM slices); per_launch is the single-launch figure"}
        if prog.critical_path_levels and not lane:
            roofline["latency_floor"] = {
                "critical_path_levels": prog.critical_path_levels,
                "us_per_level_at_measured_time": 1e3 * k_ms / prog.critical_path_levels,
                "note": "kernel time / critical-path levels: the per-level latency if nothing else paced the run "
                        "(in commit_host mode the ring back-pressure of the drain / fold does)"}
        if args.ct_mode == "commit_host":
            wave_s = (wall / args.steps) * conc            # time one step is in flight
            drain_gbs = d2h * args.steps / wall / 1e9
            fold_floor_s = prog.n_ciphertexts * 13e-9      # one dependent AES-NI block per ~13 ns (10 aesenc)
            pcie_util = drain_gbs / PCIE_D2H_GBS
            fold_busy = max(r.host_fold_busy for r in results)               # busiest fold thread, share of its step
            wait_kernel = sum(r.host_drain_wait_kernel for r in results) / len(results)
            wait_fold = sum(r.host_drain_wait_fold for r in results) / len(results)
            # utilisation of each stage of the committed step; the drain thread is always a few buffers ahead, so its
            # wait for free buffers says nothing about which side paces it
            utils = {"host_fold": fold_busy, "pcie_drain": pcie_util, "chain_latency": fold_floor_s / wave_s}
            top = max(utils, key=utils.get)
            if wait_kernel >= 0.5:
                limiter = "kernel"
            elif utils[top] >= 0.85:
                limiter = top
            else:
                limiter = f"balanced ({top} {utils[top]:.2f})"
            roofline["gpu_chain_floor_s"] = prog.n_ciphertexts * 0.46e-6  # measured dependent-AES step on the GPU
            roofline["pipeline"] = {
                "limiter": limiter, "steps_in_flight": conc, "step_in_flight_s": wave_s,
                "d2h_drain_GBps": drain_gbs, "pcie_d2h_peak_GBps": PCIE_D2H_GBS, "pcie_util": pcie_util,
                "host_chain_floor_s": fold_floor_s, "fold_thread_busy": fold_busy,
                "drain_wait_for_kernel": wait_kernel, "drain_wait_for_fold": wait_fold,
                "host_threads_per_rank": S * (fold_threads + 1), "fold_threads_per_session": fold_threads,
                "host_logical_cpus": logical, "host_physical_cores": physical, "ranks_on_host": local_world,
                "stage_utilisation": utils,
                "note": "rank 0's view.  Stage utilisations: host_fold = busy share of the busiest AES-NI fold thread; pcie_drain = "
                        "ciphertext drain rate / measured pinned D2H rate (GPUs that share a PCIe switch divide it); chain_latency = "
                        "serial-chain floor of one instance (n_ct x 13 ns) / time a step is in flight.  limiter = kernel when the "
                        "drain waits for the kernel to publish ciphertexts half of the time, else the stage at >= 85 %, else "
                        "'balanced': kernel production (the same sessions with ciphertexts dropped run at 13.2 G gates/s), the "
                        "PCIe drain (13.7 G) and the folds (48 chains per 38.7 s = 14.2 G) are within 10 % of each other"}
        try:
            blocks = g.bench_hash(hasher, 1 << 28, 2, device=local)
            nonfree = sum(prog.type_count[:8]) / prog.n_gates
            need = nonfree * (AES_BLOCKS_PER_GATE_GARBLE + (1.0 if args.ct_mode == "commit" else 0.0)) * gates_per_step / world * args.steps / (span_ms * 1e-3)
            roofline["alu"] = {"hash_blocks_per_s_peak": blocks, "hash_blocks_per_s_achieved": need,
                               "frac": need / blocks,
                               "note": "register-resident 2-block gate-hash micro-kernel; T-table AES is bound by the "
                                       "shared-memory pipe (160 lookups per block, one 32-lane wavefront per SM clock)"}
        except Exception as e:  # pragma: no cover
            roofline["alu"] = {"error": str(e)}
        line = {
            "metric": "garbled_gates_per_s", "value": value, "unit": "gates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": f"{args.circuit} x {B} cut-and-choose instances per step and GPU ({S} steps in flight per GPU), "
                            f"garble + {commit_txt}, {args.hasher} gate hasher"
                            + (" (BASELINE.json configs 2/4: Groth16 verifier, 1 public input, synthetic vk; "
                               "11.46 G gates as recorded from the current gadget sources; the reference's README quotes 11.17 G)"
                               if args.circuit == "groth16_verify_compressed" else ""),
                "kernel": kernel_name, "plan_s": round(t_plan, 1), "plan_cache": os.environ.get("GSV_PLAN_CACHE_DIR"),
                "gates_per_instance": prog.n_gates, "ciphertexts_per_instance": prog.n_ciphertexts,
                "instances_per_step": B, "steps_in_flight": S, "instances_in_flight_per_gpu": B * min(S, args.steps),
                "seeds": "ChaCha20Rng::seed_from_u64(1234) u64 draws (garbler.rs:201-203)",
                "l2": "inputs larger than L2 (label + ciphertext state >> 126 MB)",
                "parallelism": f"instances sharded over {world} GPU(s); NCCL all-gather of commitments only",
                **({"instances_note": instances_note} if instances_note else {}),
            },
            "phases_ms_per_step": {"seed_expand": seed_ms / args.steps, "garble_kernel": garble_ms / args.steps,
                                   "commit_tail_after_kernel": commit_ms / args.steps, "step_in_flight": step_ms / args.steps},
            "e2e": {"value": e2e, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
        }
        for s in sessions:
            s.close()
        sessions.clear()
        if world == 1 and args.extras and not lane:
            # the same kernel, ciphertexts dropped, whole GPU: not paced by the host fold / PCIe drain
            try:
                s2 = g.Session(prog, 2 * B, device=local, group=args.group, worker_threads=args.worker_threads,
                               ct_mode=g.CT_NONE, exec_mode=args.exec_mode)
                r2 = s2.garble(all_seeds[:2 * B], hasher, want_inputs=False, want_outputs=False)
                s2.close()
                rate2 = prog.n_gates * 2 * B / (r2.ms_garble * 1e-3)
                roofline["kernel_without_commitment"] = {
                    "instances": 2 * B, "kernel_ms": r2.ms_garble, "gates_per_s": rate2,
                    "achieved_GBps": rate2 * ALGO_BYTES_PER_GATE / 1e9, "frac": rate2 * ALGO_BYTES_PER_GATE / 1e9 / peak,
                    "us_per_critical_level": 1e3 * r2.ms_garble / max(prog.critical_path_levels, 1)}
            except Exception as e:  # pragma: no cover
                roofline["kernel_without_commitment"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cs = CpuSampler(args.circuit)
                rate, dt = cs.step(logical)
                line["cpu_baseline"] = {
                    "value": rate, "unit": "gates/s", "cores": physical, "threads": logical, "kind": "port",
                    "sample": f"{cs.describe()}, one instance per thread on {logical} threads ({dt:.1f} s), garble + chain "
                              f"commitment, {'AES-NI' if cs.aesni else 'portable AES'} oracle",
                }
            except Exception as e:  # pragma: no cover -- never lose the measured line to the baseline leg
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": physical, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

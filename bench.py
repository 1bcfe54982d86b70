#!/usr/bin/env python
"""bench.py -- garbled gates/s of the B200 engine on the Groth16 verifier circuit.

A "step" garbles one batch of B cut-and-choose instances of the workload circuit per GPU from
seeds (seed expansion -> garbling -> bit-exact ciphertext chain commitment), i.e. the first
garbling stage of the reference's cut-and-choose (src/cut_and_choose/garbler.rs:191-242) with
`AesNiHasher` + `AESAccumulatingHash`.

Workloads (--workload):
  verifier (default) : groth16_verify_compressed (11.46 G gates, 2.98 G ciphertexts; BASELINE.json
          configs 2/4) x 32 instances per GPU, levelised kernel (4 instances per worker), commitment in GSV_CT_COMMIT_HOST mode:
          every gate hash on the GPU, the strictly serial AES chain folded by host AES-NI threads that
          drain the ciphertext ring while the kernel runs (a GPU folds one dependent AES per 0.46 us:
          23 min for 2.98 G ciphertexts, whatever the batch).
  batch   : fq12_mul (20.3 M gates) x 6144 instances per GPU, lane kernel, commitment fused on the GPU
          (GSV_CT_COMMIT): the large-batch regime where thousands of chains run side by side.

  value : whole-job gates/s over the device time of the step (CUDA events inside the library, on its
          stream), seeds already resident in HBM being the only input.
  e2e   : same metric over the wall time of the public API call with HOST buffers (seeds H2D; commitments,
          input / output labels D2H; in the verifier workload also the ciphertext drain and the host
          fold, which end a few seconds after the kernel) -- the headline.
  --impl reference : the CPU oracle (AES-NI restatement of the reference's per-gate loop; the
          reference itself is Rust and cannot be built in this image) on all host cores.

Launch: `python bench.py --gpus 1`, or under torchrun for N > 1 (one rank per GPU, instances
sharded across ranks, NCCL used only to all-gather the per-instance commitments).
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


VERIFIER_SAMPLE_GATES = 600_000_000   # per core and step: the first 5 % of the verifier's emission order


def cpu_garble_rate(circuit, n_instances_per_core, cores, hasher=0, prog=None):
    """Oracle (CPU restatement, AES-NI when the host has it) on `cores` threads, one instance
    at a time per core like the reference's pinned rayon pool (cut_and_choose/mod.rs:131-186).
    Circuits that fit memory as a flat stream run `n_instances_per_core` whole instances per core; the
    Groth16 verifier is walked over its template DAG and each core garbles the first
    VERIFIER_SAMPLE_GATES gates of its own instance (a bounded sample of the same workload)."""
    import gsv_b200 as g
    from oracle import oracle as o

    if circuit == "groth16_verify_compressed":
        prog = prog or g.Program(circuit, lane_only=True)
        dag = o.TemplateDag(*prog.export_templates())
        done = [0] * cores

        def work(k):
            done[k] = dag.garble(hasher, 1000 * k, max_gates=VERIFIER_SAMPLE_GATES)["n_gates"]  # ctypes releases the GIL
        sample = f"first {VERIFIER_SAMPLE_GATES / 1e6:.0f} M gates of one {circuit} instance per core"
    else:
        prog = prog or g.Program(circuit)
        t, a, b, c, outs, nw = prog.flat_stream()
        st = o.Stream(t, a, b, c, outs, nw, prog.n_inputs).compact()  # slab-sized live set, cache resident
        done = [prog.n_gates * n_instances_per_core] * cores

        def work(k):
            for j in range(n_instances_per_core):
                st.garble(hasher, 1000 * k + j, want_ct=False)
        sample = f"{n_instances_per_core} instance(s) of {circuit} per core"

    th = [threading.Thread(target=work, args=(k,)) for k in range(cores)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt, prog, o.have_aesni(), sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rates = []
    per_core = max(1, args.ref_instances_per_core)
    for i in range(args.warmup + args.steps):
        rate, dt, prog, aesni, what = cpu_garble_rate(args.cpu_circuit, per_core, cores)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(d for _, d in rates) / len(rates)
    sample = (f"{what} on {cores} threads per step, garble + chain "
              f"commitment, {'AES-NI' if aesni else 'portable AES'}; oracle omits the reference's slab/credit "
              f"bookkeeping (upper bound on the reference's CPU speed)")
    line = {
        "impl": "reference", "metric": "garbled_gates_per_s", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.circuit} garble+commit, AES hasher (CPU oracle; bounded sample: {what}, "
                               f"independent per core)",
                   "gates_per_instance": prog.n_gates},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsv", choices=["gsv", "reference"])
    ap.add_argument("--workload", default="verifier", choices=["verifier", "batch"])
    ap.add_argument("--circuit", default=None, help="override the workload's circuit")
    ap.add_argument("--instances", type=int, default=None, help="cut-and-choose instances per GPU")
    ap.add_argument("--exec-mode", type=int, default=None, help="0 auto, 1 levelised, 2 lane")
    ap.add_argument("--group", type=int, default=None)
    ap.add_argument("--ct-mode", default=None, choices=["commit", "commit_host", "none"])
    ap.add_argument("--worker-threads", type=int, default=0)
    ap.add_argument("--hasher", default="aes", choices=["aes", "blake3"])
    ap.add_argument("--no-commit", action="store_true", help="drop ciphertexts (the `()` handler)")
    ap.add_argument("--ref-instances-per-core", type=int, default=8)
    ap.add_argument("--cpu-baseline-instances", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unthrottled", action="store_true", help="skip the extra no-commitment kernel step")
    args = ap.parse_args()
    # workload presets (explicit flags win)
    preset = {"verifier": dict(circuit="groth16_verify_compressed", instances=32, exec_mode=1, group=4,
                               ct_mode="commit_host", steps=1),
              "batch": dict(circuit="fq12_mul", instances=6144, exec_mode=2, group=0, ct_mode="commit", steps=3)}[args.workload]
    for k, v in preset.items():
        if getattr(args, k) is None:
            setattr(args, k, v)
    if args.no_commit:
        args.ct_mode = "none"
    args.cpu_circuit = args.circuit

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gsv_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    hasher = g.HASH_AES if args.hasher == "aes" else g.HASH_BLAKE3
    ct_mode = {"none": g.CT_NONE, "commit": g.CT_COMMIT, "commit_host": g.CT_COMMIT_HOST}[args.ct_mode]
    t_plan = time.perf_counter()
    prog = g.Program(args.circuit)
    t_plan = time.perf_counter() - t_plan
    B = args.instances
    sess = g.Session(prog, B, device=local, group=args.group, worker_threads=args.worker_threads, ct_mode=ct_mode,
                     exec_mode=args.exec_mode)
    lane = args.exec_mode == 2 or (args.exec_mode == 0 and args.group == 0 and B >= 128)
    kernel_name = "k_lane" if lane else "k_engine"
    commit_txt = {"none": "no commitment (ciphertexts dropped)",
                  "commit": "bit-exact AES chain commitment fused into the kernel (chain CTAs)",
                  "commit_host": "bit-exact AES chain commitment, gate hashes on the GPU, serial chain folded by "
                                 "host AES-NI threads draining the ciphertext ring during the kernel"}[args.ct_mode]
    # cut-and-choose seeds: instance i of rank r (garbler.rs:201-203 draws them from one RNG;
    # here a fixed arithmetic pattern so every rank/step is reproducible)
    def seeds_for(step):
        base = np.uint64(0x9E3779B97F4A7C15)
        idx = np.arange(B, dtype=np.uint64) + np.uint64((rank * 1_000_003 + step) * B)
        return idx * base + np.uint64(12345)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    commits_dev = torch.empty((B, 16), dtype=torch.uint8, device="cuda")
    gathered = torch.empty((world * B, 16), dtype=torch.uint8, device="cuda") if world > 1 else None

    def step(i, want_labels):
        t1 = time.perf_counter()
        res = sess.garble(seeds_for(i), hasher, want_inputs=want_labels, want_outputs=want_labels)
        res.wall_ms = 1e3 * (time.perf_counter() - t1)
        if world > 1:
            # the only collective of the path: gather the per-instance commitments
            commits_dev.copy_(torch.from_numpy(res.ct_commit))
            dist.all_gather_into_tensor(gathered, commits_dev)
        return res

    for i in range(args.warmup):
        step(i, True)

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = garble_ms = commit_ms = seed_ms = 0.0
    launches = 0
    for i in range(args.steps):
        res = step(args.warmup + i, True)
        dev_ms += res.ms_total
        garble_ms += res.ms_garble
        commit_ms += res.ms_commit
        seed_ms += res.ms_seed
        launches += res.n_launches
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # max over ranks, on-device time for `value`, wall (incl. copies) for e2e
    tt = torch.tensor([dev_ms, wall * 1e3, garble_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, garble_ms_max = [float(x) for x in tt.tolist()]

    gates_per_step = prog.n_gates * B * world
    value = gates_per_step * args.steps / (dev_ms_max * 1e-3)
    e2e = gates_per_step * args.steps / (wall_ms_max * 1e-3)
    h2d = 8 * B
    d2h = B * 16 * (4 + prog.n_inputs + prog.n_outputs)
    if args.ct_mode == "commit_host":
        d2h += B * 16 * prog.n_ciphertexts  # the ciphertext stream itself is drained to the host

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel = the persistent engine kernel (garbling + fused chain commitment); its
        # per-launch device time comes from CUDA events recorded by the library on its own stream
        k_ms = garble_ms / args.steps
        k_gates = prog.n_gates * B
        achieved = k_gates * ALGO_BYTES_PER_GATE / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "kernel": kernel_name, "kernel_ms": k_ms, "peak_source": peak_src,
                    "algorithmic_bytes_per_gate": ALGO_BYTES_PER_GATE}
        if prog.critical_path_levels and not lane:
            # what actually bounds the levelised kernel at this batch size: the circuit's dependency chain
            roofline["latency_floor"] = {
                "critical_path_levels": prog.critical_path_levels,
                "us_per_level_at_measured_time": 1e3 * k_ms / prog.critical_path_levels,
                "note": "one barrier-separated level = a dependent fixed-key AES through shared-memory tables; "
                        "kernel time / critical-path levels is the per-level latency if nothing else limited the run"}
        if args.ct_mode == "commit_host":
            roofline["gpu_chain_floor_s"] = prog.n_ciphertexts * 0.46e-6  # measured dependent-AES step, profiles/r01_chain_poll.md
        try:
            blocks = g.bench_hash(hasher, 1 << 28, 2, device=local)
            nonfree = sum(prog.type_count[:8]) / prog.n_gates
            need = nonfree * (AES_BLOCKS_PER_GATE_GARBLE + (1.0 if args.ct_mode == "commit" else 0.0)) * k_gates / (k_ms * 1e-3)
            roofline["alu"] = {"hash_blocks_per_s_peak": blocks, "hash_blocks_per_s_achieved": need,
                               "frac": need / blocks, "note": "register-resident 2-block gate-hash micro-kernel"}
        except Exception as e:  # pragma: no cover
            roofline["alu"] = {"error": str(e)}
        line = {
            "metric": "garbled_gates_per_s", "value": value, "unit": "gates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": f"{args.circuit} x {B} cut-and-choose instances per GPU, garble + {commit_txt}, "
                            f"{args.hasher} gate hasher"
                            + (" (BASELINE.json configs 2/4: Groth16 verifier, 1 public input, synthetic vk; "
                               "11.46 G gates here vs the reference's 11.17 G for its own vk)"
                               if args.circuit == "groth16_verify_compressed" else ""),
                "kernel": kernel_name, "plan_s": round(t_plan, 1),
                "gates_per_instance": prog.n_gates, "ciphertexts_per_instance": prog.n_ciphertexts,
                "instances_per_gpu": B, "l2": "inputs larger than L2 (label + ciphertext state >> 126 MB)",
                "parallelism": f"instances sharded over {world} GPU(s); NCCL all-gather of commitments only",
            },
            "phases_ms_per_step": {"seed_expand": seed_ms / args.steps, "garble": garble_ms / args.steps,
                                   "commit_tail_after_kernel": commit_ms / args.steps},
            "e2e": {"value": e2e, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and args.ct_mode == "commit_host" and not args.no_unthrottled:
            # The committed step is paced by the host fold (one dependent AES-NI chain per instance) through
            # ring back-pressure, so its kernel time understates the kernel.  One extra, untimed-for-`value`
            # step with the ciphertexts dropped shows the garbling kernel on its own.
            try:
                sess.close()
                s2 = g.Session(prog, B, device=local, group=args.group, worker_threads=args.worker_threads,
                               ct_mode=g.CT_NONE, exec_mode=args.exec_mode)
                r2 = s2.garble(seeds_for(10_000), hasher, want_inputs=False, want_outputs=False)
                s2.close()
                rate2 = prog.n_gates * B / (r2.ms_garble * 1e-3)
                roofline["kernel_without_commitment"] = {
                    "kernel_ms": r2.ms_garble, "gates_per_s": rate2,
                    "achieved_GBps": rate2 * ALGO_BYTES_PER_GATE / 1e9, "frac": rate2 * ALGO_BYTES_PER_GATE / 1e9 / peak,
                    "us_per_critical_level": 1e3 * r2.ms_garble / max(prog.critical_path_levels, 1),
                    "note": "same kernel, ciphertexts dropped: not paced by the host fold / PCIe drain"}
            except Exception as e:  # pragma: no cover
                roofline["kernel_without_commitment"] = {"error": str(e)}
        if world == 1 and args.workload == "verifier" and args.ct_mode == "commit_host" and not args.no_unthrottled:
            # The all-GPU counterpart in the same run: the large-batch regime (Fq12 mul x 6144 instances, lane
            # kernel, chain commitment fused on the GPU), 1 warm-up + 2 timed steps.  Not part of `value`.
            try:
                p2 = g.Program("fq12_mul")
                s3 = g.Session(p2, 6144, device=local, ct_mode=g.CT_COMMIT, exec_mode=2)
                sd = lambda i: (np.arange(6144, dtype=np.uint64) + np.uint64(i * 6144)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(7)
                s3.garble(sd(0), hasher, want_inputs=False, want_outputs=False)
                ms = [s3.garble(sd(1 + i), hasher, want_inputs=False, want_outputs=False).ms_total for i in range(2)]
                s3.close()
                rate3 = p2.n_gates * 6144 * len(ms) / (sum(ms) * 1e-3)
                line["gpu_fused_commit_batch"] = {
                    "workload": "fq12_mul x 6144 instances, k_lane, AES chain commitment folded by chain CTAs on the GPU",
                    "value": rate3, "unit": "gates/s", "ms_per_step": sum(ms) / len(ms),
                    "roofline_frac_hbm": rate3 * ALGO_BYTES_PER_GATE / 1e9 / peak}
            except Exception as e:  # pragma: no cover
                line["gpu_fused_commit_batch"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                rate, dt, _, aesni, what = cpu_garble_rate(args.cpu_circuit, args.cpu_baseline_instances, cores, prog=prog)
                line["cpu_baseline"] = {
                    "value": rate, "unit": "gates/s", "cores": cores, "kind": "port",
                    "sample": f"{what} on {cores} threads ({dt:.1f} s), garble + chain commitment, "
                              f"{'AES-NI' if aesni else 'portable AES'} oracle",
                }
            except Exception as e:  # pragma: no cover -- never lose the measured line to the baseline leg
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": cores, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

/*
 * gsv_cuda.h -- C ABI of libgsv_cuda.so, the B200 (sm_100a) garbling / evaluation engine.
 *
 * The reference (BitVM/garbled-snark-verifier v0.4.0) has no FFI layer: the seam this library
 * replaces is the Rust trait set around `CircuitMode` (src/circuit/modes.rs:26-51).  A Rust shim
 * crate (`gsv-cuda`) binds exactly the functions below:
 *
 *   reference interface                                       -> entry point here
 *   -----------------------------------------------------------------------------------------
 *   CircuitContext::{issue_wire,add_gate,with_named_child}     gsv_ctx_issue_wire / gsv_ctx_add_gate /
 *     (src/circuit/circuit_context_trait.rs:12-27)               gsv_ctx_component, gsv_program_record
 *   CircuitBuilder::streaming_garbling<H, CTH>                 gsv_garble_batch (GarbleMode hot loop,
 *     (src/circuit/mod.rs:185-208), GarbleMode::evaluate_gate    src/circuit/modes/garble_mode.rs:160-222)
 *   AESAccumulatingHash as CiphertextHandler                   GSV_CT_COMMIT    (src/ciphertext_hasher.rs:23-29)
 *   channel::Sender<S> / FileCiphertextHandler                 GSV_CT_KEEP + gsv_session_read_ciphertexts
 *     (src/circuit/mod.rs:160-170, cut_and_choose/ciphertext_repository.rs:59-136)
 *   CircuitBuilder::streaming_evaluation<H, SRC>               gsv_evaluate_batch (EvaluateMode hot loop,
 *     (src/circuit/mod.rs:229-250)                               src/circuit/modes/evaluate_mode.rs:123-158)
 *   commit_label / GarbledInstanceCommit::new                  gsv_commit_labels
 *     (src/cut_and_choose/mod.rs:41-65, garbler.rs:85-116)
 *   GateHasher = AesNiHasher | Blake3Hasher                    enum gsv_hasher (src/hashers/mod.rs:15-96)
 *
 * Conventions: every label crosses the ABI as 16 bytes in `S::to_bytes()` order (big-endian
 * u128, src/core/s.rs:25-32).  The caller owns all in/out buffers.  Functions return 0 on
 * success and a negative gsv_status otherwise (never unwind); gsv_last_error() describes the
 * last failure of the calling thread.  One host thread drives one device; the library uses its
 * own CUDA streams.  There is NO CPU fallback: every compute entry point fails with
 * GSV_ERR_NO_DEVICE when no CUDA device is present.
 */
#ifndef GSV_CUDA_H
#define GSV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsv_program gsv_program;
typedef struct gsv_session gsv_session;

enum gsv_status {
  GSV_OK = 0,
  GSV_ERR_INVALID = -1,
  GSV_ERR_NO_DEVICE = -2,
  GSV_ERR_CUDA = -3,
  GSV_ERR_CAPACITY = -4,
  GSV_ERR_CT_EXHAUSTED = -5 /* "Ciphertext source exhausted", evaluate_mode.rs:140-142 */
};

/* src/hashers/mod.rs:7-11 */
enum gsv_hasher { GSV_HASH_AES = 0, GSV_HASH_BLAKE3 = 1 };

/* src/core/gate_type.rs:1-15 */
enum gsv_gate_type {
  GSV_AND = 0, GSV_NAND = 1, GSV_NIMP = 2, GSV_IMP = 3, GSV_NCIMP = 4, GSV_CIMP = 5,
  GSV_NOR = 6, GSV_OR = 7, GSV_XOR = 8, GSV_XNOR = 9, GSV_NOT = 10
};

#define GSV_WIRE_FALSE 0u
#define GSV_WIRE_TRUE 1u
#define GSV_WIRE_UNREACHABLE 0xFFFFFFFFu

/* What happens to the ciphertext stream of a garbling run (the CiphertextHandler choice). */
enum gsv_ct_mode {
  GSV_CT_NONE = 0,   /* `()` handler: ciphertexts are dropped (src/circuit/mod.rs:172-178)       */
  GSV_CT_COMMIT = 1, /* AESAccumulatingHash: bit-exact chain commitment, stream not kept         */
  GSV_CT_KEEP = 2,   /* commitment + the stream stays in HBM for evaluation / read-back / P2P     */
  GSV_CT_KEEP_RAW = 3, /* stream kept, commitment not computed (channel::Sender<S> handler)         */
  GSV_CT_COMMIT_HOST = 4 /* same commitment as GSV_CT_COMMIT; every gate hash runs on the GPU, but the
                            strictly serial chain is folded by host AES-NI threads that drain the
                            ciphertext ring over PCIe while the kernel runs.  For few-instance runs of
                            billion-ciphertext circuits, where one dependent AES per 0.46 us on the GPU
                            (23 min for the Groth16 verifier) would dwarf the garbling itself.         */
};

const char* gsv_last_error(void);
/* Number of visible CUDA devices (0 when none). */
int gsv_device_count(void);
const char* gsv_version(void);

/* ------------------------------------------------------------------------------------------
 * Topology recording: the C mirror of CircuitContext (circuit_context_trait.rs:12-27).
 * The recorder runs the reference's two passes (metadata/credits, execution) itself; a caller
 * describes each component as a callback that re-emits the body on demand.
 * ------------------------------------------------------------------------------------------ */
typedef struct gsv_ctx gsv_ctx; /* opaque per-callback context */
/* body(ctx, user, inputs[n_in], outputs[arity]) must issue wires / add gates / call children
 * deterministically; it is invoked once for the credits pass and once per liveness variant. */
typedef void (*gsv_body_fn)(gsv_ctx* ctx, void* user, const uint32_t* inputs, uint32_t n_in,
                            uint32_t* outputs, uint32_t arity);

uint32_t gsv_ctx_issue_wire(gsv_ctx* ctx);
void gsv_ctx_add_gate(gsv_ctx* ctx, int gate_type, uint32_t a, uint32_t b, uint32_t c);
/* with_named_child: `key` = component name + off-circuit parameters (component_key.rs:15-39). */
void gsv_ctx_component(gsv_ctx* ctx, const char* key, const uint32_t* inputs, uint32_t n_in,
                       uint32_t arity, gsv_body_fn body, void* user, uint32_t* outputs);

typedef struct {
  uint64_t max_task_gates; /* 0 = default (600000) */
  uint32_t max_task_slots; /* 0 = default (1536 shared-memory label slots per instance) */
  uint32_t lane_only;      /* 1 = build only the lane-mode (emission order) task form: large circuits */
  uint32_t pipeline;       /* call pipelining of the levelised form: a consumer call is queued once its producers have
                              STARTED, runs once the inputs of its first window are there and gathers every later input
                              when its ready flag is set; producers publish outputs window by window.
                              0 = default (on; GSV_PIPELINE=0 in the environment turns it off), 1 = on, 2 = off */
  uint32_t window_levels;  /* device levels per pipelining window; 0 = default (64) */
} gsv_plan_options;

/* Records `root(ctx, user, inputs[n_inputs], outputs[n_outputs])` (CircuitBuilder::run_streaming,
 * src/circuit/mod.rs:253-301) and plans it into a levelised task program. */
gsv_program* gsv_program_record(const char* name, uint32_t n_inputs, uint32_t n_outputs,
                                gsv_body_fn root, void* user, const gsv_plan_options* opt);
/* The named workload circuits built by the library's own C++ gadget restatement
 * (src/gadgets/**): "fq12_mul", "fq6_mul", "fq2_mul", "fq_mul", "fq_add", "fq_expr",
 * "gate_zoo", "bn_mul<N>", "fq_inverse", "fq_sqrt", "fq2_sqrt", "g1_add", "g1_msm1", "fq12_square",
 * "fq12_cyclotomic_square", "fq12_inverse", "fq12_frobenius<i>", "final_exponentiation",
 * "miller_loop_groth16", "groth16_verify_compressed" (src/gadgets/groth16.rs:250-268) and "groth16_verify"
 * (src/gadgets/groth16.rs:57-110 on uncompressed points, 2 286 inputs, what examples/groth16_garble.rs
 * garbles), both with one public input over the deterministic synthetic verifying key of
 * gsv_groth16_synthetic_inputs. */
gsv_program* gsv_program_build(const char* circuit, const gsv_plan_options* opt);

/* ExecuteMode (src/circuit/modes/execute_mode.rs): plain boolean evaluation of the recorded
 * topology on the host.  A topology self-check (what the reference's gadget tests do), not a
 * garbling path.  input_bits: n_inputs bytes (0/1), output_bits: n_outputs bytes. */
int gsv_program_execute(const gsv_program* p, const uint8_t* input_bits, uint8_t* output_bits,
                        uint64_t* gates_executed);
/* Gate-level dependency depth of the circuit's outputs: over all live gates and over the non-free gates
 * only (each a dependent gate hash): the floor of any schedule's critical path. */
int gsv_program_depth(const gsv_program* p, uint64_t* depth_all, uint64_t* depth_nonfree);

/* The same boolean evaluation, but over the PLANNED program (tasks, calls, task-local slots, recycled
 * global slots) in call order, in the levelised (lane_form = 0) or emission-order (lane_form = 1)
 * task form: a host-side planner self-check that needs no GPU.  Fails if the plan ever reads a
 * slot that was not written, or if a read / overwrite of a (recycled) global slot is not ordered by an
 * explicit dependency edge on the slot's last writer / its readers (the dataflow scheduler orders calls
 * by those edges alone). */
int gsv_program_execute_plan(const gsv_program* p, int lane_form, const uint8_t* input_bits,
                             uint8_t* output_bits);

/* Input bits (EncodeInput order) of a synthetic proof: n_bits = 1273 for "groth16_verify_compressed"
 * (public | a.x a.flag | b.x.c0 b.x.c1 b.flag | c.x c.flag, src/garbled_groth16.rs:433-494) or 2286 for
 * "groth16_verify" (public | a.x a.y | b.x b.y | c.x c.y, src/garbled_groth16.rs:141-176).
 * The proof verifies for `public_x`; pass flip_public = 1 to get the rejecting variant. */
int gsv_groth16_synthetic_inputs(uint64_t public_x, int flip_public, uint8_t* bits, uint32_t n_bits);
void gsv_program_destroy(gsv_program* p);

typedef struct {
  uint64_t n_gates;       /* every add_gate call (dead ones included), = GateCount total     */
  uint64_t n_live_gates;
  uint64_t n_ciphertexts; /* live non-free gates                                            */
  uint64_t type_count[11];
  uint32_t n_inputs;
  uint32_t n_outputs;
  uint32_t n_tasks;
  uint32_t n_calls;
  uint32_t n_global_slots;
  uint32_t max_task_slots;
  uint32_t max_task_levels;
  uint32_t max_call_deps;
  uint64_t sum_call_levels; /* sum over calls of their task's level count */
  uint64_t critical_path_gates;  /* longest dependency chain through the calls, in gates: the
                                    per-instance latency floor of the lane mode */
  uint64_t critical_path_levels; /* the same chain in levels (levelised mode) */
} gsv_program_info;
int gsv_program_get_info(const gsv_program* p, gsv_program_info* out);

/* Flat emission-order gate stream (SSA wire ids: 0/1 constants, 2.. inputs, then one id per
 * written wire; c = GSV_WIRE_UNREACHABLE for dead gates).  Checkers feed this to the CPU
 * oracle.  Pass NULL arrays to query sizes.  Returns the gate count or a negative status. */
int64_t gsv_program_flat_stream(const gsv_program* p, uint8_t* type, uint32_t* a, uint32_t* b,
                                uint32_t* c, uint64_t capacity, uint32_t* outputs,
                                uint32_t* n_wires);

/* ------------------------------------------------------------------------------------------
 * Execution.  A session owns the device state of `n_instances` instances of one program on
 * one GPU: per-instance global label slots, deltas, the interleaved ciphertext buffer.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int device;             /* CUDA device ordinal */
  uint32_t n_instances;   /* batch size B (cut-and-choose instances on this GPU) */
  uint32_t group;         /* instances per work item: 1,2,4,8; 0 = auto */
  uint32_t worker_threads;/* threads per worker: 64/128/256/512; 0 = auto (256) */
  uint32_t ct_mode;       /* enum gsv_ct_mode */
  uint32_t ct_ring_log2;  /* GSV_CT_COMMIT: cap the ciphertext ring at 2^n entries per instance; 0 = auto */
  uint32_t exec_mode;     /* 0 = auto, 1 = levelised (labels in shared memory, small batches),
                             2 = lane (one warp = 32 instances, emission order; large batches) */
  uint32_t sm_limit;      /* 0 = every SM; otherwise the persistent grid takes this many SMs, so that several
                             sessions (software-pipelined cut-and-choose batches) share one GPU spatially */
  uint64_t ct_buffer_bytes; /* 0 = auto (85 % of the free HBM); otherwise an upper bound for the ciphertext ring */
  uint32_t host_threads;  /* GSV_CT_COMMIT_HOST: host fold threads of this session; 0 = auto (half the hardware
                             threads, at most one per four instances) */
  uint32_t reserved;
} gsv_session_options;

gsv_session* gsv_session_create(const gsv_program* p, const gsv_session_options* opt);
void gsv_session_destroy(gsv_session* s);

/* Garbler -> evaluator streaming (streaming_garbling_with_sender + streaming_evaluation over a channel,
 * examples/groth16_garble.rs:170-267, tests/garbler_evaluator_connection.rs:64-156): links a garbling
 * session to an evaluating session of the same program and batch size.  The ciphertexts never touch the
 * host: they travel through a ring in the EVALUATOR's device memory that the garbler's kernel fills
 * directly -- peer stores over NVLink when the two sessions are on different GPUs -- behind progress words;
 * nothing is kept.  ring_bytes caps the ring (0 = 85 % of the evaluator GPU's free memory).
 * Afterwards gsv_garble_batch(garbler) and gsv_evaluate_batch(evaluator) (ct_streams = NULL) must be called
 * CONCURRENTLY from two host threads, once per run; the garbler's ct_mode is ignored (its ct_commit is not
 * written), the evaluator's ct_commit is the chain hash of what it received (host AES-NI threads draining its
 * own ring), to be compared with the garbler's earlier commitment exactly as the reference's evaluator does. */
int gsv_session_link(gsv_session* garbler, gsv_session* evaluator, uint64_t ring_bytes);

typedef struct {
  /* all optional (NULL = not wanted); host pointers */
  uint8_t* delta;         /* B * 16        secret                                        */
  uint8_t* false_label0;  /* B * 16        GarbleMode::false_value().label0              */
  uint8_t* true_label0;   /* B * 16                                                      */
  uint8_t* input_label0;  /* B * n_inputs * 16, EncodeInput order                        */
  uint8_t* output_label0; /* B * n_outputs * 16                                          */
  uint8_t* ct_commit;     /* B * 16        AESAccumulatingHash::finalize                 */
  /* filled by the call */
  uint64_t n_ciphertexts; /* per instance */
  float ms_seed;          /* device time of the seed-expansion kernel (CUDA events)      */
  float ms_garble;        /* device time of the garbling kernel                          */
  float ms_commit;        /* device time of the chain-commitment kernel                  */
  float ms_total;         /* first launch -> last kernel done                            */
  uint32_t n_launches;    /* kernels launched by this call                               */
  uint32_t reserved;
  /* GSV_CT_COMMIT_HOST: where the step's time went on the host side (shares of the call's duration) */
  float host_fold_busy;         /* busiest fold thread: time inside the AES-NI fold                 */
  float host_drain_wait_kernel; /* drain loop waiting for the kernel to publish more ciphertexts    */
  float host_drain_wait_fold;   /* drain loop waiting for a host buffer the fold threads still hold */
  float reserved2;
} gsv_garble_result;

/* The recorded circuit as its memoised template DAG (what gsv_program_flat_stream expands), for checkers
 * that walk circuits too large to flatten (the 11 G-gate verifier is a few MB in this form).
 *   tmpl   : 12 words per template: n_in, n_wires, gate_off, n_gates, call_off, n_calls, item_off,
 *            n_items, call_wire_off, n_call_wires, out_off, n_outs (offsets into the arrays below)
 *   gates  : 4 words per gate: a, b, c (template-local wire ids; c = GSV_WIRE_UNREACHABLE when dead), type
 *   calls  : 3 words per call: callee template, in_off, out_off (relative to the template's call wires)
 *   items  : emission order inside a template: bit 31 = call, low bits = index into its gates / calls
 *   call_wires, outs : local wire ids (0/1 constants, 2.. inputs, then internal)
 * Pass NULL arrays to query the six sizes (in words) in sizes[0..5]; root = index of the circuit's template.
 * Local ids: 0 / 1 constants, [2, 2 + n_in) inputs.  A callee output that is one of the callee's own
 * inputs or constants is a pass-through: the caller already aliases it. */
int gsv_program_export_templates(const gsv_program* p, uint64_t sizes[6], uint32_t* root, uint32_t* tmpl,
                                 uint32_t* gates, uint32_t* calls, uint32_t* items, uint32_t* call_wires,
                                 uint32_t* outs);

/* Host half of GSV_CT_COMMIT_HOST, exposed for checkers: folds n_pos stream positions of n_inst
 * instances, h[i] <- AES_K(h[i] ^ block(p, i)), block(p, i) = base + (p * pos_stride + i * inst_stride)
 * * 16 (src/ciphertext_hasher.rs:22-29).  Needs AES-NI (GSV_ERR_INVALID otherwise). */
int gsv_host_chain_fold(uint8_t* h, const uint8_t* base, uint64_t pos_stride, uint64_t inst_stride,
                        uint64_t n_pos, uint32_t n_inst);

/* The same fold over the drain layout of GSV_CT_COMMIT_HOST: n_quads quads of chains, quad q's rows at
 * base + q * quad_bytes, row p = the four chains' 16-byte blocks at stream position p (64 bytes); h holds
 * 4 * n_quads states.  One 512-bit load per quad and step on VAES hosts. */
int gsv_host_chain_fold_quads(uint8_t* h, const uint8_t* base, uint64_t quad_bytes, uint64_t n_pos,
                              uint32_t n_quads);

/* streaming_garbling for B instances: instance i uses seeds[i] (garble_mode.rs:80-97). */
int gsv_garble_batch(gsv_session* s, int hasher, const uint64_t* seeds, gsv_garble_result* res);

/* FileCiphertextHandler (src/cut_and_choose/ciphertext_repository.rs:59-136): while a GSV_CT_COMMIT_HOST session
 * garbles (or a linked evaluator receives), instance i's ciphertexts are also written to the open file
 * descriptor fds[i] in the gc_{i}.bin format (16-byte records, emission order, no header; pwrite at the
 * record's offset).  fds: B descriptors (-1 = skip that instance); NULL clears the sink.  The caller owns the
 * descriptors.  This is how finalized cut-and-choose instances (47.7 GB each for the verifier) leave the GPU:
 * straight from the pinned drain buffers, never through a resident copy of the stream. */
int gsv_session_set_ciphertext_files(gsv_session* s, const int* fds);

/* Only the seed expansion of gsv_garble_batch (GarbleMode::new + EncodeInput, garble_mode.rs:80-97,116-118):
 * fills res->delta / false_label0 / true_label0 / input_label0 (each optional) for seeds[0..B) without
 * garbling.  What a garbler needs to build the evaluator's input message before a streamed run. */
int gsv_session_expand_seeds(gsv_session* s, const uint64_t* seeds, gsv_garble_result* res);

/* Copies ciphertexts [first, first+count) of `instance` to `out` in the reference stream
 * format (count * 16 bytes, emission order = gc_{i}.bin, ciphertext_source.rs:95-101).
 * Requires GSV_CT_KEEP. */
int gsv_session_read_ciphertexts(gsv_session* s, uint32_t instance, uint64_t first, uint64_t count,
                                 uint8_t* out);

typedef struct {
  /* inputs (host pointers) */
  const uint8_t* true_label;    /* B * 16   true constant's label1  (mod.rs:278-279)            */
  const uint8_t* false_label;   /* B * 16   false constant's label0                              */
  const uint8_t* input_active;  /* B * n_inputs * 16                                             */
  const uint8_t* input_bits;    /* B * n_inputs (0/1)                                            */
  /* ciphertext source: NULL = the session's own kept stream (garbler and evaluator share the GPU) or the linked
   * garbler; otherwise B host streams of n_ciphertexts*16 bytes each (FileSource layout, e.g. mmap'ed
   * gc_{i}.bin files).                                                                            */
  const uint8_t* const* ct_streams;
  uint64_t ct_stream_len;       /* ciphertexts available per host stream                         */
  /* outputs (host pointers, optional) */
  uint8_t* output_active;       /* B * n_outputs * 16                                            */
  uint8_t* output_bits;         /* B * n_outputs                                                 */
  uint8_t* ct_commit;           /* B * 16  chain hash of the consumed ciphertexts                */
  float ms_evaluate;
  float ms_commit;
  float ms_total;
  uint32_t n_launches;
  uint32_t ct_ring_log2;        /* in, with ct_streams: 0 = automatic (the streams are uploaded whole when they fit in
                                   half of the free HBM, otherwise FED through a ring while the kernel runs and hashed
                                   by host threads, FileSource); n > 0 forces a ring of 2^n ciphertexts per instance */
} gsv_evaluate_io;

/* streaming_evaluation for B instances (evaluate_mode.rs:70-158). */
int gsv_evaluate_batch(gsv_session* s, int hasher, gsv_evaluate_io* io);

/* ExecuteMode on the GPU (src/circuit/modes/execute_mode.rs; the reference's IS_PRE_BOOLEAN_EXEC pre-check,
 * examples/groth16_cut_and_choose.rs:235-254): plain boolean evaluation of the planned circuit for n_exec
 * independent inputs at once, bit-sliced 128 per instance slot of the session (n_exec <= 128 * B), on the
 * levelised kernel with bitwise gates and no hashing.  input_bits: n_exec * n_inputs bytes (0/1), output_bits:
 * n_exec * n_outputs.  *ms (optional) receives the kernel's device time. */
int gsv_execute_batch(gsv_session* s, const uint8_t* input_bits, uint32_t n_exec, uint8_t* output_bits, float* ms);

/* commit(label) = AES128_K(label) for n labels (src/cut_and_choose/mod.rs:41-48). */
int gsv_commit_labels(int device, const uint8_t* labels, uint64_t n, uint8_t* out);

/* Raw AES / BLAKE3 gate-hash micro-kernels: n blocks, register resident; returns blocks/s in
 * *rate (the integer-ALU roof of SURVEY.md section 8d).  out (n*16, optional) gets H(x_i, gid_i)
 * for x_i = counter pattern -- used by the parity tests of the device primitives. */
int gsv_hash_blocks(int device, int hasher, const uint8_t* x, const uint64_t* gid, uint64_t n,
                    uint8_t* out);
int gsv_bench_hash(int device, int hasher, uint64_t n_blocks, int iters, double* blocks_per_s);
/* Dependent-hash latency: every thread of `warps_per_sm` warps on each SM chains n gate hashes; returns SM
 * cycles per hash (the least one barrier-separated level of dependent non-free gates can cost). */
int gsv_bench_hash_latency(int device, int hasher, uint32_t warps_per_sm, uint64_t n, double* cycles_per_hash);

#ifdef __cplusplus
}
#endif
#endif /* GSV_CUDA_H */

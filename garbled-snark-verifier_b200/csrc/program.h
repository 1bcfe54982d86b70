// program.h -- the flattened, levelised gate program the GPU executes.
//
// The reference walks gates one at a time through a slab of live wires with run-time credits
// (src/circuit/modes/garble_mode.rs:160-222, src/storage.rs).  On the GPU that per-gate
// bookkeeping is resolved ONCE at flatten time:
//
//   * the template DAG (circuit.h) is cut into TASKS: component bodies small enough that all of
//     their live wires fit in shared memory.  A task is flattened, levelised (dependency depth)
//     and its wires are packed into compact shared-memory SLOTS by interval colouring -- the
//     static replacement of the credits slab.
//   * the program is the emission-ordered list of CALLS of those tasks.  Each call carries the
//     running gate index base (the AES/BLAKE3 tweak, hashers/mod.rs:56-64) and the running
//     ciphertext index base (position in the commitment stream, ciphertext_hasher.rs:23-29), so
//     gates may execute in any dependency-respecting order and still produce the reference's
//     bytes at the reference's positions.
//   * wires that cross task boundaries live in a per-instance GLOBAL slot array in HBM/L2;
//     calls list their producer calls (RAW) and, when slots are recycled, the readers of the
//     previous occupant (WAR) as dependencies.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "circuit.h"

namespace gsv {

// 16-byte device gate record.  Slots are task-local shared-memory label slots.
constexpr uint32_t LEVEL_WIDTH_MAX = 128;  // gates per device level (7-bit width field, ring staging)

struct alignas(16) DevGate {
  uint16_t a, b, c;
  uint8_t type;
  uint8_t flags;    // bit0: has ciphertext; bits 1-7 (levelised form, first gate of a level): level width - 1
  uint32_t gid_off; // gate index relative to the call's gid_base (dead gates counted)
  uint32_t ct_off;  // bits 0-23: ciphertext index relative to the call's ct_base; bits 24-31 (levelised form,
                    // first gate of a level): number of non-free gates in the level (they come first)
};
static_assert(sizeof(DevGate) == 16, "DevGate must be 16 bytes");

struct Task {
  std::string key;
  uint32_t n_in = 0;        // input positions
  uint32_t n_out = 0;       // produced (non-passthrough, live) outputs
  uint32_t n_slots = 0;     // shared-memory slots incl. 0/1 constants
  uint32_t n_levels = 0;
  uint32_t max_width = 0;
  uint64_t n_gates_total = 0;  // gate-index advance (dead gates included)
  uint64_t n_ct = 0;
  uint64_t n_live = 0;
  std::vector<DevGate> gates;        // sorted by level, non-free first inside a level
  std::vector<uint32_t> level_off;   // n_levels + 1
  std::vector<uint16_t> in_slot;     // per input position; 0xFFFF when the input is never read
  std::vector<uint16_t> out_slot;    // per produced output
  std::vector<uint32_t> out_pos;     // callee output position of each produced output
  // Lane-mode form (one warp = 32 instances, gates in EMISSION order, no levels): slots are
  // recycled by emission-order liveness -- exactly the reference's credits slab behaviour
  // (src/storage.rs:158-179: freed on the last read) -- and index a per-worker scratch array.
  std::vector<DevGate> seq_gates;
  std::vector<uint16_t> seq_in_slot, seq_out_slot;
  uint32_t n_seq_slots = 0;
  // Pipelining analysis (levelised form): under a schedule that keeps every output at its earliest level
  // and everything else as late as that allows, the level before which input i is first read and the
  // level at which produced output k is complete.  Feeds Program-level "what if calls overlapped level by
  // level" statistics; the device programs do not use it yet.
  std::vector<uint32_t> pipe_in_need, pipe_out_ready;
  uint32_t pipe_depth = 0;
  // Windows of the levelised form (call pipelining, planner.cpp): window w = device levels [w * window_levels,
  // (w + 1) * window_levels).  win_in[win_in_off[w] .. win_in_off[w + 1]) = input positions to gather before window
  // w (the window that first reads them), win_out likewise = indices of the produced outputs that are final at the
  // end of window w.  Plans without pipelining have one window.
  uint32_t window_levels = 0xFFFFFFFFu;
  std::vector<uint16_t> win_in, win_out;
  std::vector<uint32_t> win_in_off, win_out_off;
  std::vector<uint32_t> in_need_level, out_ready_level;  // per input position (0xFFFFFFFF: never read) / produced output
};

struct Call {
  uint32_t task;
  uint64_t gid_base;
  uint64_t ct_base;
  uint32_t in_off;    // into Program::call_slots, task.n_in entries (global slots)
  uint32_t out_off;   // into Program::call_slots, task.n_out entries
  uint32_t dep_off;   // into Program::deps
  uint32_t n_deps;
  // The first n_start_deps entries are START dependencies (producers of this call's inputs: with pipelining the call
  // may start once they have STARTED and then waits for each input's ready flag), the rest DONE dependencies (readers
  // of the global slots this call overwrites: they must have completed).  Without pipelining all are DONE.
  uint32_t n_start_deps = 0;
};

struct Program {
  std::vector<Task> tasks;
  std::vector<Call> calls;
  std::vector<uint32_t> call_slots;
  std::vector<uint32_t> deps;
  uint32_t n_global_slots = 0;       // incl. slots 0/1 (constants) and the circuit inputs
  uint32_t n_inputs = 0;             // circuit inputs live in global slots [2, 2+n_inputs)
  std::vector<uint32_t> output_slots;  // global slot per circuit output (0/1 = constants)
  uint64_t total_gates = 0;
  uint64_t total_ct = 0;
  uint64_t total_live = 0;
  uint32_t max_task_slots = 0;
  uint32_t max_task_seq_slots = 0;
  bool has_levelised = true;         // false: planned lane-only (PlanOptions::build_levelised off)
  bool pipelined = false;            // calls carry start dependencies and tasks window tables (PlanOptions::pipeline)
  uint32_t max_task_in = 0;
  uint32_t max_call_deps = 0;
  uint64_t type_count[11] = {0};
};

struct PlanOptions {
  uint64_t max_task_gates = 600000;  // templates above this are structural (split into children)
  uint32_t max_task_slots = 1536;    // shared-memory label slots per instance
  bool alap = true;                  // schedule gates as late as possible (smaller live sets)
  uint32_t reuse_distance = 512;     // global slot blocks are recycled no sooner than this many calls
                                     // after their last reader (0: never recycle)
  uint32_t max_global_slots = 6u << 20;  // fresh global slots are preferred over delaying WAR edges up to here
                                         // (16 B x instances each: 3 GB at 32 instances)
  uint64_t small_task_gates = 8192;  // bodies up to this size are tasks even when they occur once
  uint64_t min_shared_calls = 4;     // larger bodies become tasks only if they occur this often
  bool build_levelised = true;       // false: lane-mode form only (large circuits)
  bool pipeline = false;             // pipeline calls level-window by level-window (levelised form only)
  uint32_t window_levels = 64;       // device levels per window
};

// Cuts, flattens, levelises and packs.  Throws on circuits it cannot plan.
Program plan_program(const Builder& b, uint32_t root, const PlanOptions& opt);

// Schedules one flattened leaf (exposed for tests / statistics).
Task compile_task(const Builder& b, uint32_t tmpl, const PlanOptions& opt);

}  // namespace gsv

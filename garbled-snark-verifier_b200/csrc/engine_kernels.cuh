// engine_kernels.cuh -- the persistent task-dataflow kernels of the garbling engine (sm_100a).
//
// Execution model:
//   grid  = one CTA per SM (persistent); a CTA = NW garble workers of NT threads
//           + NC chain warps (commitment consumers);
//   work  = items (call, instance-group) scheduled by dataflow: an item becomes ready when its last
//           RAW / WAR predecessor completes (per-item counters, reverse-edge lists); the finishing
//           worker keeps one ready successor, the rest go through a global FIFO ready queue;
//   task  = gather the call's input labels from the instance's global slot array into shared
//           memory, run the task level by level (named barrier per level, labels never leave
//           shared memory, gate records streamed through a cp.async ring), write ciphertexts
//           straight to their stream position (a ring when the stream is not kept), scatter the
//           produced labels back, release the successors;
//   chain = each chain warp owns 8 instances (a lane quad per instance) and folds their ciphertext
//           streams into the bit-exact commitment h <- AES_K(h ^ ct) in emission order, call by
//           call, as soon as the producing items are done; its progress counter is the ring's
//           back-pressure (ring governor).  This is the "commitment fused into the garbling kernel"
//           of the north star: the serial chain runs concurrently with garbling instead of after
//           it.  GSV_CT_COMMIT_HOST replaces the chain warps by a publisher warp and host threads.
// Thread mapping inside a level (garbling): lane pair u -> (gate = u / G, instance = u % G), the two
// half-gate hashes on the two lanes of the pair; evaluation: thread idx -> (idx / G, idx % G).  The G
// instances of a gate sit in adjacent lanes so label accesses are contiguous 16-byte vectors and the
// gate record load is a broadcast.
#pragma once
#include "device_hash.cuh"

namespace gsvdev {

struct DevTaskD {
  uint32_t gate_off, n_gates, n_levels, n_in, n_out, n_slots, in_slot_off, out_slot_off;
  uint32_t n_ct;
  uint32_t seq_gate_off, n_seq_gates, n_seq_slots;  // lane-mode (emission order) form
  // windows of the levelised form (program.h): n_windows + 1 offsets each, starting at win_off; the offsets index
  // win_in / win_out relative to win_in_base / win_out_base
  uint32_t n_windows, window_levels, win_off, win_in_base, win_out_base;
};
struct DevCallD {
  uint32_t task, in_off, out_off, dep_off, n_deps, pad;
  unsigned long long gid_base, ct_base;
};

struct EngineParams {
  const uint4* gates;  // gsv::DevGate records as uint4
  const uint16_t* in_slot;
  const uint16_t* out_slot;
  const DevTaskD* tasks;
  const DevCallD* calls;
  const uint32_t* call_slots;
  const uint32_t* deps;
  uint4* labels;        // [group][global slot][G]
  uint8_t* vals;        // evaluate: plaintext bit per label, same indexing
  const uint4* delta;   // [B]
  uint4* ct;            // [ct position][B]  (garble: written, evaluate: read)
  // garble stores: ciphertext (position, instance) lives at
  //   ct[position * ct_pos_stride + (instance >> ct_qshift) * ct_quad_stride + (instance & ((1 << ct_qshift) - 1))]
  // position-major [position][B] (qshift 31, pos stride B) for the GPU consumers; GSV_CT_COMMIT_HOST keeps
  // [instance quad][ring position][4] (qshift 2, pos stride 4, quad stride 4 * ring capacity): each quad of
  // chains drains as ONE sequential host stream of 64-byte rows, a 512-bit VAES load per fold step
  unsigned long long ct_pos_stride, ct_quad_stride;
  uint32_t ct_qshift;
  uint32_t* flags;      // [call][group] == epoch when done
  // dataflow scheduler: ready queue of work items (call * n_groups + group)
  unsigned long long* queue;  // [1 << queue_log2]: (round + 1) << 32 | item
  uint32_t* pending;          // [item]: producers / slot owners still running
  const uint32_t* succ_off;   // [n_calls + 1] CSR of the reverse DONE-dependency edges (released at completion)
  const uint32_t* succ;
  // call pipelining (program.h): reverse START-dependency edges, released when an item starts; window tables;
  // one ready flag per (group, global slot), == epoch once the slot's current value is valid
  const uint32_t* start_succ_off;
  const uint32_t* start_succ;
  const uint32_t* win_in_off;
  const uint32_t* win_out_off;
  const uint16_t* win_in;
  const uint16_t* win_out;
  uint32_t* slot_flags;       // null: the plan is not pipelined (every input is complete before an item starts)
  uint32_t* sched;            // [0] queue head, [1] queue tail, [2] items completed, [3] park buckets released
  uint32_t queue_log2;
  // ring governor (commit modes with a ciphertext ring): items that would overrun the ring are parked
  unsigned long long* sched_limit;  // ciphertext index up to which ring space is free (min consumer + ring)
  uint32_t* park_head;              // [n_buckets] lock-free lists of parked items, by ring-space need
  uint32_t* park_next;              // [item]
  unsigned long long park_q;        // bucket width in ciphertexts
  uint32_t n_buckets;
  uint32_t n_progress;              // consumer progress counters to take the minimum of
  // Flow control of the ciphertext ring (the governor publishes sched_limit = min(progress words) + limit_add):
  //   garbling into a ring   : progress = what the consumers (chain warps / host drain / a linked evaluator)
  //                            have released, limit_add = ring capacity, free_until = ring capacity (first lap);
  //   evaluating from a ring : progress = what the producer has made available, limit_add = free_until = 0.
  uint32_t flow_control;            // items are admitted against sched_limit (else: everything is resident)
  unsigned long long limit_add, free_until;
  unsigned long long flow_total;    // ciphertexts in the whole stream
  const unsigned long long* ext_progress;  // optional progress word in mapped host memory (another device / the host
                                           // writes it): read at system scope
  uint32_t ct_sys;                  // the ciphertext ring is written to / read by another device: system-scope fences
  uint32_t* error_flag;       // evaluate: set to 1 on ciphertext exhaustion
  unsigned long long* chain_progress;  // [chain warp] ciphertexts folded so far
  uint4* commit;        // [B] chain result
  unsigned long long ct_ring;      // ring capacity in ciphertexts: position = index % ct_ring
                                   // (0: whole stream resident, no wrap, no back-pressure)
  unsigned long long ct_capacity;  // evaluate: ciphertexts available per instance
  uint32_t n_calls, n_groups, n_global_slots, B;
  uint32_t slots_per_worker;  // shared-memory label slots per instance reserved per worker
  uint32_t worker_threads, n_workers;
  uint32_t n_chain_warps;     // chain warps per chain CTA
  uint32_t n_chain_ctas;      // trailing CTAs of the grid that only run chain warps
  uint32_t epoch;
  uint32_t write_ct;          // garble: store ciphertexts
  uint32_t G;                 // instances per group (32 in lane mode)
  // GSV_CT_COMMIT_HOST: the (single) chain CTA only publishes how much of the stream is complete
  uint32_t host_chain;
  unsigned long long* host_ready;  // mapped host memory: ciphertexts complete in emission order
  // lane mode
  const uint4* seq_gates;
  const uint16_t* seq_in_slot;
  const uint16_t* seq_out_slot;
  uint4* scratch;             // [worker warp][scratch slot][32]
  uint8_t* scratch_vals;
  uint32_t scratch_stride;    // scratch slots per worker warp
  // GSV_PROFILE=1: per-phase cycle totals of the levelised workers (thread 0 of each worker), see PROF_*
  unsigned long long* prof;
};
enum { PROF_WAIT = 0, PROF_GATHER, PROF_AES_LEVELS, PROF_FREE_LEVELS, PROF_SCATTER, PROF_COMPLETE, PROF_N_AES_LEVELS,
       PROF_N_FREE_LEVELS, PROF_N_ITEMS, PROF_N_PASSES, PROF_TOTAL, PROF_WORDS };

// Gate-record staging of the levelised mode: a task's records are contiguous in level order, so
// they are streamed global -> shared in chunks of GATE_CHUNK records (2 KB), GATE_CHUNKS chunks in
// flight per worker, independent of the level structure (levels are <= GATE_CHUNK wide and carry
// their width in their first record).  The level loop then never waits on L2.
// GSV_RECORD_TMA (default): one thread per worker issues each chunk as ONE bulk async copy
// (cp.async.bulk global -> shared, the TMA engine's 1-D form) that completes on the mbarrier of the
// chunk's ring slot; the worker's threads wait on that mbarrier's phase when the level loop crosses
// into the chunk's look-ahead window.  GSV_RECORD_TMA=0: 16-byte cp.async per thread (Ampere style).
#ifndef GSV_RECORD_TMA
#define GSV_RECORD_TMA 1
#endif
constexpr uint32_t GATE_CHUNK = 128;
constexpr uint32_t GATE_CHUNKS = 4;
constexpr uint32_t GATE_RING = GATE_CHUNK * GATE_CHUNKS;  // records (8 KB) per worker
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// bulk async copy global -> shared (bytes a multiple of 16, both sides 16-byte aligned), completion counted on mbar
__device__ __forceinline__ void bulk_g2s(uint32_t smem_addr, const void* g, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr),
               "l"(g), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(mbar), "r"(parity)
      : "memory");
  return done != 0u;
}

// shared memory through 32-bit shared-window addresses
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Flag polling uses RELAXED gpu-scope loads: an ld.acquire.gpu compiles to LDG + CCTL.IVALL, and
// a spinning warp would invalidate its SM's L1 on every poll (52 M invalidates in the first
// profile, profiles/r01_chain_poll.md).  The acquire is a single fence after the wait succeeds.
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// words another device or the host writes (mapped host memory): system scope
__device__ __forceinline__ unsigned long long ld_sys64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- commitment consumer.  The chain h <- AES_K(h ^ ct) is strictly serial per instance, so
// its LATENCY is the floor of a committed garbling run.  One AES is therefore spread over a
// quad of lanes: lane (i, c) owns column c of instance i's state, exchanges the other three
// columns with its quad by shuffle each round and does 4 table lookups instead of 16.  A chain
// warp serves CHAIN_INST = 8 instances.  It folds the stream call by call in emission order as
// soon as the producing work items are flagged done; its progress counter is the ring's
// back-pressure.
constexpr uint32_t CHAIN_INST = 8;

__device__ __forceinline__ uint32_t quad_aes(const uint32_t* __restrict__ te, uint32_t s, const uint32_t (&rk)[11],
                                             uint32_t src1, uint32_t src2, uint32_t src3, uint32_t lb0,
                                             uint32_t lb2) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  s ^= rk[0];
#pragma unroll
  for (int r = 1; r < 10; r++) {
    const uint32_t s1 = __shfl_sync(FULL, s, src1), s2 = __shfl_sync(FULL, s, src2), s3 = __shfl_sync(FULL, s, src3);
    s = GSV_AES_COL(s, s1, s2, s3, rk[r]);
  }
  const uint32_t s1 = __shfl_sync(FULL, s, src1), s2 = __shfl_sync(FULL, s, src2), s3 = __shfl_sync(FULL, s, src3);
  return GSV_AES_LASTCOL(s, s1, s2, s3, rk[10]);
}

__device__ __forceinline__ void chain_warp(const EngineParams& p, const uint32_t* te, uint32_t cw) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t first = cw * CHAIN_INST;
  if (first >= p.B) return;
  const uint32_t n_inst = min(CHAIN_INST, p.B - first);
  const uint32_t G = p.G;
  const uint32_t g0 = first / G, ng = (first + n_inst - 1) / G - g0 + 1;
  const uint32_t col = lane & 3u, qi = lane >> 2;
  const bool active = qi < n_inst;
  const uint32_t q = lane & ~3u;
  const uint32_t src1 = q | ((col + 1) & 3u), src2 = q | ((col + 2) & 3u), src3 = q | ((col + 3) & 3u);
  const uint32_t lb0 = lane << 2, lb2 = lb0 | 128u;
  uint32_t rk[11];
#pragma unroll
  for (int r = 0; r < 11; r++) rk[r] = c_rk[4 * r + col];
  // word `col` of instance (first + qi)'s ciphertext at stream position 0
  const uint32_t* base = reinterpret_cast<const uint32_t*>(p.ct + first + (active ? qi : 0u)) + col;
  const size_t row = (size_t)p.B * 4u;  // words per stream position
  uint32_t h = 0;
  for (uint32_t c = 0; c < p.n_calls; ++c) {
    const DevCallD call = p.calls[c];
    const uint32_t n = p.tasks[call.task].n_ct;
    if (n == 0) continue;
    if (lane < ng) {
      const uint32_t* f = p.flags + (size_t)c * p.n_groups + g0 + lane;
      while (ld_acquire(f) != p.epoch) __nanosleep(200);
    }
    fence_acquire();
    __syncwarp();
    unsigned long long k = call.ct_base;
    const unsigned long long end = k + n;
    // ring position of the next ciphertext to LOAD (advanced incrementally, one wrap test per load)
    unsigned long long pos = p.ct_ring ? k % p.ct_ring : k;
    const unsigned long long wrap = p.ct_ring ? p.ct_ring : ~0ull;
    constexpr int U = 8;
    uint32_t v[U];
    if (k + U <= end) {
#pragma unroll
      for (int j = 0; j < U; j++) {
        v[j] = __ldcg(base + (size_t)pos * row);
        if (++pos == wrap) pos = 0;
      }
    }
    for (; k + U <= end; k += U) {
      uint32_t w[U];
#pragma unroll
      for (int j = 0; j < U; j++) w[j] = v[j];
      if (k + 2 * U <= end) {  // prefetch the next batch under this batch's AES latency
#pragma unroll
        for (int j = 0; j < U; j++) {
          v[j] = __ldcg(base + (size_t)pos * row);
          if (++pos == wrap) pos = 0;
        }
      }
#pragma unroll
      for (int j = 0; j < U; j++) h = quad_aes(te, h ^ w[j], rk, src1, src2, src3, lb0, lb2);
    }
    for (; k < end; k++) {
      h = quad_aes(te, h ^ __ldcg(base + (size_t)pos * row), rk, src1, src2, src3, lb0, lb2);
      if (++pos == wrap) pos = 0;
    }
    __syncwarp();
    if (lane == 0) st_release64(p.chain_progress + cw, end);
  }
  if (active) reinterpret_cast<uint32_t*>(p.commit + first + qi)[col] = h;
}

// ---- dataflow scheduler.  A work item (call, instance group) enters the ready queue when its last
// producer (RAW) or last reader of the global slots it will overwrite (WAR) completes; workers pop
// tickets in FIFO order.  Unlike claiming items in emission order, a worker never sits on an item
// whose inputs are not there, so all of the circuit's call-level parallelism is exposed however many
// instance groups share the workers.  Each slot of the circular queue is tagged with its round.
// named barrier over n threads that also ANDs a predicate across them
__device__ __forceinline__ bool named_bar_and(uint32_t id, uint32_t n, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %3, 0;\n\tbarrier.red.and.pred p, %1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(n), "r"((uint32_t)pred)
      : "memory");
  return r != 0u;
}
constexpr uint32_t SCHED_DONE = 0xFFFFFFFFu;
__device__ __forceinline__ void sched_push(const EngineParams& p, uint32_t item) {
  const uint32_t pos = atomicAdd(p.sched + 1, 1u);
  const unsigned long long e = ((unsigned long long)((pos >> p.queue_log2) + 1u) << 32) | item;
  st_release64(p.queue + (pos & ((1u << p.queue_log2) - 1u)), e);
}
// one thread: next ready item, or SCHED_DONE once every item has completed
__device__ __forceinline__ uint32_t sched_pop(const EngineParams& p, uint32_t n_items) {
  const uint32_t t = atomicAdd(p.sched, 1u);
  const unsigned long long* slot = p.queue + (t & ((1u << p.queue_log2) - 1u));
  const uint32_t want = (t >> p.queue_log2) + 1u;
  for (;;) {
    const unsigned long long e = ld_acquire64(slot);
    if ((uint32_t)(e >> 32) == want) {
      fence_acquire();
      return (uint32_t)e;
    }
    if (ld_acquire(p.sched + 2) >= n_items) return SCHED_DONE;
    __nanosleep(100);
  }
}
// ---- ring governor.  With a ciphertext ring, an item may only start once the ring has room for all
// of its ciphertexts (need = ct_base + n_ct <= limit = slowest consumer + ring).  Out-of-order
// scheduling can run far ahead of the consumer, so such items are PARKED in per-bucket lists
// (bucket = need / park_q) instead of occupying a worker or cycling through the queue; one governor
// warp tracks the consumers, publishes the limit and re-queues whole buckets once they are covered.
// The item the consumer is waiting for always fits (ring >= 2 tasks), so progress is guaranteed.
constexpr uint32_t PARK_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ void park_release_bucket(const EngineParams& p, uint32_t b) {
  uint32_t h = atomicExch(p.park_head + b, PARK_EMPTY);
  __threadfence();
  while (h != PARK_EMPTY) {
    const uint32_t nx = *reinterpret_cast<volatile uint32_t*>(p.park_next + h);
    sched_push(p, h);
    h = nx;
  }
}
__device__ __forceinline__ void sched_park(const EngineParams& p, uint32_t item, unsigned long long need) {
  const uint32_t b = (uint32_t)min((unsigned long long)(p.n_buckets - 1), need / p.park_q);
  uint32_t h = *reinterpret_cast<volatile uint32_t*>(p.park_head + b);
  for (;;) {
    *reinterpret_cast<volatile uint32_t*>(p.park_next + item) = h;
    __threadfence();
    const uint32_t seen = atomicCAS(p.park_head + b, h, item);
    if (seen == h) break;
    h = seen;
  }
  __threadfence();
  // the governor may have swept this bucket just before the insertion
  if (ld_acquire(p.sched + 3) > b) park_release_bucket(p, b);
}
// true when the item may start now; otherwise it has been parked
__device__ __forceinline__ bool sched_ring_admit(const EngineParams& p, uint32_t item) {
  const uint32_t ci = item / p.n_groups;
  const uint32_t n_ct = p.tasks[p.calls[ci].task].n_ct;
  if (n_ct == 0) return true;  // touches no ring position
  const unsigned long long need = p.calls[ci].ct_base + n_ct;
  if (need <= p.free_until || need <= ld_acquire64(p.sched_limit)) return true;
  sched_park(p, item, need);
  return false;
}
__device__ __forceinline__ void governor_warp(const EngineParams& p) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_items = p.n_calls * p.n_groups;
  uint32_t released = 0;
  for (;;) {
    const bool done = ld_acquire(p.sched + 2) >= n_items;
    unsigned long long m = ~0ull;
    for (uint32_t i = lane; i < p.n_progress; i += 32) m = min(m, ld_acquire64(p.chain_progress + i));
    if (p.ext_progress && lane == 0) {
      m = min(m, ld_sys64(p.ext_progress));
      fence_sys();  // what the other side wrote before advancing the word is visible from here on
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
    const unsigned long long limit = m + p.limit_add;
    if (lane == 0) st_release64(p.sched_limit, limit);
    // buckets the limit covers entirely; everything once the whole stream is covered (when evaluating, the
    // limit ends exactly at the stream's length, inside the last bucket)
    const uint32_t nb = limit >= p.flow_total ? p.n_buckets : (uint32_t)min((unsigned long long)p.n_buckets, limit / p.park_q);
    if (nb > released) {
      if (lane == 0) st_release(p.sched + 3, nb);
      __threadfence();
      __syncwarp();
      for (uint32_t b = released + lane; b < nb; b += 32) park_release_bucket(p, b);
      released = nb;
    }
    if (done) break;
    __nanosleep(2000);
  }
}

// Completion of an item: its successors' counters drop; those that reach zero become ready.  One
// ready successor is KEPT by the completing worker (no queue round trip on the dependent chains
// that make up a circuit's critical path), the others are queued.  Every thread must have fenced
// its global stores and synchronised with the others before.
constexpr uint32_t SCHED_NONE = 0xFFFFFFFEu;
// worker of nthreads threads; `keep` is the worker's shared-memory word (SCHED_NONE on entry)
// an item has STARTED (its output slots' ready flags are invalidated): consumers waiting only for that may start
__device__ __forceinline__ void sched_started_cta(const EngineParams& p, uint32_t call_i, uint32_t grp, uint32_t tid,
                                                  uint32_t nthreads) {
  const uint32_t lo = p.start_succ_off[call_i], hi = p.start_succ_off[call_i + 1];
  for (uint32_t k = lo + tid; k < hi; k += nthreads) {
    const uint32_t it = p.start_succ[k] * p.n_groups + grp;
    if (atomicSub(p.pending + it, 1u) == 1u) {
      __threadfence();
      sched_push(p, it);
    }
  }
}
__device__ __forceinline__ void sched_complete_cta(const EngineParams& p, uint32_t call_i, uint32_t grp, uint32_t tid,
                                                   uint32_t nthreads, uint32_t* keep) {
  const uint32_t lo = p.succ_off[call_i], hi = p.succ_off[call_i + 1];
  for (uint32_t k = lo + tid; k < hi; k += nthreads) {
    const uint32_t it = p.succ[k] * p.n_groups + grp;
    if (atomicSub(p.pending + it, 1u) == 1u) {
      __threadfence();  // order after the other producers' releases observed through the counter
      if (atomicCAS(keep, SCHED_NONE, it) != SCHED_NONE) sched_push(p, it);
    }
  }
  if (tid == 0) {
    st_release(p.flags + (size_t)call_i * p.n_groups + grp, p.epoch);
    atomicAdd(p.sched + 2, 1u);
  }
}
// one warp; returns the kept item (uniform) or SCHED_NONE
__device__ __forceinline__ uint32_t sched_complete_warp(const EngineParams& p, uint32_t call_i, uint32_t grp,
                                                        uint32_t lane) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lo = p.succ_off[call_i], hi = p.succ_off[call_i + 1];
  uint32_t keep = SCHED_NONE;
  for (uint32_t k = lo; k < hi; k += 32) {
    uint32_t it = 0;
    bool rdy = false;
    if (k + lane < hi) {
      it = p.succ[k + lane] * p.n_groups + grp;
      rdy = atomicSub(p.pending + it, 1u) == 1u;
      if (rdy) __threadfence();
    }
    const uint32_t m = __ballot_sync(FULL, rdy);
    if (m) {
      uint32_t first = 32;
      if (keep == SCHED_NONE) {
        first = __ffs(m) - 1;
        keep = __shfl_sync(FULL, it, first);
      }
      if (rdy && lane != first) sched_push(p, it);
    }
  }
  if (lane == 0) {
    st_release(p.flags + (size_t)call_i * p.n_groups + grp, p.epoch);
    atomicAdd(p.sched + 2, 1u);
  }
  return keep;
}

// pending[item] = number of dependencies; items without any are pushed right away
__global__ void k_sched_init(const EngineParams p) {
  const uint32_t n_items = p.n_calls * p.n_groups;
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.sched_limit = p.free_until;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const uint32_t nd = p.calls[i / p.n_groups].n_deps;
    p.pending[i] = nd;
    if (nd == 0) sched_push(p, i);
  }
}

// GSV_CT_COMMIT_HOST: instead of folding, one warp tracks the emission-order frontier of finished
// calls (all instance groups) and publishes the number of complete stream positions to mapped host
// memory; the host drains the ring by DMA and folds the chains with AES-NI (host_chain.h).  Ring
// back-pressure comes back through chain_progress[0], which the host advances after each drain.
__device__ __forceinline__ void publish_warp(const EngineParams& p) {
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long end = 0;
  for (uint32_t c = 0; c < p.n_calls; ++c) {
    const DevCallD call = p.calls[c];
    const uint32_t n = p.tasks[call.task].n_ct;
    if (n == 0) continue;
    for (uint32_t g = lane; g < p.n_groups; g += 32) {
      const uint32_t* f = p.flags + (size_t)c * p.n_groups + g;
      while (ld_acquire(f) != p.epoch) __nanosleep(200);
    }
    fence_acquire();
    __syncwarp();
    end = call.ct_base + n;
    if (lane == 0) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(p.host_ready) = end;
    }
  }
  if (lane == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(p.host_ready) = end;
  }
}

// MODE 0 = garble (labels are label0, ciphertexts produced), MODE 1 = evaluate, MODE 2 = execute: plain boolean
// evaluation (ExecuteMode, src/circuit/modes/execute_mode.rs) bit-sliced 128 wide -- a "label" is then the
// wire's value in 128 independent executions, gates are bitwise operations, nothing is hashed.
// At most ENGINE_MAX_THREADS threads per CTA: 128 registers per thread are available, which lets ptxas keep
// a round's table lookups in flight together (with the 64 registers of a 1024-thread CTA every lookup was
// consumed ~8 instructions after its issue and a worker alone on its scheduler paid the shared-memory
// latency ~16 times per AES round; profiles/r02_level_loop.md).
constexpr uint32_t ENGINE_MAX_THREADS = 512;
template <int G, int HASH, int MODE>
__global__ void __launch_bounds__(ENGINE_MAX_THREADS, 1) k_engine(const EngineParams p) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);  // 64 KB table block
  constexpr uint32_t TE_Q = AES_TABLE_BYTES / 16;    // in uint4 units
  const uint32_t NT = p.worker_threads;
  const uint32_t n_workers = p.n_workers;
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();

  if (blockIdx.x >= gridDim.x - p.n_chain_ctas) {
    // ---- chain CTA: the last n_chain_ctas SMs only run commitment consumers, a few warps per
    // SMSP, so the latency-bound chain never competes with garbling warps for issue slots (evaluating
    // from a ring: one such CTA with the governor and the publisher of the consumed frontier)
    const uint32_t warp = threadIdx.x >> 5;
    if (p.flow_control && blockIdx.x == gridDim.x - p.n_chain_ctas && warp == (blockDim.x >> 5) - 1) {
      governor_warp(p);
    } else if (p.host_chain) {
      if (warp == 0) publish_warp(p);
    } else if (MODE == 0 && warp < p.n_chain_warps) {
      chain_warp(p, te, (blockIdx.x - (gridDim.x - p.n_chain_ctas)) * p.n_chain_warps + warp);
    }
    return;
  }

  if (threadIdx.x >= n_workers * NT) return;
  const uint32_t worker = threadIdx.x / NT;
  const uint32_t wt = threadIdx.x - worker * NT;
  const uint32_t bar_id = worker + 1;
  // per-worker regions
  const uint32_t lab_words = p.slots_per_worker * G;  // uint4 entries
  uint4* lab = smem + TE_Q + worker * lab_words;
  uint8_t* tail = reinterpret_cast<uint8_t*>(smem + TE_Q + n_workers * lab_words);
  uint8_t* sval = tail + worker * lab_words;
  uint8_t* tail2 = tail + (MODE == 1 ? n_workers * lab_words : 0);
  volatile uint32_t* ctrl = reinterpret_cast<volatile uint32_t*>(tail2) + 2 * worker;
  uint32_t* keep = reinterpret_cast<uint32_t*>(tail2) + 2 * worker + 1;  // successor kept at completion
  if (wt == 0) *keep = SCHED_NONE;
  // gate-record ring, 16-byte aligned behind the value bytes and the ticket words
  const uint32_t tail_bytes = ((MODE == 1 ? n_workers * lab_words : 0u) + n_workers * 8u + 15u) & ~15u;
  const uint4* ring = reinterpret_cast<const uint4*>(tail + tail_bytes) + worker * GATE_RING;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);

  const uint32_t inst = wt % G;  // NT % G == 0, so a thread always serves the same instance lane
  const uint32_t n_items = p.n_calls * p.n_groups;

  // GSV_PROFILE: thread 0 of each worker accumulates per-phase cycles in its shared-memory area (a few
  // instructions per level; global reductions per level measurably slowed the run they were measuring)
  const bool prof = p.prof != nullptr && wt == 0;
  volatile unsigned long long* wprof = reinterpret_cast<volatile unsigned long long*>(
      reinterpret_cast<uint4*>(tail + tail_bytes) + n_workers * GATE_RING) + worker * PROF_WORDS;
  if (prof)
    for (int i = 0; i < PROF_WORDS; i++) wprof[i] = 0;
#if GSV_RECORD_TMA
  // one mbarrier per slot of the worker's record ring, behind the profile words; ring_par holds, per slot, the
  // phase parity the next wait expects (every issued chunk is waited for exactly once, by every thread)
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<const uint4*>(tail + tail_bytes) + n_workers * GATE_RING) +
                          n_workers * PROF_WORDS * 8u + worker * (8u * GATE_CHUNKS);
  uint32_t ring_par = 0;
  if (wt == 0) {
#pragma unroll
    for (uint32_t k = 0; k < GATE_CHUNKS; k++) mbar_init(mbar_s + 8u * k, 1u);
    mbar_init_fence();
  }
#endif
  long long t_prev = prof ? clock64() : 0;
  const long long t_begin = t_prev;
  auto lap = [&](int slot) {
    if (prof) {
      const long long t = clock64();
      wprof[slot] += (unsigned long long)(t - t_prev);
      t_prev = t;
    }
  };
  for (;;) {
    if (wt == 0) {
      uint32_t it = *keep;
      *keep = SCHED_NONE;
      for (;;) {
        if (it == SCHED_NONE) it = sched_pop(p, n_items);
        if (it == SCHED_DONE || !p.flow_control || sched_ring_admit(p, it)) break;
        it = SCHED_NONE;  // parked until the ring has room
      }
      *ctrl = it;
    }
    named_bar(bar_id, NT);
    const uint32_t item = *ctrl;
    if (item == SCHED_DONE) break;
    const uint32_t call_i = item / p.n_groups;
    const uint32_t grp = item - call_i * p.n_groups;
    const DevCallD call = p.calls[call_i];
    const DevTaskD task = p.tasks[call.task];
    // Call pipelining (program.h): the item's levels are cut into windows; the inputs a window reads first are
    // gathered at its start -- each behind the ready flag of its global slot, its producer may still be running --
    // and the outputs a window completes are published (label, fence, flag) at its end.  Plans without
    // pipelining have one window and no flags.
    const size_t gbase = (size_t)grp * p.n_global_slots;
    uint32_t* const sflags = p.slot_flags ? p.slot_flags + gbase : nullptr;
    if (sflags) {
      // An item is queued as soon as its producers have STARTED.  It only occupies the worker once the inputs of
      // its first window are there; until then it goes back to the end of the queue, so workers are not held by
      // items whose producers are still far from publishing.
      bool ok = true;
      const uint32_t lo = p.win_in_off[task.win_off], hi = p.win_in_off[task.win_off + 1];
      for (uint32_t k = lo + wt; k < hi; k += NT)
        ok &= *(volatile uint32_t*)(sflags + p.call_slots[call.in_off + p.win_in[task.win_in_base + k]]) == p.epoch;
      if (!named_bar_and(bar_id, NT, ok)) {
        if (wt == 0) {
          __nanosleep(256);
          sched_push(p, item);
        }
        continue;
      }
    }
    lap(PROF_WAIT);

    // ---- start streaming the task's gate records (independent of the producers)
    const uint4* gsrc = p.gates + task.gate_off;
    uint32_t issued = 0;  // chunks requested so far (uniform over the worker)
#if GSV_RECORD_TMA
    const uint32_t n_chunks = (task.n_gates + GATE_CHUNK - 1) / GATE_CHUNK;
    auto issue_chunk = [&]() {
      if (wt == 0 && issued < n_chunks) {
        const uint32_t base = issued * GATE_CHUNK, slot = issued & (GATE_CHUNKS - 1);
        const uint32_t bytes = min(GATE_CHUNK, task.n_gates - base) * 16u;
        mbar_expect_tx(mbar_s + 8u * slot, bytes);
        bulk_g2s(ring_s + slot * (GATE_CHUNK * 16u), gsrc + base, bytes, mbar_s + 8u * slot);
      }
      issued++;
    };
    auto wait_chunk = [&](uint32_t j) {  // chunk j of this task, requested earlier; every thread of the worker
      if (j < n_chunks) {
        const uint32_t slot = j & (GATE_CHUNKS - 1);
        while (!mbar_try_wait(mbar_s + 8u * slot, (ring_par >> slot) & 1u)) {
        }
        ring_par ^= 1u << slot;
      }
    };
#else
    auto issue_chunk = [&]() {
      const uint32_t base = issued * GATE_CHUNK;
      for (uint32_t r = base + wt; r < base + GATE_CHUNK && r < task.n_gates; r += NT)
        cp_async16(ring_s + ((r & (GATE_RING - 1)) << 4), gsrc + r);
      cp_async_commit();
      issued++;
    };
#endif
#pragma unroll
    for (uint32_t k = 0; k < GATE_CHUNKS; k++) issue_chunk();

    // ring position of the call's first ciphertext (a task never exceeds half the ring)
    const unsigned long long ct_pos0 = p.ct_ring ? call.ct_base % p.ct_ring : call.ct_base;
    // ---- gather inputs (and the two constant wires) into shared memory
    uint4 delta = make_uint4(0, 0, 0, 0);
    if (MODE == 0) delta = p.delta[grp * G + inst];
    // this thread's instance column of the ciphertext buffer (written when garbling, read when evaluating)
    const uint32_t ginst = grp * G + inst;
    uint4* const ct_out = p.ct + (size_t)(ginst >> p.ct_qshift) * p.ct_quad_stride + (ginst & ((1u << p.ct_qshift) - 1u));
    if (sflags) {
      // the values in this call's output slots are dead (their readers are done dependencies): invalidate them,
      // then let the consumers that only waited for that start
      for (uint32_t k = wt; k < task.n_out; k += NT) *(volatile uint32_t*)(sflags + p.call_slots[call.out_off + k]) = 0u;
      __threadfence();
      named_bar(bar_id, NT);
      sched_started_cta(p, call_i, grp, wt, NT);
    }
    auto gather_window = [&](uint32_t w) {
      const uint32_t lo = p.win_in_off[task.win_off + w], hi = p.win_in_off[task.win_off + w + 1];
      for (uint32_t k = wt; k < (hi - lo) * G; k += NT) {
        const uint32_t pos = p.win_in[task.win_in_base + lo + k / G];
        const uint32_t s = p.in_slot[task.in_slot_off + pos];
        const uint32_t gs = p.call_slots[call.in_off + pos];
        // The label loads below are ld.global.cg (served by L2, the coherence point) and are issued behind the
        // control dependency on the flag's value; the producer ordered label stores -> __threadfence -> barrier ->
        // flag store.  No fence here: a membar would also wait for this thread's ciphertext stores in flight.
        if (sflags)
          while (ld_acquire(sflags + gs) != p.epoch) __nanosleep(64);
        const size_t gi = (gbase + gs) * G + inst;
        lab[s * G + inst] = __ldcg(p.labels + gi);
        if (MODE == 1) sval[s * G + inst] = __ldcg(p.vals + gi);
      }
    };
    auto publish_window = [&](uint32_t w) {
      const uint32_t lo = p.win_out_off[task.win_off + w], hi = p.win_out_off[task.win_off + w + 1];
      for (uint32_t k = wt; k < (hi - lo) * G; k += NT) {
        const uint32_t idx = p.win_out[task.win_out_base + lo + k / G];
        const uint32_t s = p.out_slot[task.out_slot_off + idx];
        const size_t gi = (gbase + p.call_slots[call.out_off + idx]) * G + inst;
        p.labels[gi] = lab[s * G + inst];
        if (MODE == 1) p.vals[gi] = sval[s * G + inst];
      }
      if (sflags) {
        __threadfence();
        named_bar(bar_id, NT);
        for (uint32_t k = lo + wt; k < hi; k += NT)
          *(volatile uint32_t*)(sflags + p.call_slots[call.out_off + p.win_out[task.win_out_base + k]]) = p.epoch;
      }
    };
    if (wt < 2 * G) {
      const uint32_t s = wt / G;
      lab[s * G + inst] = __ldcg(p.labels + (gbase + s) * G + inst);
      if (MODE == 1) sval[s * G + inst] = (uint8_t)s;
    }
    gather_window(0);
#if GSV_RECORD_TMA
#pragma unroll
    for (uint32_t k = 0; k < GATE_CHUNKS; k++) wait_chunk(k);  // chunks 0..3 have landed
#else
    cp_async_wait<0>();  // chunks 0..3 have landed
#endif
    named_bar(bar_id, NT);
    lap(PROF_GATHER);
    if (prof) wprof[PROF_N_ITEMS]++;

    // ---- level loop.  A level = its non-free gates (first), then its free gates; the header in the
    // level's first record gives both counts.  Everything that does not depend on the previous level's
    // labels is software-pipelined out of the dependent path: at the start of level L a thread issues the
    // loads of its first gate records of level L + 1 and of the header of level L + 2 (their ring
    // positions follow from headers it already holds), then does level L from registers.  Between two
    // barriers the dependent chain is: label loads -> hash / XOR -> label store.
    // Ring invariant at a level start (pos in chunk c): chunks <= c + 2 are visible, chunk c + 3 is
    // requested; levels are at most one chunk wide, so L + 1's records and L + 2's header are visible.
    // All shared-memory traffic of the loop uses 32-bit shared-window addresses (ld/st.shared): with
    // generic pointers the compiler re-derived the window base (S2UR SR_CgaCtaId) at every use.
    //   garbling  : AES gates on lane pairs (lane, lane ^ G): the two hashes of a half-gate, H(A_sel) and
    //               H(A_sel ^ delta), run one per lane and the odd lane fetches the other by shuffle
    //               (it forms the ciphertext, the even lane the output label); free gates one thread
    //               per (gate, instance);
    //   evaluating: one thread per (gate, instance) for both kinds.
    constexpr uint32_t RMASK = GATE_RING - 1;
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t SLOT_B = 16u * G;                      // bytes per label slot (G instances)
    const uint32_t lab_s = (uint32_t)__cvta_generic_to_shared(lab) + inst * 16u;
    const uint32_t sval_s = (uint32_t)__cvta_generic_to_shared(sval) + inst;
    // gate index inside a free pass, counted from the worker's LAST threads: the non-free gates start at its
    // first threads, so in a narrow level the free gates run in other warps, beside the hashes, not after them
    const uint32_t f_idx = (NT - 1u - wt) / G;
    const uint32_t f_per_pass = NT / G;
    const uint32_t a_idx = MODE == 0 ? wt / (2 * G) : wt / G;  // gate index inside an AES pass
    const uint32_t a_per_pass = MODE == 0 ? NT / (2 * G) : f_per_pass;
    const uint32_t half = (wt / G) & 1u;                      // garbling: which of the two hashes
    auto rec_at = [&](uint32_t idx) { return lds128(ring_s + ((idx & RMASK) << 4)); };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    uint32_t pos = 0, chunk = 0;
    uint32_t n_tot = 0, n_nf = 0, n_tot1 = 0, n_nf1 = 0;      // headers of the current and the next level
    uint4 rec_a = zero4, rec_f = zero4;                       // this thread's first AES / free record of the level
    if (task.n_levels) {
      const uint4 h = rec_at(0);
      n_tot = ((h.y >> 25) & 0x7Fu) + 1u;
      n_nf = h.w >> 24;
      if (a_idx < n_nf) rec_a = rec_at(a_idx);
      if (n_nf + f_idx < n_tot) rec_f = rec_at(n_nf + f_idx);
      if (task.n_levels > 1) {
        const uint4 h1 = rec_at(n_tot);
        n_tot1 = ((h1.y >> 25) & 0x7Fu) + 1u;
        n_nf1 = h1.w >> 24;
      }
    }
    uint32_t lvl = 0;
    for (uint32_t win = 0; win < task.n_windows; ++win) {
    if (win) {
      gather_window(win);
      named_bar(bar_id, NT);
    }
    const uint32_t lvl_end = task.n_levels - lvl > task.window_levels ? lvl + task.window_levels : task.n_levels;
    for (; lvl < lvl_end; ++lvl) {
      // ---- prefetch: first records of level lvl + 1, header of level lvl + 2
      const uint32_t pos1 = pos + n_tot, pos2 = pos1 + n_tot1;
      uint4 rec_a1 = zero4, rec_f1 = zero4;
      if (a_idx < n_nf1) rec_a1 = rec_at(pos1 + a_idx);
      if (n_nf1 + f_idx < n_tot1) rec_f1 = rec_at(pos1 + n_nf1 + f_idx);
      const uint4 h2 = rec_at(pos2);
      // ---- non-free gates
      for (uint32_t g0 = 0; g0 < n_nf; g0 += a_per_pass) {  // uniform trip count: shuffles are full-warp
        const uint32_t gi = g0 + a_idx;
        const bool act = gi < n_nf;
        uint4 r = rec_a;
        if (g0 != 0 && act) r = rec_at(pos + gi);
        const uint32_t sa = r.x & 0xFFFFu, sb = r.x >> 16, sc = r.y & 0xFFFFu;
        const uint32_t type = (r.y >> 16) & 0xFFu;
        const unsigned long long gid = call.gid_base + r.z;
        if (MODE == 0) {
          uint4 hs = zero4;
          if (act) {
            uint4 x = xor4(lds128(lab_s + sa * SLOT_B), and4(delta, 0u - ((type >> 2) & 1u)));  // selected label
            if (half) x = xor4(x, delta);                                                      // the other one
            hs = hash1<HASH>(te, x, gid);
          }
          uint4 ho;
          ho.x = __shfl_xor_sync(FULL, hs.x, G);
          ho.y = __shfl_xor_sync(FULL, hs.y, G);
          ho.z = __shfl_xor_sync(FULL, hs.z, G);
          ho.w = __shfl_xor_sync(FULL, hs.w, G);
          if (act) {
            if (half) {
              if (p.write_ct) {
                const uint4 ct = xor4(xor4(hs, ho), xor4(lds128(lab_s + sb * SLOT_B), and4(delta, 0u - ((type >> 1) & 1u))));
                unsigned long long cpos = ct_pos0 + (r.w & 0xFFFFFFu);
                if (p.ct_ring && cpos >= p.ct_ring) cpos -= p.ct_ring;
                __stcg(ct_out + (size_t)cpos * p.ct_pos_stride, ct);
              }
            } else {
              sts128(lab_s + sc * SLOT_B, xor4(hs, and4(delta, 0u - (type & 1u))));
            }
          }
        } else if (MODE == 2) {
          if (act) {  // ((a ^ alpha_a) & (b ^ alpha_b)) ^ alpha_c on 128 executions (gate_type.rs:20-37)
            const uint4 la = lds128(lab_s + sa * SLOT_B), lb = lds128(lab_s + sb * SLOT_B);
            const uint32_t ma = 0u - ((type >> 2) & 1u), mb = 0u - ((type >> 1) & 1u), mc = 0u - (type & 1u);
            sts128(lab_s + sc * SLOT_B, make_uint4(((la.x ^ ma) & (lb.x ^ mb)) ^ mc, ((la.y ^ ma) & (lb.y ^ mb)) ^ mc,
                                                   ((la.z ^ ma) & (lb.z ^ mb)) ^ mc, ((la.w ^ ma) & (lb.w ^ mb)) ^ mc));
          }
        } else if (act) {
          const uint4 la = lds128(lab_s + sa * SLOT_B), lb = lds128(lab_s + sb * SLOT_B);
          const uint32_t va = lds8(sval_s + sa * G), vb = lds8(sval_s + sb * G);
          const unsigned long long cti = call.ct_base + (r.w & 0xFFFFFFu);
          uint4 ct = zero4;
          if (cti < p.ct_capacity) {
            unsigned long long cpos = ct_pos0 + (r.w & 0xFFFFFFu);
            if (p.ct_ring && cpos >= p.ct_ring) cpos -= p.ct_ring;
            ct = __ldcg(ct_out + (size_t)cpos * p.ct_pos_stride);  // L2 only: ring positions are rewritten
          } else {
            *p.error_flag = 1u;
          }
          sts128(lab_s + sc * SLOT_B, degarble_nonfree<HASH>(te, type, ct, la, va, lb, gid));
          sts8(sval_s + sc * G, gate_value(type, va, vb));
        }
      }
      // ---- free gates
      for (uint32_t g0 = n_nf; g0 < n_tot; g0 += f_per_pass) {
        const uint32_t gi = g0 + f_idx;
        if (gi < n_tot) {
          uint4 r = rec_f;
          if (g0 != n_nf) r = rec_at(pos + gi);
          const uint32_t sa = r.x & 0xFFFFu, sb = r.x >> 16, sc = r.y & 0xFFFFu;
          const uint32_t type = (r.y >> 16) & 0xFFu;
          const uint4 la = lds128(lab_s + sa * SLOT_B), lb = lds128(lab_s + sb * SLOT_B);
          uint4 lc = (type == 10) ? la : xor4(la, lb);
          if (MODE == 0) {
            if (type != 8) lc = xor4(lc, delta);  // Xnor / Not flip the zero label
          } else if (MODE == 2) {
            if (type != 8) lc = make_uint4(~lc.x, ~lc.y, ~lc.z, ~lc.w);  // Xnor / Not
          } else {
            sts8(sval_s + sc * G, gate_value(type, lds8(sval_s + sa * G), lds8(sval_s + sb * G)));
          }
          sts128(lab_s + sc * SLOT_B, lc);
        }
      }
      const bool cross = pos1 / GATE_CHUNK != chunk;  // uniform over the worker; at most one chunk per level
#if GSV_RECORD_TMA
      if (cross && chunk) wait_chunk(chunk + 3);      // chunk c + 3, requested one crossing ago (0..3: at the start)
#else
      if (cross) cp_async_wait<0>();                  // chunk c + 3, requested one crossing ago
#endif
      named_bar(bar_id, NT);
      if (cross) {                                    // the chunk left behind is free: request chunk c + 4
        issue_chunk();
        chunk++;
      }
      if (prof) {
        lap(n_nf ? PROF_AES_LEVELS : PROF_FREE_LEVELS);
        wprof[n_nf ? PROF_N_AES_LEVELS : PROF_N_FREE_LEVELS]++;
        wprof[PROF_N_PASSES] += (n_nf + a_per_pass - 1) / a_per_pass;
      }
      pos = pos1;
      n_tot = n_tot1;
      n_nf = n_nf1;
      rec_a = rec_a1;
      rec_f = rec_f1;
      n_tot1 = n_nf1 = 0;
      if (lvl + 2 < task.n_levels) {
        n_tot1 = ((h2.y >> 25) & 0x7Fu) + 1u;
        n_nf1 = h2.w >> 24;
      }
    }
    // ---- scatter the labels this window completed to the instance's global slots
    publish_window(win);
    }
#if !GSV_RECORD_TMA
    cp_async_wait<0>();
#endif
    if (p.ct_sys) __threadfence_system();  // ciphertexts stored to a peer's ring are ordered before the flags
    else __threadfence();
    named_bar(bar_id, NT);
    lap(PROF_SCATTER);
    sched_complete_cta(p, call_i, grp, wt, NT, keep);
    named_bar(bar_id, NT);
    lap(PROF_COMPLETE);
  }
  if (prof) {
    wprof[PROF_TOTAL] = (unsigned long long)(clock64() - t_begin);
    for (int i = 0; i < PROF_WORDS; i++) atomicAdd(p.prof + i, wprof[i]);
  }
}

// ---- lane mode: one WARP per work item, lane = instance (32 instances of the batch), the task's
// gates run sequentially in emission order.  Control flow is warp-uniform (every lane executes
// the same gate), every AES issue does 32 useful blocks, there are no barriers, and label /
// ciphertext accesses are 512-byte coalesced rows.  Task-internal wires live in a per-warp
// scratch array in global memory (L1/L2 resident: the active window of an emission-order walk is
// small) whose slots are recycled by emission-order liveness.  This is the throughput mode for
// cut-and-choose batches of hundreds to thousands of instances; k_engine (levelised, labels in
// shared memory) is the latency mode for small batches.
// 512 threads per CTA (16 worker warps, 128 registers per thread): with 1024 threads the 64-register budget
// spilled 200+ bytes per thread, and kernels with local memory make a first launch resize the context's
// local-memory pool, which synchronises the device -- a dead-lock when another session's persistent
// kernel on the same GPU is waiting for this one (linked garbler / evaluator).
constexpr uint32_t LANE_WARPS = 16;
template <int HASH, int MODE>
__global__ void __launch_bounds__(32 * LANE_WARPS, 1) k_lane(const EngineParams p) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  if (blockIdx.x >= gridDim.x - p.n_chain_ctas) {
    // chain CTA (see k_engine): dedicated SMs for the serial commitment
    if (p.flow_control && blockIdx.x == gridDim.x - p.n_chain_ctas && warp == (blockDim.x >> 5) - 1) {
      governor_warp(p);
    } else if (p.host_chain) {
      if (warp == 0) publish_warp(p);
    } else if (MODE == 0 && warp < p.n_chain_warps) {
      chain_warp(p, te, (blockIdx.x - (gridDim.x - p.n_chain_ctas)) * p.n_chain_warps + warp);
    }
    return;
  }
  if (warp >= p.n_workers) return;
  const uint32_t wid = blockIdx.x * p.n_workers + warp;
  uint4* my = p.scratch + (size_t)wid * p.scratch_stride * 32u + lane;
  uint8_t* myv = p.scratch_vals + (size_t)wid * p.scratch_stride * 32u + lane;
  const uint32_t n_items = p.n_calls * p.n_groups;
  constexpr uint32_t FULL = 0xFFFFFFFFu;

  uint32_t kept = SCHED_NONE;
  for (;;) {
    uint32_t item = kept;
    kept = SCHED_NONE;
    for (;;) {
      if (item == SCHED_NONE) {
        if (lane == 0) item = sched_pop(p, n_items);
        item = __shfl_sync(FULL, item, 0);
      }
      if (item == SCHED_DONE || !p.flow_control) break;
      uint32_t ok = 0;
      if (lane == 0) ok = sched_ring_admit(p, item) ? 1u : 0u;
      if (__shfl_sync(FULL, ok, 0)) break;
      item = SCHED_NONE;  // parked until the ring has room
    }
    if (item == SCHED_DONE) break;
    const uint32_t call_i = item / p.n_groups;
    const uint32_t grp = item - call_i * p.n_groups;
    const DevCallD call = p.calls[call_i];
    const DevTaskD task = p.tasks[call.task];
    const uint32_t instance = grp * 32u + lane;
    const bool act = instance < p.B;
    fence_acquire();
    __syncwarp();

    const unsigned long long ct_pos0 = p.ct_ring ? call.ct_base % p.ct_ring : call.ct_base;
    // ---- gather: constants + inputs -> scratch (rows of 32 labels, 512 B each)
    const size_t gbase = (size_t)grp * p.n_global_slots;
    const uint4* glab = p.labels + gbase * 32u + lane;
    const uint8_t* gval = p.vals + gbase * 32u + lane;
    uint4 delta = make_uint4(0, 0, 0, 0);
    if (MODE == 0) delta = p.delta[instance];
    uint4* const ct_out = p.ct + (size_t)(instance >> p.ct_qshift) * p.ct_quad_stride + (instance & ((1u << p.ct_qshift) - 1u));
    my[0] = __ldcg(glab);
    my[32] = __ldcg(glab + 32);
    if (MODE == 1) { myv[0] = 0; myv[32] = 1; }
    for (uint32_t base = 0; base < task.n_in; base += 32) {
      const uint32_t cnt = min(32u, task.n_in - base);
      uint32_t ms = 0xFFFFu, mg = 0;
      if (lane < cnt) {
        ms = p.seq_in_slot[task.in_slot_off + base + lane];
        mg = p.call_slots[call.in_off + base + lane];
      }
      for (uint32_t j0 = 0; j0 < cnt; j0 += 8) {
        uint4 v[8];
        uint8_t vb[8];
        uint32_t s[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s[j] = __shfl_sync(FULL, ms, (j0 + j) & 31);
          const uint32_t gs = __shfl_sync(FULL, mg, (j0 + j) & 31);
          if (j0 + j < cnt && s[j] != 0xFFFFu) {
            v[j] = __ldcg(glab + (size_t)gs * 32u);
            if (MODE == 1) vb[j] = __ldcg(gval + (size_t)gs * 32u);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j0 + j < cnt && s[j] != 0xFFFFu) {
            my[s[j] * 32u] = v[j];
            if (MODE == 1) myv[s[j] * 32u] = vb[j];
          }
      }
    }

    // ---- the task's gates, emission order, 32 records per coalesced fetch
    const uint4* gates = p.seq_gates + task.seq_gate_off;
    const uint32_t n = task.n_seq_gates;
    uint4 rec_next = make_uint4(0, 0, 0, 0);
    if (lane < n) rec_next = __ldg(gates + lane);
    for (uint32_t g0 = 0; g0 < n; g0 += 32) {
      const uint4 rec = rec_next;
      if (g0 + 32 + lane < n) rec_next = __ldg(gates + g0 + 32 + lane);
      const uint32_t cnt = min(32u, n - g0);
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t rx = __shfl_sync(FULL, rec.x, j), ry = __shfl_sync(FULL, rec.y, j);
        const uint32_t sa = rx & 0xFFFFu, sb = rx >> 16, sc = ry & 0xFFFFu;
        const uint32_t type = (ry >> 16) & 0xFFu;
        const uint4 la = my[sa * 32u];
        const uint4 lb = my[sb * 32u];
        uint4 lc;
        if (MODE == 0) {
          if (type >= 8) {
            lc = (type == 10) ? xor4(la, delta) : xor4(la, lb);
            if (type == 9) lc = xor4(lc, delta);
          } else {
            const unsigned long long gid = call.gid_base + __shfl_sync(FULL, rec.z, j);
            unsigned long long pos = ct_pos0 + __shfl_sync(FULL, rec.w, j);
            if (p.ct_ring && pos >= p.ct_ring) pos -= p.ct_ring;
            uint4 ct;
            lc = garble_nonfree<HASH>(te, type, la, lb, delta, gid, ct);
            if (p.write_ct && act) __stcg(ct_out + (size_t)pos * p.ct_pos_stride, ct);
          }
        } else {
          const uint32_t va = myv[sa * 32u], vb = myv[sb * 32u];
          if (type >= 8) {
            lc = (type == 10) ? la : xor4(la, lb);
          } else {
            const unsigned long long gid = call.gid_base + __shfl_sync(FULL, rec.z, j);
            const unsigned long long cti = call.ct_base + __shfl_sync(FULL, rec.w, j);
            uint4 ct = make_uint4(0, 0, 0, 0);
            if (cti < p.ct_capacity) {
              unsigned long long pos = ct_pos0 + __shfl_sync(FULL, rec.w, j);
              if (p.ct_ring && pos >= p.ct_ring) pos -= p.ct_ring;
              if (act) ct = __ldcg(ct_out + (size_t)pos * p.ct_pos_stride);
            } else if (lane == 0) {
              *p.error_flag = 1u;
            }
            lc = degarble_nonfree<HASH>(te, type, ct, la, va, lb, gid);
          }
          myv[sc * 32u] = (uint8_t)gate_value(type, va, vb);
        }
        my[sc * 32u] = lc;
      }
    }

    // ---- scatter produced labels to the instance group's global slots
    uint4* wlab = p.labels + gbase * 32u + lane;
    uint8_t* wval = p.vals + gbase * 32u + lane;
    for (uint32_t base = 0; base < task.n_out; base += 32) {
      const uint32_t cnt = min(32u, task.n_out - base);
      uint32_t ms = 0, mg = 0;
      if (lane < cnt) {
        ms = p.seq_out_slot[task.out_slot_off + base + lane];
        mg = p.call_slots[call.out_off + base + lane];
      }
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t s = __shfl_sync(FULL, ms, j), gs = __shfl_sync(FULL, mg, j);
        wlab[(size_t)gs * 32u] = my[s * 32u];
        if (MODE == 1) wval[(size_t)gs * 32u] = myv[s * 32u];
      }
    }
    if (p.ct_sys) __threadfence_system();
    else __threadfence();
    __syncwarp();
    kept = sched_complete_warp(p, call_i, grp, lane);
  }
}

// ---- seed expansion: ChaCha20Rng::seed_from_u64(seed) -> delta, constants, input label0s
// (garble_mode.rs:80-97,116-118; rand_core 0.6.4 seed_from_u64; rand_chacha 0.3.1).
// One thread per (instance, 64-byte ChaCha block) = 4 u128 draws.
__device__ __forceinline__ uint32_t rotl(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
#define GSV_QR(a, b, c, d) \
  a += b; d ^= a; d = rotl(d, 16); \
  c += d; b ^= c; b = rotl(b, 12); \
  a += b; d ^= a; d = rotl(d, 8);  \
  c += d; b ^= c; b = rotl(b, 7);

template <int DUMMY>
__global__ void k_seed_expand(const unsigned long long* seeds, uint32_t B, uint32_t G, uint32_t n_inputs,
                              uint32_t n_global_slots, uint4* labels, uint4* delta) {
  const uint32_t n_draws = 3 + n_inputs;
  const uint32_t n_blocks = (n_draws + 3) / 4;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)B * n_blocks) return;
  const uint32_t instance = (uint32_t)(gid / n_blocks);
  const uint32_t blk = (uint32_t)(gid - (size_t)instance * n_blocks);
  // PCG32 expansion of the u64 seed into the 256-bit key
  uint32_t key[8];
  unsigned long long st = seeds[instance];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    st = st * 6364136223846793005ull + 11634580027462260723ull;
    uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
    uint32_t rot = (uint32_t)(st >> 59);
    key[i] = __funnelshift_r(xs, xs, rot);
  }
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                    key[4], key[5], key[6], key[7], blk, 0u, 0u, 0u};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    GSV_QR(x[0], x[4], x[8], x[12])
    GSV_QR(x[1], x[5], x[9], x[13])
    GSV_QR(x[2], x[6], x[10], x[14])
    GSV_QR(x[3], x[7], x[11], x[15])
    GSV_QR(x[0], x[5], x[10], x[15])
    GSV_QR(x[1], x[6], x[11], x[12])
    GSV_QR(x[2], x[7], x[8], x[13])
    GSV_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] += s[i];
  const uint32_t grp = instance / G, lane = instance % G;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t draw = blk * 4 + q;
    if (draw >= n_draws) break;
    // u128 = w0 | w1<<32 | w2<<64 | w3<<96 ; to_bytes() is big-endian
    uint4 v = make_uint4(__byte_perm(x[4 * q + 3], 0, 0x0123), __byte_perm(x[4 * q + 2], 0, 0x0123),
                         __byte_perm(x[4 * q + 1], 0, 0x0123), __byte_perm(x[4 * q + 0], 0, 0x0123));
    if (draw == 0) delta[instance] = v;
    else {
      const uint32_t slot = draw - 1;  // 0 = false, 1 = true, 2.. = inputs
      labels[((size_t)grp * n_global_slots + slot) * G + lane] = v;
    }
  }
}

// ---- stand-alone serial commitment (evaluator side: FileSource hashes what it consumed,
// src/circuit/ciphertext_source.rs:35-106).  One lane per instance.
template <int DUMMY>
__global__ void __launch_bounds__(32) k_chain(const uint4* __restrict__ ct, unsigned long long n_ct, uint32_t B,
                                              uint4* __restrict__ out) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint32_t instance = blockIdx.x * blockDim.x + threadIdx.x;
  if (instance >= B) return;
  uint4 h = make_uint4(0, 0, 0, 0);
  const uint4* p = ct + instance;
  unsigned long long k = 0;
  constexpr int U = 8;
  for (; k + U <= n_ct; k += U) {
    uint4 c[U];
#pragma unroll
    for (int j = 0; j < U; j++) c[j] = __ldcs(p + (size_t)(k + j) * B);
#pragma unroll
    for (int j = 0; j < U; j++) h = aes_fixed(te, xor4(h, c[j]));
  }
  for (; k < n_ct; k++) h = aes_fixed(te, xor4(h, __ldcs(p + (size_t)k * B)));
  out[instance] = h;
}

// ---- small utility kernels
// call pipelining: the constants and the circuit inputs (global slots [0, n) of every instance group) are valid from the start
template <int DUMMY>
__global__ void k_mark_slots(uint32_t* flags, uint32_t n_groups, uint32_t n_global_slots, uint32_t n, uint32_t epoch) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)n_groups * n) flags[(i / n) * n_global_slots + i % n] = epoch;
}

// labels of selected global slots -> dense [instance][j] (and optionally the value bits)
template <int DUMMY>
__global__ void k_gather_slots(const uint4* labels, const uint8_t* vals, const uint32_t* slots, uint32_t n,
                               uint32_t B, uint32_t G, uint32_t n_global_slots, uint4* out, uint8_t* out_vals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * n) return;
  const uint32_t instance = (uint32_t)(i / n), j = (uint32_t)(i % n);
  const size_t gi = ((size_t)(instance / G) * n_global_slots + slots[j]) * G + instance % G;
  out[i] = labels[gi];
  if (out_vals) out_vals[i] = vals[gi];
}
// dense [instance][j] evaluator inputs -> global slots 2.. ; constants into slots 0/1
template <int DUMMY>
__global__ void k_scatter_inputs(const uint4* in_labels, const uint8_t* in_bits, const uint4* true_l,
                                 const uint4* false_l, uint32_t n_inputs, uint32_t B, uint32_t G,
                                 uint32_t n_global_slots, uint4* labels, uint8_t* vals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t per = n_inputs + 2;
  if (i >= (size_t)B * per) return;
  const uint32_t instance = (uint32_t)(i / per), j = (uint32_t)(i % per);
  const size_t gi = ((size_t)(instance / G) * n_global_slots + j) * G + instance % G;
  if (j == 0) { labels[gi] = false_l[instance]; vals[gi] = 0; }
  else if (j == 1) { labels[gi] = true_l[instance]; vals[gi] = 1; }
  else {
    labels[gi] = in_labels[(size_t)instance * n_inputs + (j - 2)];
    vals[gi] = in_bits[(size_t)instance * n_inputs + (j - 2)] ? 1 : 0;
  }
}
// interleaved ct[k][B] <-> one instance's contiguous stream
template <int DUMMY>
__global__ void k_ct_extract(const uint4* ct, uint32_t B, uint32_t instance, unsigned long long first,
                             unsigned long long count, uint4* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = ct[(first + i) * B + instance];
}
template <int DUMMY>
__global__ void k_ct_insert(uint4* ct, uint32_t B, uint32_t instance, unsigned long long first,
                            unsigned long long count, const uint4* in) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) ct[(first + i) * B + instance] = in[i];
}
// commit(label) = AES_K(label)
template <int DUMMY>
__global__ void k_commit_labels(const uint4* in, unsigned long long n, uint4* out) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = aes_fixed(te, in[i]);
}
// H(x_i, gid_i) for the primitive parity tests
template <int HASH>
__global__ void k_hash_blocks(const uint4* x, const unsigned long long* gid, unsigned long long n, uint4* out) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = hash1<HASH>(te, x[i], gid[i]);
}
// register-resident hash throughput probe (the integer-ALU roof): each thread chains
// `per_thread` two-block hashes, like a stream of AND gates with no memory traffic.
template <int HASH>
__global__ void __launch_bounds__(1024, 1) k_bench_hash(unsigned long long per_thread, uint4* sink) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 a = make_uint4((uint32_t)t, 1, 2, 3), b = make_uint4(4, 5, 6, (uint32_t)t);
  for (unsigned long long i = 0; i < per_thread; i++) {
    hash2<HASH>(te, a, b, (unsigned long long)t * per_thread + i);
    a.x ^= b.w;
  }
  if (a.x == 0x12345678u && b.y == 0x9abcdef0u) sink[0] = xor4(a, b);  // defeat DCE
}

// dependent-hash latency probe: every thread chains n one-block gate hashes; thread 0 of block 0 reports
// the cycles (what one barrier-separated level of dependent non-free gates costs at the least)
template <int HASH>
__global__ void k_hash_latency(unsigned long long n, unsigned long long* out) {
  extern __shared__ uint4 smem[];
  uint32_t* te = reinterpret_cast<uint32_t*>(smem);
  load_tables(te, threadIdx.x, blockDim.x);
  __syncthreads();
  uint4 x = make_uint4(threadIdx.x, blockIdx.x, 2, 3);
  const long long t0 = clock64();
  for (unsigned long long i = 0; i < n; i++) x = hash1<HASH>(te, x, i);
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (x.x == 0x12345678u && x.y == 0x9abcdef0u) out[1] = x.z;  // defeat DCE
}

}  // namespace gsvdev

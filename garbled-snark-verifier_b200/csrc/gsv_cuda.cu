// gsv_cuda.cu -- host runtime + C ABI of libgsv_cuda.so (see include/gsv_cuda.h).
// No CPU fallback lives here: every compute entry point needs a CUDA device.
#include "gsv_cuda.h"

#include <cuda_runtime.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/file.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <chrono>
#include <thread>
#include <atomic>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "engine_kernels.cuh"
#include "gadgets.h"
#include "program.h"
#include "host_chain.h"

using namespace gsvdev;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// ---- AES tables for the device (computed, not typed in)
uint8_t gmul(uint8_t a, uint8_t b) {
  uint8_t p = 0;
  for (int i = 0; i < 8; i++) {
    if (b & 1) p ^= a;
    uint8_t hi = a & 0x80;
    a = (uint8_t)(a << 1);
    if (hi) a ^= 0x1b;
    b >>= 1;
  }
  return p;
}
struct AesTables {
  uint32_t te0[256];
  uint32_t rk[44];
  AesTables() {
    uint8_t sbox[256];
    for (int x = 0; x < 256; x++) {
      uint8_t inv = 0;
      if (x)
        for (int y = 1; y < 256; y++)
          if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
      uint8_t s = inv, r = inv;
      for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
      sbox[x] = s ^ 0x63;
    }
    for (int x = 0; x < 256; x++) {
      uint8_t s = sbox[x], s2 = gmul(s, 2), s3 = gmul(s, 3);
      te0[x] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
    }
    // FIPS-197 key schedule of K = 0x42 * 16 (src/hashers/aes_ni.rs:165,179-216)
    static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
    uint8_t k[11][16];
    memset(k[0], 0x42, 16);
    for (int r = 1; r <= 10; r++) {
      const uint8_t* p = k[r - 1];
      uint8_t t[4] = {sbox[p[13]], sbox[p[14]], sbox[p[15]], sbox[p[12]]};
      t[0] ^= rcon[r - 1];
      for (int i = 0; i < 4; i++) k[r][i] = p[i] ^ t[i];
      for (int i = 4; i < 16; i++) k[r][i] = p[i] ^ k[r][i - 4];
    }
    for (int r = 0; r < 11; r++)
      for (int j = 0; j < 4; j++)
        rk[4 * r + j] = (uint32_t)k[r][4 * j] | ((uint32_t)k[r][4 * j + 1] << 8) |
                        ((uint32_t)k[r][4 * j + 2] << 16) | ((uint32_t)k[r][4 * j + 3] << 24);
  }
};

// CUDA loads kernels lazily, and loading one may synchronise the context.  Sessions that share a GPU run
// persistent kernels which wait for each other's progress (linked garbler / evaluator), so a first-use load
// behind a running kernel can dead-lock: every kernel of the library is loaded when a device is first used.
int g_smem_optin = 0;  // the device's opt-in shared memory per block
template <typename K>
void preload(K kernel) {
  cudaFuncAttributes a;
  CUDA_TRY(cudaFuncGetAttributes(&a, kernel));
  // the dynamic shared-memory opt-in is set here, once, for the same reason: changing a function's
  // attributes can reconfigure the SMs' shared-memory carve-out, which waits for running kernels
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin - (int)a.sharedSizeBytes));
}
template <int G>
void preload_engine() {
  preload(k_engine<G, HASH_AES, 0>);
  preload(k_engine<G, HASH_AES, 1>);
  preload(k_engine<G, HASH_BLAKE3, 0>);
  preload(k_engine<G, HASH_BLAKE3, 1>);
  preload(k_engine<G, HASH_AES, 2>);
}
void preload_kernels() {
  preload_engine<1>();
  preload_engine<2>();
  preload_engine<4>();
  preload_engine<8>();
  preload(k_lane<HASH_AES, 0>);
  preload(k_lane<HASH_AES, 1>);
  preload(k_lane<HASH_BLAKE3, 0>);
  preload(k_lane<HASH_BLAKE3, 1>);
  preload(k_sched_init);
  preload(k_mark_slots<0>);
  preload(k_seed_expand<0>);
  preload(k_chain<0>);
  preload(k_gather_slots<0>);
  preload(k_scatter_inputs<0>);
  preload(k_ct_extract<0>);
  preload(k_ct_insert<0>);
  preload(k_commit_labels<0>);
  preload(k_hash_blocks<HASH_AES>);
  preload(k_hash_blocks<HASH_BLAKE3>);
  preload(k_bench_hash<HASH_AES>);
  preload(k_bench_hash<HASH_BLAKE3>);
  preload(k_hash_latency<HASH_AES>);
  preload(k_hash_latency<HASH_BLAKE3>);
}

bool g_tables_loaded[64] = {false};
void ensure_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    throw std::runtime_error("no CUDA device visible (libgsv_cuda has no CPU fallback)");
  if (device < 0 || device >= n) throw std::runtime_error("bad device ordinal");
  CUDA_TRY(cudaSetDevice(device));
  if (!g_tables_loaded[device]) {
    static const AesTables T;
    CUDA_TRY(cudaMemcpyToSymbol(c_te0, T.te0, sizeof(T.te0)));
    CUDA_TRY(cudaMemcpyToSymbol(c_rk, T.rk, sizeof(T.rk)));
    CUDA_TRY(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    preload_kernels();
    g_tables_loaded[device] = true;
  }
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    n = count;
    if (count) CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const std::vector<T>& v) {
    alloc(v.size());
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

}  // namespace

// =============================================================================== program
struct gsv_program {
  std::unique_ptr<gsv::Builder> builder;
  uint32_t root = 0;
  gsv::Program prog;
  uint32_t max_task_levels = 0;
  uint64_t sum_call_levels = 0;
  uint64_t critical_path_gates = 0;   // longest dependency chain through the calls, weighted by gates
  uint64_t critical_path_levels = 0;  // ... weighted by the tasks' level counts
};

struct gsv_ctx {
  gsv::Builder* b;
};

namespace {

// ---- plan cache.  Planning the verifier takes ~23 s and ~17 GB of host memory per process; ranks of one job (and
// successive runs on one box) share the result through a file: GSV_PLAN_CACHE_DIR/<circuit>.<options>.plan, written
// once under an advisory lock.  A program loaded from the cache has no recorded circuit (template DAG): the
// host-side walkers (gsv_program_execute, _flat_stream, _export_templates, _depth) refuse it.
constexpr uint64_t PLAN_MAGIC = 0x67737650'4c414e32ull;  // "gsvPLAN2"
struct PlanWriter {
  FILE* f;
  bool ok = true;
  void raw(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; }
  template <typename T> void pod(const T& v) { raw(&v, sizeof(T)); }
  template <typename T> void vec(const std::vector<T>& v) { pod<uint64_t>(v.size()); raw(v.data(), v.size() * sizeof(T)); }
  void str(const std::string& v) { pod<uint64_t>(v.size()); raw(v.data(), v.size()); }
};
struct PlanReader {
  FILE* f;
  bool ok = true;
  void raw(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) ok = false; }
  template <typename T> void pod(T& v) { raw(&v, sizeof(T)); }
  template <typename T> void vec(std::vector<T>& v) {
    uint64_t n = 0;
    pod(n);
    if (!ok || n > (1ull << 36) / sizeof(T)) { ok = false; return; }
    v.resize(n);
    raw(v.data(), n * sizeof(T));
  }
  void str(std::string& v) {
    uint64_t n = 0;
    pod(n);
    if (!ok || n > (1u << 20)) { ok = false; return; }
    v.resize(n);
    raw(&v[0], n);
  }
};
template <typename IO, typename P>
void plan_io(IO& io, P& g) {  // the same field walk for writing (const-cast) and reading
  uint64_t nt = g.tasks.size();
  io.pod(nt);
  g.tasks.resize(nt);
  for (auto& t : g.tasks) {
    io.str(t.key);
    io.pod(t.n_in); io.pod(t.n_out); io.pod(t.n_slots); io.pod(t.n_levels); io.pod(t.max_width);
    io.pod(t.n_gates_total); io.pod(t.n_ct); io.pod(t.n_live);
    io.vec(t.gates); io.vec(t.level_off); io.vec(t.in_slot); io.vec(t.out_slot); io.vec(t.out_pos);
    io.vec(t.seq_gates); io.vec(t.seq_in_slot); io.vec(t.seq_out_slot); io.pod(t.n_seq_slots);
    io.vec(t.pipe_in_need); io.vec(t.pipe_out_ready); io.pod(t.pipe_depth);
    io.pod(t.window_levels); io.vec(t.win_in); io.vec(t.win_out); io.vec(t.win_in_off); io.vec(t.win_out_off);
    io.vec(t.in_need_level); io.vec(t.out_ready_level);
  }
  io.vec(g.calls); io.vec(g.call_slots); io.vec(g.deps);
  io.pod(g.n_global_slots); io.pod(g.n_inputs); io.vec(g.output_slots);
  io.pod(g.total_gates); io.pod(g.total_ct); io.pod(g.total_live);
  io.pod(g.max_task_slots); io.pod(g.max_task_seq_slots); io.pod(g.has_levelised); io.pod(g.max_task_in);
  io.pod(g.max_call_deps); io.pod(g.pipelined);
  for (auto& c : g.type_count) io.pod(c);
}
void derive_program_stats(gsv_program* p) {
  p->max_task_levels = 0;
  p->sum_call_levels = p->critical_path_gates = p->critical_path_levels = 0;
  for (const auto& t : p->prog.tasks) p->max_task_levels = std::max(p->max_task_levels, t.n_levels);
  for (const auto& c : p->prog.calls) p->sum_call_levels += p->prog.tasks[c.task].n_levels;
  // dependency-chain lengths: what bounds one instance's latency however many SMs there are.  Levels: per-wire
  // model (pipelined consumers need an input at the start of the window that first reads it, producers publish an
  // output at the end of the window that completes it, + 2 levels of polling; otherwise at the call's end); done
  // dependencies always wait for the end.
  const auto& g = p->prog;
  std::vector<uint64_t> fg(g.calls.size()), fl(g.calls.size());
  std::vector<uint32_t> st(g.n_global_slots, 0);
  for (size_t i = 0; i < g.calls.size(); i++) {
    const auto& c = g.calls[i];
    const gsv::Task& t = g.tasks[c.task];
    uint64_t sg = 0, S = 0;
    for (uint32_t d = 0; d < c.n_deps; d++) {
      sg = std::max(sg, fg[g.deps[c.dep_off + d]]);
      if (d >= c.n_start_deps) S = std::max(S, fl[g.deps[c.dep_off + d]]);
    }
    if (g.pipelined) {
      for (uint32_t k = 0; k < t.n_in; k++) {
        if (k >= t.in_need_level.size() || t.in_need_level[k] == 0xFFFFFFFFu) continue;
        const uint64_t need = t.in_need_level[k] / t.window_levels * (uint64_t)t.window_levels;
        const uint64_t avail = (uint64_t)st[g.call_slots[c.in_off + k]] + 2;
        if (avail > need) S = std::max(S, avail - need);
      }
      for (uint32_t k = 0; k < t.n_out; k++) {
        uint64_t r = t.n_levels;
        if (k < t.out_ready_level.size()) r = std::min<uint64_t>(r, (t.out_ready_level[k] / t.window_levels + 1) * (uint64_t)t.window_levels);
        st[g.call_slots[c.out_off + k]] = (uint32_t)(S + r);
      }
    } else {
      for (uint32_t d = 0; d < c.n_start_deps; d++) S = std::max(S, fl[g.deps[c.dep_off + d]]);
    }
    fg[i] = sg + t.n_gates_total;
    fl[i] = S + t.n_levels;
    p->critical_path_gates = std::max(p->critical_path_gates, fg[i]);
    p->critical_path_levels = std::max(p->critical_path_levels, fl[i]);
  }
}
// gsv_plan_options.pipeline: 1 on, 2 off, 0 = default (on; GSV_PIPELINE=0 in the environment turns it off)
uint32_t pipeline_option(const gsv_plan_options* opt) {
  if (opt && opt->pipeline) return opt->pipeline == 1 ? 1u : 2u;
  const char* e = getenv("GSV_PIPELINE");
  return (e && atoi(e) == 0) ? 2u : 1u;
}

std::string plan_cache_path(const std::string& circuit, const gsv_plan_options* opt) {
  const char* dir = getenv("GSV_PLAN_CACHE_DIR");
  if (!dir || !*dir) return "";
  // the file name carries the plan options and this library build's stamp: a rebuilt library never reads old plans
  uint32_t stamp = 2166136261u;
  for (const char* c = __DATE__ " " __TIME__; *c; c++) stamp = (stamp ^ (uint8_t)*c) * 16777619u;
  char buf[160];
  const uint32_t pl = pipeline_option(opt);
  snprintf(buf, sizeof buf, ".g%llu.s%u.l%u.p%u.w%u.%08x.plan", opt ? (unsigned long long)opt->max_task_gates : 0ull,
           opt ? opt->max_task_slots : 0u, opt ? opt->lane_only : 0u, pl, opt ? opt->window_levels : 0u, stamp);
  std::string name = circuit;
  for (char& c : name)
    if (!isalnum((unsigned char)c) && c != '_') c = '_';
  return std::string(dir) + "/" + name + buf;
}
gsv_program* load_plan(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return nullptr;
  PlanReader r{f};
  uint64_t magic = 0;
  r.pod(magic);
  auto p = std::make_unique<gsv_program>();
  if (r.ok && magic == PLAN_MAGIC) plan_io(r, p->prog);
  uint64_t tail = 0;
  r.pod(tail);
  fclose(f);
  if (!r.ok || magic != PLAN_MAGIC || tail != PLAN_MAGIC) return nullptr;  // truncated / foreign file: plan afresh
  derive_program_stats(p.get());
  return p.release();
}
void store_plan(const std::string& path, gsv_program* p) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  PlanWriter w{f};
  w.pod(PLAN_MAGIC);
  plan_io(w, p->prog);
  w.pod(PLAN_MAGIC);
  const bool ok = w.ok && fclose(f) == 0;
  if (ok) rename(tmp.c_str(), path.c_str());
  else remove(tmp.c_str());
}

gsv_program* finish_program(std::unique_ptr<gsv::Builder> b, uint32_t root, const gsv_plan_options* opt) {
  gsv::PlanOptions po;
  if (opt) {
    if (opt->max_task_gates) po.max_task_gates = opt->max_task_gates;
    if (opt->max_task_slots) po.max_task_slots = opt->max_task_slots;
    if (opt->lane_only) {
      po.build_levelised = false;
      po.max_global_slots = 1u << 20;  // lane mode serves thousands of instances: memory before critical path
    }
    if (opt->window_levels) po.window_levels = opt->window_levels;
  }
  po.pipeline = pipeline_option(opt) == 1 && po.build_levelised;
  auto p = std::make_unique<gsv_program>();
  p->prog = gsv::plan_program(*b, root, po);
  p->builder = std::move(b);
  p->root = root;
  if (getenv("GSV_PLAN_DEBUG")) {  // largest working sets, for choosing max_task_slots
    std::vector<const gsv::Task*> ts;
    for (const auto& t : p->prog.tasks) ts.push_back(&t);
    std::sort(ts.begin(), ts.end(), [](const gsv::Task* a, const gsv::Task* b) { return a->n_slots > b->n_slots; });
    for (size_t i = 0; i < ts.size() && i < 12; i++)
      fprintf(stderr, "[plan] slots %u levels %u gates %llu in %u out %u  %s\n", ts[i]->n_slots, ts[i]->n_levels,
              (unsigned long long)ts[i]->n_gates_total, ts[i]->n_in, ts[i]->n_out, ts[i]->key.substr(0, 80).c_str());
  }
  derive_program_stats(p.get());
  if (getenv("GSV_PLAN_DEBUG") && p->prog.has_levelised) {
    {  // worker time (levels, AES gates) by the shared-memory footprint of the task: what smaller workers could take
      const uint32_t edges[] = {128, 256, 384, 512, 640, 768, 1024, 1280, 1536, 2048, 0xFFFFFFFFu};
      uint64_t lv[11] = {0}, ct[11] = {0}, n[11] = {0};
      for (const auto& c : p->prog.calls) {
        const gsv::Task& t = p->prog.tasks[c.task];
        int b = 0;
        while (t.n_slots > edges[b]) b++;
        lv[b] += t.n_levels;
        ct[b] += t.n_ct;
        n[b]++;
      }
      for (int b = 0; b < 11; b++)
        fprintf(stderr, "[plan] tasks with <= %u slots: %llu calls, %llu levels, %llu ciphertexts\n", edges[b],
                (unsigned long long)n[b], (unsigned long long)lv[b], (unsigned long long)ct[b]);
    }
    // what the critical path would be if a consumer level only waited for the producer LEVELS it needs
    const auto& g = p->prog;
    std::vector<uint64_t> slot_time(g.n_global_slots, 0);
    uint64_t crit = 0, crit_plain = 0;
    std::vector<uint64_t> slot_end(g.n_global_slots, 0);
    for (const auto& c : g.calls) {
      const gsv::Task& t = g.tasks[c.task];
      uint64_t S = 0, S_plain = 0;
      for (uint32_t i = 0; i < t.n_in; i++) {
        if (t.in_slot[i] == 0xFFFF) continue;
        const uint32_t sl = g.call_slots[c.in_off + i];
        const uint64_t need = i < t.pipe_in_need.size() ? t.pipe_in_need[i] : 1;
        if (slot_time[sl] + 1 > need) S = std::max(S, slot_time[sl] + 1 - need);
        S_plain = std::max(S_plain, slot_end[sl]);
      }
      for (uint32_t k = 0; k < t.n_out; k++) {
        const uint32_t sl = g.call_slots[c.out_off + k];
        slot_time[sl] = S + (k < t.pipe_out_ready.size() ? t.pipe_out_ready[k] : t.pipe_depth);
        slot_end[sl] = S_plain + t.pipe_depth;
      }
      crit = std::max(crit, S + t.pipe_depth);
      crit_plain = std::max(crit_plain, S_plain + t.pipe_depth);
    }
    // the same call-granular chain over the explicit edges (RAW + WAR), true depths vs device levels
    std::vector<uint64_t> f1(g.calls.size()), f2(g.calls.size());
    uint64_t c1 = 0, c2 = 0, sum_true = 0, sum_dev = 0;
    for (size_t i = 0; i < g.calls.size(); i++) {
      const auto& c = g.calls[i];
      uint64_t a = 0, b2 = 0;
      for (uint32_t d = 0; d < c.n_deps; d++) {
        a = std::max(a, f1[g.deps[c.dep_off + d]]);
        b2 = std::max(b2, f2[g.deps[c.dep_off + d]]);
      }
      f1[i] = a + g.tasks[c.task].pipe_depth;
      f2[i] = b2 + g.tasks[c.task].n_levels;
      c1 = std::max(c1, f1[i]);
      c2 = std::max(c2, f2[i]);
      sum_true += g.tasks[c.task].pipe_depth;
      sum_dev += g.tasks[c.task].n_levels;
    }
    for (uint64_t W : {16ull, 32ull, 64ull, 128ull}) {
      // the same with windows of W levels: inputs gathered at the start of the window that first reads them, outputs
      // published at the end of the window that completes them (+ 2 levels of polling latency)
      std::vector<uint64_t> st(g.n_global_slots, 0);
      uint64_t cw = 0;
      for (const auto& c : g.calls) {
        const gsv::Task& t = g.tasks[c.task];
        uint64_t S = 0;
        for (uint32_t i = 0; i < t.n_in; i++) {
          if (t.in_slot[i] == 0xFFFF) continue;
          const uint32_t sl = g.call_slots[c.in_off + i];
          const uint64_t need = (i < t.pipe_in_need.size() ? t.pipe_in_need[i] : 1);
          const uint64_t need_w = (need ? (need - 1) / W * W : 0);  // level at whose start the input must be there
          if (st[sl] + 2 > need_w) S = std::max(S, st[sl] + 2 - need_w);
        }
        for (uint32_t k = 0; k < t.n_out; k++) {
          const uint32_t sl = g.call_slots[c.out_off + k];
          const uint64_t rdy = k < t.pipe_out_ready.size() ? t.pipe_out_ready[k] : t.pipe_depth;
          st[sl] = S + std::min<uint64_t>((rdy + W - 1) / W * W, t.pipe_depth);
        }
        cw = std::max(cw, S + t.pipe_depth);
      }
      fprintf(stderr, "[plan] windowed pipelining, W = %llu levels: critical path %llu\n", (unsigned long long)W, (unsigned long long)cw);
    }
    fprintf(stderr, "[plan] critical path in levels: RAW only / true depths %llu, RAW+WAR / true depths %llu, "
            "RAW+WAR / device levels (128-gate cap) %llu; level-pipelined %llu; sum of call levels true %llu device %llu\n",
            (unsigned long long)crit_plain, (unsigned long long)c1, (unsigned long long)c2, (unsigned long long)crit,
            (unsigned long long)sum_true, (unsigned long long)sum_dev);
  }
  return p.release();
}
}  // namespace

extern "C" {

const char* gsv_last_error(void) { return g_err.c_str(); }
const char* gsv_version(void) { return "gsv-cuda 0.1 (sm_100a)"; }
int gsv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

uint32_t gsv_ctx_issue_wire(gsv_ctx* ctx) { return ctx->b->issue_wire(); }
void gsv_ctx_add_gate(gsv_ctx* ctx, int gate_type, uint32_t a, uint32_t b, uint32_t c) {
  ctx->b->add_gate((uint8_t)gate_type, a, b, c);
}
void gsv_ctx_component(gsv_ctx* ctx, const char* key, const uint32_t* inputs, uint32_t n_in, uint32_t arity,
                       gsv_body_fn body, void* user, uint32_t* outputs) {
  gsv::Wires in(inputs, inputs + n_in);
  gsv::Wires out = ctx->b->component(key, in, arity, [body, user, arity](gsv::Builder& b, const gsv::Wires& ins) {
    gsv_ctx c{&b};
    gsv::Wires o(arity);
    body(&c, user, ins.data(), (uint32_t)ins.size(), o.data(), arity);
    return o;
  });
  std::copy(out.begin(), out.end(), outputs);
}

gsv_program* gsv_program_record(const char* name, uint32_t n_inputs, uint32_t n_outputs, gsv_body_fn root,
                                void* user, const gsv_plan_options* opt) {
  try {
    auto b = std::make_unique<gsv::Builder>();
    uint32_t r = b->build_root(name ? name : "root", n_inputs, [root, user, n_outputs](gsv::Builder& bb, const gsv::Wires& ins) {
      gsv_ctx c{&bb};
      gsv::Wires o(n_outputs);
      root(&c, user, ins.data(), (uint32_t)ins.size(), o.data(), n_outputs);
      return o;
    });
    return finish_program(std::move(b), r, opt);
  } catch (const std::exception& e) {
    fail(GSV_ERR_INVALID, e.what());
    return nullptr;
  }
}

gsv_program* gsv_program_build(const char* circuit, const gsv_plan_options* opt) {
  try {
    std::string c = circuit ? circuit : "";
    const std::string cache = plan_cache_path(c, opt);
    int lock_fd = -1;
    if (!cache.empty()) {
      if (gsv_program* hit = load_plan(cache)) return hit;
      // one process plans, the others wait for its file
      lock_fd = open((cache + ".lock").c_str(), O_CREAT | O_RDWR, 0644);
      if (lock_fd >= 0) flock(lock_fd, LOCK_EX);
      if (gsv_program* hit = load_plan(cache)) {
        if (lock_fd >= 0) close(lock_fd);
        return hit;
      }
    }
    auto b = std::make_unique<gsv::Builder>();
    gsv_program* p = nullptr;
    try {
      const uint32_t r = gsv::build_named_circuit(*b, c);
      p = finish_program(std::move(b), r, opt);
      if (p && !cache.empty()) store_plan(cache, p);
    } catch (...) {
      if (lock_fd >= 0) close(lock_fd);
      throw;
    }
    if (lock_fd >= 0) close(lock_fd);
    return p;
  } catch (const std::exception& e) {
    fail(GSV_ERR_INVALID, e.what());
    return nullptr;
  }
}

void gsv_program_destroy(gsv_program* p) { delete p; }

int gsv_program_execute(const gsv_program* p, const uint8_t* input_bits, uint8_t* output_bits, uint64_t* gates_executed) {
  if (!p || !input_bits || !output_bits) return fail(GSV_ERR_INVALID, "null argument");
  if (!p->builder) return fail(GSV_ERR_INVALID, "program came from the plan cache: no recorded circuit to walk");
  try {
    const gsv::Template& rt = p->builder->tmpl(p->root);
    std::vector<uint8_t> in(input_bits, input_bits + rt.n_in);
    std::vector<uint8_t> out = gsv::execute(*p->builder, p->root, in, gates_executed);
    memcpy(output_bits, out.data(), out.size());
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_INVALID, e.what());
  }
}

int gsv_program_export_templates(const gsv_program* p, uint64_t sizes[6], uint32_t* root, uint32_t* tmpl,
                                 uint32_t* gates, uint32_t* calls, uint32_t* items, uint32_t* call_wires,
                                 uint32_t* outs) {
  if (!p || !sizes) return fail(GSV_ERR_INVALID, "null argument");
  if (!p->builder) return fail(GSV_ERR_INVALID, "program came from the plan cache: no recorded circuit to export");
  if (!gsv::export_templates(*p->builder, sizes, tmpl, gates, calls, items, call_wires, outs))
    return fail(GSV_ERR_INVALID, "template DAG too large to export");
  if (root) *root = p->root;
  return GSV_OK;
}

int gsv_host_chain_fold(uint8_t* h, const uint8_t* base, uint64_t pos_stride, uint64_t inst_stride, uint64_t n_pos, uint32_t n_inst) {
  if (!h || (!base && n_pos)) return fail(GSV_ERR_INVALID, "null argument");
  if (!gsv::host_chain_available()) return fail(GSV_ERR_INVALID, "host CPU has no AES-NI");
  gsv::host_chain_fold(h, base, pos_stride, inst_stride, n_pos, n_inst);
  return GSV_OK;
}

int gsv_host_chain_fold_quads(uint8_t* h, const uint8_t* base, uint64_t quad_bytes, uint64_t n_pos, uint32_t n_quads) {
  if (!h || (!base && n_pos)) return fail(GSV_ERR_INVALID, "null argument");
  if (!gsv::host_chain_available()) return fail(GSV_ERR_INVALID, "host CPU has no AES-NI");
  gsv::host_chain_fold_quads(h, base, quad_bytes, n_pos, n_quads);
  return GSV_OK;
}

int gsv_program_depth(const gsv_program* p, uint64_t* depth_all, uint64_t* depth_nonfree) {
  if (!p) return fail(GSV_ERR_INVALID, "null argument");
  if (!p->builder) return fail(GSV_ERR_INVALID, "program came from the plan cache: no recorded circuit");
  try {
    gsv::circuit_depth(*p->builder, p->root, depth_all, depth_nonfree);
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_INVALID, e.what());
  }
}

int gsv_program_execute_plan(const gsv_program* p, int lane_form, const uint8_t* input_bits, uint8_t* output_bits) {
  if (!p || !input_bits || !output_bits) return fail(GSV_ERR_INVALID, "null argument");
  try {
    const gsv::Program& g = p->prog;
    if (!lane_form && !g.has_levelised) throw std::runtime_error("program was planned lane-only");
    auto eval = [](uint8_t type, uint8_t a, uint8_t b) -> uint8_t {
      if (type == gsv::NOT) return a ^ 1;
      if (type == gsv::XOR) return a ^ b;
      if (type == gsv::XNOR) return a ^ b ^ 1;
      const uint8_t aa = (type >> 2) & 1, ab = (type >> 1) & 1, ac = type & 1;  // gate_type.rs alphas
      return ((a ^ aa) & (b ^ ab)) ^ ac;
    };
    // 0xFF marks a slot nobody has written in this pass: reading one is a planner bug
    std::vector<uint8_t> glob(g.n_global_slots, 0xFF), loc;
    glob[0] = 0;
    glob[1] = 1;
    for (uint32_t i = 0; i < g.n_inputs; i++) glob[2 + i] = input_bits[i] & 1;
    // Dependency audit: under the dataflow scheduler only the explicit edges order the calls, so every
    // read must list the slot's last writer (RAW) and every write the slot's readers since then, or the
    // last writer if nobody read it (WAR / WAW), directly in the call's dependency list.
    std::vector<int64_t> last_writer(g.n_global_slots, -1);
    std::vector<std::vector<uint32_t>> readers_since(g.n_global_slots);
    // the dependency list is two sorted ranges: start dependencies, then done dependencies
    auto has_done_dep = [&](const gsv::Call& c, uint32_t d) {
      return std::binary_search(g.deps.begin() + c.dep_off + c.n_start_deps, g.deps.begin() + c.dep_off + c.n_deps, d);
    };
    auto has_dep = [&](const gsv::Call& c, uint32_t d) {
      return std::binary_search(g.deps.begin() + c.dep_off, g.deps.begin() + c.dep_off + c.n_start_deps, d) || has_done_dep(c, d);
    };
    for (size_t ci = 0; ci < g.calls.size(); ci++) {
      const gsv::Call& c = g.calls[ci];
      const gsv::Task& t = g.tasks[c.task];
      for (uint32_t i = 0; i < t.n_in; i++) {
        if ((lane_form ? t.seq_in_slot[i] : t.in_slot[i]) == 0xFFFF) continue;
        const uint32_t sl = g.call_slots[c.in_off + i];
        if (last_writer[sl] >= 0 && !has_dep(c, (uint32_t)last_writer[sl]))
          throw std::runtime_error("plan misses a RAW edge: call " + std::to_string(ci) + " reads slot " +
                                   std::to_string(sl) + " of call " + std::to_string(last_writer[sl]));
        if (readers_since[sl].empty() || readers_since[sl].back() != (uint32_t)ci) readers_since[sl].push_back((uint32_t)ci);
      }
      for (uint32_t k = 0; k < t.n_out; k++) {
        const uint32_t sl = g.call_slots[c.out_off + k];
        for (uint32_t r : readers_since[sl])
          if (r != (uint32_t)ci && !has_done_dep(c, r))
            throw std::runtime_error("plan misses a WAR edge (must be a done dependency): call " + std::to_string(ci) + " overwrites slot " +
                                     std::to_string(sl) + " read by call " + std::to_string(r));
        if (readers_since[sl].empty() && last_writer[sl] >= 0 && last_writer[sl] != (int64_t)ci &&
            !has_done_dep(c, (uint32_t)last_writer[sl]))
          throw std::runtime_error("plan misses a WAW edge on slot " + std::to_string(sl));
      }
      for (uint32_t k = 0; k < t.n_out; k++) {
        const uint32_t sl = g.call_slots[c.out_off + k];
        last_writer[sl] = (int64_t)ci;
        readers_since[sl].clear();
      }
      const auto& gates = lane_form ? t.seq_gates : t.gates;
      const auto& in_slot = lane_form ? t.seq_in_slot : t.in_slot;
      const auto& out_slot = lane_form ? t.seq_out_slot : t.out_slot;
      loc.assign(std::max<uint32_t>(lane_form ? t.n_seq_slots : t.n_slots, 2), 0xFF);
      loc[0] = 0;
      loc[1] = 1;
      auto run = [&](const gsv::DevGate& dg) {
        const uint8_t a = loc[dg.a], b = loc[dg.type == gsv::NOT ? dg.a : dg.b];
        if ((a | b) > 1) throw std::runtime_error("plan reads an unwritten slot in call " + std::to_string(ci) + " (" + t.key + ")");
        loc[dg.c] = eval(dg.type, a, b);
      };
      if (lane_form) {
        for (uint32_t i = 0; i < t.n_in; i++)
          if (in_slot[i] != 0xFFFF) loc[in_slot[i]] = glob[g.call_slots[c.in_off + i]];
        for (const gsv::DevGate& dg : gates) run(dg);
        for (uint32_t k = 0; k < t.n_out; k++) glob[g.call_slots[c.out_off + k]] = loc[out_slot[k]];
      } else {
        // exactly what the kernel does: window by window, inputs gathered when their window starts, outputs
        // published when their window ends (one window for plans without pipelining)
        const uint32_t n_win = (uint32_t)t.win_in_off.size() - 1;
        size_t n_gathered = 0, n_published = 0;
        for (uint32_t w = 0; w < n_win; w++) {
          for (uint32_t q = t.win_in_off[w]; q < t.win_in_off[w + 1]; q++, n_gathered++) {
            const uint32_t i = t.win_in[q];
            loc[in_slot[i]] = glob[g.call_slots[c.in_off + i]];
          }
          const uint64_t l0 = (uint64_t)w * t.window_levels, l1 = std::min<uint64_t>(t.n_levels, l0 + t.window_levels);
          for (uint64_t l = l0; l < l1; l++)
            for (uint32_t k = t.level_off[l]; k < t.level_off[l + 1]; k++) run(gates[k]);
          for (uint32_t q = t.win_out_off[w]; q < t.win_out_off[w + 1]; q++, n_published++) {
            const uint32_t k = t.win_out[q];
            if (loc[out_slot[k]] > 1) throw std::runtime_error("plan publishes an output before it is written: " + t.key);
            glob[g.call_slots[c.out_off + k]] = loc[out_slot[k]];
          }
        }
        size_t n_read = 0;
        for (uint32_t i = 0; i < t.n_in; i++) n_read += in_slot[i] != 0xFFFF;
        if (n_gathered != n_read || n_published != t.n_out) throw std::runtime_error("window tables incomplete in " + t.key);
        for (uint32_t k = 0; k < t.n_out; k++)
          if (glob[g.call_slots[c.out_off + k]] != loc[out_slot[k]])
            throw std::runtime_error("an output slot of " + t.key + " changed after it was published");
      }
    }
    for (size_t j = 0; j < g.output_slots.size(); j++) {
      if (glob[g.output_slots[j]] > 1) throw std::runtime_error("circuit output slot never written");
      output_bits[j] = glob[g.output_slots[j]];
    }
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_INVALID, e.what());
  }
}

int gsv_groth16_synthetic_inputs(uint64_t public_x, int flip_public, uint8_t* bits, uint32_t n_bits) {
  // 1273 bits: groth16_verify_compressed (x + sign flag per point); 2286 bits: groth16_verify (affine x, y)
  if (!bits || (n_bits != 1273 && n_bits != 2286)) return fail(GSV_ERR_INVALID, "need a 1273-bit (compressed) or 2286-bit buffer");
  try {
    using namespace gsv::host;
    VerifyingKey vk;
    Proof pr;
    synthetic_groth16(7, gsv::U256(public_x), vk, pr);
    size_t o = 0;
    auto put = [&](const gsv::U256& v) { for (unsigned i = 0; i < 254; i++) bits[o++] = v.bit(i); };
    put(gsv::U256(flip_public ? public_x + 1 : public_x));
    if (n_bits == 2286) {
      put(gsv::mont254(pr.a.x.to_u256())); put(gsv::mont254(pr.a.y.to_u256()));
      put(gsv::mont254(pr.b.x.c0.to_u256())); put(gsv::mont254(pr.b.x.c1.to_u256()));
      put(gsv::mont254(pr.b.y.c0.to_u256())); put(gsv::mont254(pr.b.y.c1.to_u256()));
      put(gsv::mont254(pr.c.x.to_u256())); put(gsv::mont254(pr.c.y.to_u256()));
      return GSV_OK;
    }
    const gsv::U256 e_sqrt = gsv::U256::from_dec("5472060717959818805561601436314318772174077789324455915672259473661306552146");
    const gsv::U256& pm = FpCtx::get().p;
    auto g1flag = [&](const G1Affine& a) { return fp_pow(a.x * a.x * a.x + Fp::from_u64(3), e_sqrt) == a.y; };
    // mirror of Fq2::sqrt_general_montgomery (fq2.rs:425-447) to learn which root the circuit picks
    Fp2 y2 = pr.b.x * pr.b.x * pr.b.x + Params::get().g2_b;
    Fp alpha_sqrt = fp_pow(y2.c0 * y2.c0 + y2.c1 * y2.c1, e_sqrt);
    Fp half = fp_inv(Fp::from_u64(2));
    Fp delta = (alpha_sqrt + y2.c0) * half;
    bool is_qnr = fp_pow(delta, gsv::shr1(gsv::sub(pm, gsv::U256(1)))) == -Fp::from_u64(1);
    Fp delta_final = is_qnr ? delta - alpha_sqrt : delta;
    Fp c0 = fp_pow(delta_final, e_sqrt);
    Fp c1 = fp_inv(c0) * (y2.c1 * half);
    bool bflag = (c0 == pr.b.y.c0) && (c1 == pr.b.y.c1);
    put(gsv::mont254(pr.a.x.to_u256())); bits[o++] = g1flag(pr.a);
    put(gsv::mont254(pr.b.x.c0.to_u256())); put(gsv::mont254(pr.b.x.c1.to_u256())); bits[o++] = bflag;
    put(gsv::mont254(pr.c.x.to_u256())); bits[o++] = g1flag(pr.c);
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_INVALID, e.what());
  }
}

int gsv_program_get_info(const gsv_program* p, gsv_program_info* out) {
  if (!p || !out) return fail(GSV_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  const gsv::Program& g = p->prog;
  out->n_gates = g.total_gates;
  out->n_live_gates = g.total_live;
  out->n_ciphertexts = g.total_ct;
  for (int i = 0; i < 11; i++) out->type_count[i] = g.type_count[i];
  out->n_inputs = g.n_inputs;
  out->n_outputs = (uint32_t)g.output_slots.size();
  out->n_tasks = (uint32_t)g.tasks.size();
  out->n_calls = (uint32_t)g.calls.size();
  out->n_global_slots = g.n_global_slots;
  out->max_task_slots = g.max_task_slots;
  out->max_task_levels = p->max_task_levels;
  out->max_call_deps = g.max_call_deps;
  out->sum_call_levels = p->sum_call_levels;
  out->critical_path_gates = p->critical_path_gates;
  out->critical_path_levels = p->critical_path_levels;
  return GSV_OK;
}

int64_t gsv_program_flat_stream(const gsv_program* p, uint8_t* type, uint32_t* a, uint32_t* b, uint32_t* c,
                                uint64_t capacity, uint32_t* outputs, uint32_t* n_wires) {
  if (!p) return fail(GSV_ERR_INVALID, "null program");
  if (!p->builder) return fail(GSV_ERR_INVALID, "program came from the plan cache: no recorded circuit to flatten");
  try {
    const gsv::Template& rt = p->builder->tmpl(p->root);
    if (!type) {
      if (n_wires) *n_wires = 0;
      return (int64_t)rt.total_gates;
    }
    if (capacity < rt.total_gates) return fail(GSV_ERR_CAPACITY, "flat stream buffers too small");
    gsv::FlatStream fs = gsv::flatten(*p->builder, p->root, rt.total_gates + 1);
    memcpy(type, fs.type.data(), fs.type.size());
    memcpy(a, fs.a.data(), fs.a.size() * 4);
    memcpy(b, fs.b.data(), fs.b.size() * 4);
    memcpy(c, fs.c.data(), fs.c.size() * 4);
    if (outputs) memcpy(outputs, fs.outputs.data(), fs.outputs.size() * 4);
    if (n_wires) *n_wires = fs.n_wires;
    return (int64_t)fs.type.size();
  } catch (const std::exception& e) {
    return fail(GSV_ERR_INVALID, e.what());
  }
}

}  // extern "C"

// =============================================================================== session
// Garbler -> evaluator streaming (gsv_session_link): a ciphertext ring in the evaluator's device memory that the
// garbler's kernel fills (peer stores over NVLink when the sessions sit on two GPUs) and three progress words in
// mapped host memory, each in its own cache line.
struct gsv_link {
  int dev_g = 0, dev_e = 0;
  uint4* ring = nullptr;            // on dev_e, layout [instance quad][position][4]
  uint64_t ring_positions = 0;      // 0: the whole stream fits (no wrap)
  uint64_t cap = 0;                 // positions per instance in the buffer
  unsigned long long* words = nullptr;  // mapped + portable host memory, 4 x 64 bytes:
  //   words[0]  stream positions the garbler has completed        (garbler's publisher warp writes)
  //   words[8]  stream positions the evaluate kernel has consumed (evaluator's publisher warp writes)
  //   words[16] stream positions the garbler may overwrite        (evaluator's host thread writes)
  unsigned long long *ready_g = nullptr, *release_g = nullptr;   // device views for the garbler's GPU
  unsigned long long *ready_e = nullptr, *done_e = nullptr;      // device views for the evaluator's GPU
  std::atomic<uint64_t> runs_started{0}, runs_consumed{0};
  std::atomic<bool> failed{false};
  ~gsv_link() {
    if (ring) {
      cudaSetDevice(dev_e);
      cudaFree(ring);
    }
    if (words) cudaFreeHost(words);
  }
};

struct gsv_session {
  const gsv_program* prog = nullptr;
  int device = 0;
  uint32_t B = 0, G = 1, NT = 256, n_workers = 4, n_groups = 0;
  uint32_t slots_per_worker = 0;
  uint32_t ct_mode = GSV_CT_COMMIT;
  int sm_count = 0;
  size_t smem_garble = 0, smem_eval = 0;
  uint32_t n_chain_warps = 0;      // chain warps per chain CTA
  uint32_t n_chain_ctas = 0;       // trailing CTAs dedicated to the commitment chain
  uint32_t host_threads = 1;       // GSV_CT_COMMIT_HOST: fold threads
  uint64_t ct_ring = 0;  // ring capacity in ciphertexts (0 = whole stream kept)
  uint32_t epoch = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // program on device
  DevBuf<uint4> d_gates;
  DevBuf<uint16_t> d_in_slot, d_out_slot;
  DevBuf<DevTaskD> d_tasks;
  DevBuf<DevCallD> d_calls;
  DevBuf<uint32_t> d_call_slots, d_deps, d_output_slots;
  // instance state
  DevBuf<uint4> d_labels, d_delta, d_ct, d_commit, d_io, d_stage;
  DevBuf<uint8_t> d_vals, d_io_bits;
  // result / input staging, allocated once: cudaMalloc / cudaFree in a call would synchronise the device,
  // i.e. wait for the persistent kernels of the other sessions sharing the GPU
  DevBuf<uint32_t> d_gather_slots;   // 0, 1, 2 .. 2 + n_inputs (constants, inputs)
  DevBuf<uint4> d_ev_true, d_ev_false, d_ev_in;
  DevBuf<uint8_t> d_ev_bits;
  DevBuf<uint32_t> d_flags, d_ctrl;  // d_ctrl[1] = error flag, [4..6] = scheduler head / tail / completed
  DevBuf<uint32_t> d_succ_off, d_succ, d_pending;
  // call pipelining (program.h): reverse start-dependency edges, window tables, per (group, slot) ready flags
  DevBuf<uint32_t> d_start_succ_off, d_start_succ, d_win_in_off, d_win_out_off, d_slot_flags;
  DevBuf<uint16_t> d_win_in, d_win_out;
  DevBuf<unsigned long long> d_queue, d_limit;
  DevBuf<uint32_t> d_park_head, d_park_next;
  uint64_t park_q = 1;
  uint32_t queue_log2 = 0;
  DevBuf<unsigned long long> d_progress;
  DevBuf<unsigned long long> d_seeds;
  DevBuf<unsigned long long> d_prof;  // GSV_PROFILE=1: per-phase worker cycle totals
  // lane mode (one warp = 32 instances, emission-order tasks)
  bool lane_mode = false;
  uint32_t B_pad = 0;            // B rounded up to whole groups
  uint32_t scratch_stride = 0;   // scratch slots per worker warp
  DevBuf<uint4> d_seq_gates, d_scratch;
  DevBuf<uint16_t> d_seq_in_slot, d_seq_out_slot;
  DevBuf<uint8_t> d_scratch_vals;
  bool ct_valid = false;
  std::vector<int> ct_sink_fds;    // gsv_session_set_ciphertext_files: gc_{i}.bin writers fed by the host drain
  std::shared_ptr<gsv_link> link;  // gsv_session_link: this session is one end of a garbler -> evaluator stream
  bool link_garbler = false;
  // GSV_CT_COMMIT_HOST: ring drained to pinned host buffers, chains folded by AES-NI threads
  static constexpr int HC_BUFS = 4;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t hc_ev[HC_BUFS] = {nullptr, nullptr, nullptr, nullptr};
  uint8_t* hc_buf[HC_BUFS] = {nullptr, nullptr, nullptr, nullptr};
  size_t hc_buf_bytes = 0;
  unsigned long long* hc_ready = nullptr;      // mapped: stream positions complete (device writes)
  unsigned long long* hc_ready_dev = nullptr;  // its device address
  unsigned long long* hc_consumed = nullptr;   // pinned [HC_BUFS]: sources of the back-pressure copies
  ~gsv_session() {
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (auto& e : hc_ev)
      if (e) cudaEventDestroy(e);
    for (auto& b : hc_buf)
      if (b) cudaFreeHost(b);
    if (hc_ready) cudaFreeHost(hc_ready);
    if (hc_consumed) cudaFreeHost(hc_consumed);
    if (stream) cudaStreamDestroy(stream);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};

namespace {

void upload_program(gsv_session* s) {
  const gsv::Program& g = s->prog->prog;
  std::vector<uint4> gates;
  std::vector<uint16_t> in_slot, out_slot;
  std::vector<uint4> seq_gates;
  std::vector<uint16_t> seq_in_slot, seq_out_slot;
  std::vector<DevTaskD> tasks;
  std::vector<uint32_t> win_in_off, win_out_off;
  std::vector<uint16_t> win_in, win_out;
  for (const gsv::Task& t : g.tasks) {
    DevTaskD d;
    d.n_windows = (uint32_t)t.win_in_off.size() - 1;
    d.window_levels = t.window_levels;
    d.win_off = (uint32_t)win_in_off.size();
    d.win_in_base = (uint32_t)win_in.size();
    d.win_out_base = (uint32_t)win_out.size();
    win_in_off.insert(win_in_off.end(), t.win_in_off.begin(), t.win_in_off.end());
    win_out_off.insert(win_out_off.end(), t.win_out_off.begin(), t.win_out_off.end());
    win_in.insert(win_in.end(), t.win_in.begin(), t.win_in.end());
    win_out.insert(win_out.end(), t.win_out.begin(), t.win_out.end());
    d.gate_off = (uint32_t)gates.size();
    d.n_gates = (uint32_t)t.gates.size();
    d.n_levels = t.n_levels;
    d.n_in = t.n_in;
    d.n_out = t.n_out;
    d.n_slots = t.n_slots;
    d.in_slot_off = (uint32_t)in_slot.size();
    d.out_slot_off = (uint32_t)out_slot.size();
    d.n_ct = (uint32_t)t.n_ct;
    d.seq_gate_off = (uint32_t)seq_gates.size();
    d.n_seq_gates = (uint32_t)t.seq_gates.size();
    d.n_seq_slots = t.n_seq_slots;
    for (const gsv::DevGate& dg : t.gates) {
      uint4 v;
      memcpy(&v, &dg, 16);
      gates.push_back(v);
    }
    for (const gsv::DevGate& dg : t.seq_gates) {
      uint4 v;
      memcpy(&v, &dg, 16);
      seq_gates.push_back(v);
    }
    seq_in_slot.insert(seq_in_slot.end(), t.seq_in_slot.begin(), t.seq_in_slot.end());
    seq_out_slot.insert(seq_out_slot.end(), t.seq_out_slot.begin(), t.seq_out_slot.end());
    in_slot.insert(in_slot.end(), t.in_slot.begin(), t.in_slot.end());
    out_slot.insert(out_slot.end(), t.out_slot.begin(), t.out_slot.end());
    tasks.push_back(d);
  }
  // level offsets are task-relative already (gate_off added on device)
  std::vector<DevCallD> calls;
  for (const gsv::Call& c : g.calls) {
    DevCallD d;
    d.task = c.task;
    d.in_off = c.in_off;
    d.out_off = c.out_off;
    d.dep_off = c.dep_off;
    d.n_deps = c.n_deps;
    d.pad = 0;
    d.gid_base = c.gid_base;
    d.ct_base = c.ct_base;
    calls.push_back(d);
  }
  if (gates.empty()) gates.push_back(make_uint4(0, 0, 0, 0));
  if (in_slot.empty()) in_slot.push_back(0);
  if (out_slot.empty()) out_slot.push_back(0);
  if (s->lane_mode) {
    // lane mode only needs the emission-order form
    if (seq_gates.empty()) seq_gates.push_back(make_uint4(0, 0, 0, 0));
    if (seq_in_slot.empty()) seq_in_slot.push_back(0);
    if (seq_out_slot.empty()) seq_out_slot.push_back(0);
    s->d_seq_gates.upload(seq_gates);
    s->d_seq_in_slot.upload(seq_in_slot);
    s->d_seq_out_slot.upload(seq_out_slot);
    gates.assign(1, make_uint4(0, 0, 0, 0));
  }
  s->d_gates.upload(gates);
  s->d_in_slot.upload(in_slot);
  s->d_out_slot.upload(out_slot);
  s->d_tasks.upload(tasks);
  if (win_in.empty()) win_in.push_back(0);
  if (win_out.empty()) win_out.push_back(0);
  s->d_win_in_off.upload(win_in_off);
  s->d_win_out_off.upload(win_out_off);
  s->d_win_in.upload(win_in);
  s->d_win_out.upload(win_out);
  s->d_calls.upload(calls);
  std::vector<uint32_t> cs = g.call_slots, dp = g.deps, os = g.output_slots;
  if (cs.empty()) cs.push_back(0);
  if (dp.empty()) dp.push_back(0);
  if (os.empty()) os.push_back(0);
  s->d_call_slots.upload(cs);
  s->d_deps.upload(dp);
  s->d_output_slots.upload(os);
}

// everything gsv_evaluate_batch needs on the device besides the ciphertexts
void ensure_eval_buffers(gsv_session* s) {
  const gsv::Program& g = s->prog->prog;
  const size_t B = s->B, n_in = g.n_inputs, n_out = g.output_slots.size();
  if (s->d_vals.n < (size_t)s->B_pad * g.n_global_slots) s->d_vals.alloc((size_t)s->B_pad * g.n_global_slots);
  if (s->d_ev_true.n < B) {
    s->d_ev_true.alloc(B);
    s->d_ev_false.alloc(B);
    s->d_ev_in.alloc(std::max<size_t>(B * n_in, 1));
    s->d_ev_bits.alloc(std::max<size_t>(B * n_in, 1));
  }
  if (s->d_io_bits.n < B * std::max<size_t>(n_out, 1)) s->d_io_bits.alloc(B * std::max<size_t>(n_out, 1));
}

// pinned drain buffers, copy stream and the mapped progress word of the host-folded chain
void ensure_host_chain_resources(gsv_session* s) {
  if (s->copy_stream) return;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  s->hc_buf_bytes = 64u << 20;
  if (const char* e = getenv("GSV_HOST_CHAIN_BUF_MB")) s->hc_buf_bytes = (size_t)std::max(1, atoi(e)) << 20;
  for (int b = 0; b < gsv_session::HC_BUFS; b++) {
    CUDA_TRY(cudaEventCreateWithFlags(&s->hc_ev[b], cudaEventDisableTiming));
    CUDA_TRY(cudaHostAlloc((void**)&s->hc_buf[b], s->hc_buf_bytes, cudaHostAllocDefault));
  }
  CUDA_TRY(cudaHostAlloc((void**)&s->hc_ready, 64, cudaHostAllocMapped));
  CUDA_TRY(cudaHostGetDevicePointer((void**)&s->hc_ready_dev, s->hc_ready, 0));
  CUDA_TRY(cudaHostAlloc((void**)&s->hc_consumed, 8 * gsv_session::HC_BUFS, cudaHostAllocDefault));
}

EngineParams make_params(gsv_session* s) {
  EngineParams p;
  memset(&p, 0, sizeof(p));
  p.gates = s->d_gates.p;
  p.in_slot = s->d_in_slot.p;
  p.out_slot = s->d_out_slot.p;
  p.tasks = s->d_tasks.p;
  p.calls = s->d_calls.p;
  p.call_slots = s->d_call_slots.p;
  p.deps = s->d_deps.p;
  p.labels = s->d_labels.p;
  p.vals = s->d_vals.p;
  p.delta = s->d_delta.p;
  p.ct = s->d_ct.p;
  p.flags = s->d_flags.p;
  p.sched = s->d_ctrl.p + 4;
  p.queue = s->d_queue.p;
  p.pending = s->d_pending.p;
  p.succ_off = s->d_succ_off.p;
  p.succ = s->d_succ.p;
  p.start_succ_off = s->d_start_succ_off.p;
  p.start_succ = s->d_start_succ.p;
  p.win_in_off = s->d_win_in_off.p;
  p.win_out_off = s->d_win_out_off.p;
  p.win_in = s->d_win_in.p;
  p.win_out = s->d_win_out.p;
  p.slot_flags = s->d_slot_flags.n ? s->d_slot_flags.p : nullptr;
  p.queue_log2 = s->queue_log2;
  p.sched_limit = s->d_limit.p;
  p.park_head = s->d_park_head.p;
  p.park_next = s->d_park_next.p;
  p.park_q = s->park_q;
  p.n_buckets = (uint32_t)s->d_park_head.n;
  p.n_progress = s->ct_mode == GSV_CT_COMMIT_HOST ? 1u : (s->B + CHAIN_INST - 1) / CHAIN_INST;
  p.error_flag = s->d_ctrl.p + 1;
  p.chain_progress = s->d_progress.p;
  p.commit = s->d_commit.p;
  p.ct_ring = s->ct_ring;
  p.n_workers = s->n_workers;
  p.n_chain_warps = 0;
  p.n_calls = (uint32_t)s->prog->prog.calls.size();
  p.n_groups = s->n_groups;
  p.n_global_slots = s->prog->prog.n_global_slots;
  p.B = s->B;
  p.slots_per_worker = s->slots_per_worker;
  p.worker_threads = s->NT;
  p.epoch = s->epoch;
  p.G = s->G;
  p.seq_gates = s->d_seq_gates.p;
  p.seq_in_slot = s->d_seq_in_slot.p;
  p.seq_out_slot = s->d_seq_out_slot.p;
  p.scratch = s->d_scratch.p;
  p.scratch_vals = s->d_scratch_vals.p;
  p.scratch_stride = s->scratch_stride;
  p.host_chain = s->ct_mode == GSV_CT_COMMIT_HOST ? 1u : 0u;
  p.ct_pos_stride = s->B;
  p.ct_quad_stride = 0;
  p.ct_qshift = 31;
  if (p.host_chain) {  // [instance quad][ring position][4]: each quad of chains drains as one sequential host stream
    p.ct_pos_stride = 4;
    p.ct_quad_stride = s->d_ct.n / ((s->B + 3) / 4);
    p.ct_qshift = 2;
  }
  p.host_ready = s->hc_ready_dev;
  p.prof = s->d_prof.p;
  p.flow_control = s->ct_ring ? 1u : 0u;  // garbling into a ring: admit items against the consumers' progress
  p.limit_add = p.free_until = s->ct_ring;
  p.flow_total = s->prog->prog.total_ct;
  return p;
}

template <int MODE>
void launch_engine(gsv_session* s, int hasher, const EngineParams& p) {
  // scheduler reset: empty queue, counters at zero, dependency counts loaded, root items queued
  CUDA_TRY(cudaMemsetAsync(s->d_ctrl.p + 4, 0, 16, s->stream));
  CUDA_TRY(cudaMemsetAsync(s->d_queue.p, 0, s->d_queue.n * 8, s->stream));
  if (s->d_park_head.n) CUDA_TRY(cudaMemsetAsync(s->d_park_head.p, 0xFF, s->d_park_head.n * 4, s->stream));
  if (p.slot_flags) {  // pipelined plan: constants and circuit inputs are valid from the start
    const uint32_t n = 2 + s->prog->prog.n_inputs;
    k_mark_slots<0><<<(unsigned)(((size_t)p.n_groups * n + 255) / 256), 256, 0, s->stream>>>(p.slot_flags, p.n_groups, p.n_global_slots, n, p.epoch);
    CUDA_TRY(cudaGetLastError());
  }
  {
    const size_t n_items = (size_t)p.n_calls * p.n_groups;
    k_sched_init<<<(unsigned)std::min<size_t>((n_items + 255) / 256 + 1, 4096), 256, 0, s->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
  }
  if (s->lane_mode) {
    if constexpr (MODE == 2) throw std::runtime_error("execute mode runs on the levelised kernel (exec_mode 1)");
    dim3 lgrid(s->sm_count), lblock(32 * s->n_workers);
    if constexpr (MODE != 2) {
      if (hasher == GSV_HASH_AES) {
        k_lane<HASH_AES, MODE><<<lgrid, lblock, AES_TABLE_BYTES, s->stream>>>(p);
      } else if (hasher == GSV_HASH_BLAKE3) {
        k_lane<HASH_BLAKE3, MODE><<<lgrid, lblock, AES_TABLE_BYTES, s->stream>>>(p);
      } else {
        throw std::runtime_error("unknown hasher");
      }
    }
    CUDA_TRY(cudaGetLastError());
    return;
  }
  const size_t smem = MODE == 0 ? s->smem_garble : s->smem_eval;
  dim3 grid(s->sm_count), block(std::max<uint32_t>(s->n_workers * s->NT, 32 * (p.n_chain_warps + (p.n_chain_ctas ? 1 : 0))));  // + governor warp
#define GSV_LAUNCH(GG, HH)                                                                              \
  do {                                                                                                  \
    k_engine<GG, HH, MODE><<<grid, block, smem, s->stream>>>(p);                                        \
  } while (0)
#define GSV_LAUNCH_G(HH)                         \
  switch (s->G) {                                \
    case 1: GSV_LAUNCH(1, HH); break;            \
    case 2: GSV_LAUNCH(2, HH); break;            \
    case 4: GSV_LAUNCH(4, HH); break;            \
    case 8: GSV_LAUNCH(8, HH); break;            \
    default: throw std::runtime_error("bad group size"); \
  }
  if constexpr (MODE == 2) {
    GSV_LAUNCH_G(HASH_AES)  // execute mode hashes nothing: one instantiation
  } else {
    if (hasher == GSV_HASH_AES) { GSV_LAUNCH_G(HASH_AES) }
    else if (hasher == GSV_HASH_BLAKE3) { GSV_LAUNCH_G(HASH_BLAKE3) }
    else throw std::runtime_error("unknown hasher");
  }
#undef GSV_LAUNCH_G
#undef GSV_LAUNCH
  CUDA_TRY(cudaGetLastError());
}

// Drains the ciphertext ring of a running GSV_CT_COMMIT_HOST garbling kernel and folds the B chains
// on host threads.  Called right after the kernel launch; returns when every chain is complete.
// Device layout [instance quad][ring position][4] (make_params): every drain is one 2-D copy of
// n_quads rows of n * 64 contiguous bytes, and a fold step of a quad is one 64-byte load.
// Waiting threads sleep (the drain loop and the fold threads of several sessions / ranks share the
// host cores; spinning on yield() would take the cores the folds need).
struct HostChainStats {
  double fold_busy = 0;         // busiest fold thread: share of the run spent folding
  double drain_wait_ready = 0;  // drain loop: share of the run spent waiting for the kernel to publish data
  double drain_wait_slot = 0;   // ... waiting for a host buffer the fold threads still hold
};
// What the drain loop works on.  Garbling with GSV_CT_COMMIT_HOST: the session's own ring, published by its
// own kernel, ring space returned through a device word (stream-ordered copy behind each drain).  A linked
// evaluator: the link's ring (written by the garbler's GPU), published by the garbler, ring space returned
// through the link's release word = min(drained, consumed by the evaluate kernel).
struct DrainSource {
  const uint4* ring = nullptr;       // device pointer, layout [instance quad][position][4]
  uint64_t cap = 0;                  // positions per instance in the buffer
  uint64_t ring_positions = 0;       // 0: the whole stream is resident (no wrap)
  const volatile unsigned long long* ready = nullptr;  // host-visible: stream positions complete
  cudaStream_t watch = nullptr;      // a kernel whose failure (or end before `total`) aborts the drain; may be null
  unsigned long long* dev_progress = nullptr;          // device word to advance behind every drain copy (or null)
  std::atomic<uint64_t>* copied_done = nullptr;        // out: positions whose copy has completed (or null)
  std::function<bool()> tick;        // called every loop iteration; return false to abort (or empty)
};
void run_host_chain(gsv_session* s, const DrainSource& src, uint8_t* commits, HostChainStats* stats) {
  const auto t_run0 = std::chrono::steady_clock::now();
  auto secs = [](std::chrono::steady_clock::duration d) { return std::chrono::duration<double>(d).count(); };
  std::vector<double> busy(64, 0.0);
  double wait_ready = 0, wait_slot = 0;
  const gsv::Program& g = s->prog->prog;
  const uint32_t B = s->B, nq = (B + 3) / 4;
  const uint64_t total = g.total_ct;
  const size_t row_bytes = (size_t)nq * 64;
  const uint64_t chunk_pos = std::max<uint64_t>(1, s->hc_buf_bytes / row_bytes);
  const uint64_t cap = src.cap, ring = src.ring_positions;
  constexpr int NB = gsv_session::HC_BUFS;
  const uint32_t T = std::max<uint32_t>(1, std::min(s->host_threads, nq));
  std::vector<uint8_t> h((size_t)nq * 64, 0);
  std::atomic<uint64_t> slot_job[NB];   // 1 + index of the job whose copy was enqueued into the slot
  std::atomic<uint64_t> slot_done[NB];  // hasher completions on the slot, over all jobs
  uint64_t slot_npos[NB] = {0, 0, 0, 0}, slot_end[NB] = {0, 0, 0, 0};
  for (int b = 0; b < NB; b++) {
    slot_job[b].store(0);
    slot_done[b].store(0);
  }
  std::atomic<int64_t> n_jobs{-1};
  std::atomic<bool> abort{false};
  const auto nap = [] { std::this_thread::sleep_for(std::chrono::microseconds(50)); };
  std::atomic<bool> sink_failed{false};
  std::vector<uint8_t> sink_buf;  // sized before the threads start: one run of a chunk's records per fold thread
  if (!s->ct_sink_fds.empty()) sink_buf.resize((size_t)T * chunk_pos * 16);
  auto hasher = [&](uint32_t t) {
    const uint32_t q0 = (uint32_t)((uint64_t)nq * t / T), q1 = (uint32_t)((uint64_t)nq * (t + 1) / T);
    cudaSetDevice(s->device);
    for (uint64_t j = 0;; j++) {
      const int b = (int)(j % NB);
      while (slot_job[b].load(std::memory_order_acquire) != j + 1) {
        const int64_t nj = n_jobs.load(std::memory_order_acquire);
        if (abort.load() || (nj >= 0 && (int64_t)j >= nj)) return;
        nap();
      }
      if (cudaEventSynchronize(s->hc_ev[b]) != cudaSuccess) {
        abort.store(true);
        return;
      }
      if (t == 0 && src.copied_done) src.copied_done->store(slot_end[b], std::memory_order_release);  // jobs are in stream order
      if (q1 > q0) {
        const auto t0 = std::chrono::steady_clock::now();
        gsv::host_chain_fold_quads(h.data() + (size_t)q0 * 64, s->hc_buf[b] + (size_t)q0 * chunk_pos * 64, chunk_pos * 64,
                                   slot_npos[b], q1 - q0);
        busy[t % busy.size()] += secs(std::chrono::steady_clock::now() - t0);
      }
      if (!s->ct_sink_fds.empty() && q1 > q0) {
        // gc_{i}.bin (ciphertext_repository.rs:94-106): the drained rows are [quad][position][4 chains]; each chain's
        // 16-byte records are gathered into a contiguous run and written at their stream offset
        const uint64_t n = slot_npos[b], first = slot_end[b] - n;
        uint8_t* out = sink_buf.data() + (size_t)t * chunk_pos * 16;
        for (uint32_t q = q0; q < q1 && !abort.load(); q++)
          for (uint32_t j = 0; j < 4 && 4 * q + j < B; j++) {
            const int fd = s->ct_sink_fds[4 * q + j];
            if (fd < 0) continue;
            const uint8_t* in = s->hc_buf[b] + (size_t)q * chunk_pos * 64 + 16 * j;
            for (uint64_t k = 0; k < n; k++) memcpy(out + 16 * k, in + 64 * k, 16);
            size_t off = 0;
            while (off < n * 16) {
              const ssize_t wr = pwrite(fd, out + off, n * 16 - off, (off_t)(first * 16 + off));
              if (wr <= 0) {
                sink_failed.store(true);
                abort.store(true);
                break;
              }
              off += (size_t)wr;
            }
          }
      }
      slot_done[b].fetch_add(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> threads;
  for (uint32_t t = 0; t < T; t++) threads.emplace_back(hasher, t);
  std::string err;
  try {
    uint64_t copied = 0, jobs = 0;
    auto last_query = std::chrono::steady_clock::now();
    while (copied < total) {
      if (src.tick && !src.tick()) throw std::runtime_error("the linked session failed");
      const uint64_t ready = *src.ready;
      uint64_t n = ready > copied ? std::min<uint64_t>(ready - copied, chunk_pos) : 0;
      uint64_t pos = copied;
      if (ring) {
        pos = copied % ring;
        n = std::min<uint64_t>(n, ring - pos);
      }
      // small drains waste DMA launches: unless the stream or the ring ends, wait for half a buffer (or a
      // quarter of a small ring: the producer stalls once the ring is full, so never wait for more than it holds)
      uint64_t thresh = chunk_pos / 2;
      if (ring) thresh = std::min<uint64_t>(thresh, std::max<uint64_t>(1, ring / 4));
      const bool worth = n > 0 && (n >= thresh || copied + n == total || (ring && pos + n == ring));
      if (!worth) {
        auto now = std::chrono::steady_clock::now();
        if (src.watch && now - last_query > std::chrono::milliseconds(5)) {
          last_query = now;
          cudaError_t q = cudaStreamQuery(src.watch);
          if (q != cudaSuccess && q != cudaErrorNotReady) throw std::runtime_error(std::string("kernel failed: ") + cudaGetErrorString(q));
          if (q == cudaSuccess && *src.ready < total && !src.tick)
            throw std::runtime_error("garbling kernel ended before the stream was complete");
        }
        nap();
        wait_ready += secs(std::chrono::steady_clock::now() - now);
        continue;
      }
      const int b = (int)(jobs % NB);
      {
        const auto t0 = std::chrono::steady_clock::now();
        while (slot_done[b].load(std::memory_order_acquire) != (uint64_t)T * (jobs / NB)) {
          if (abort.load()) throw std::runtime_error("host chain thread failed");
          if (src.tick && !src.tick()) throw std::runtime_error("the linked session failed");
          nap();
        }
        wait_slot += secs(std::chrono::steady_clock::now() - t0);
      }
      CUDA_TRY(cudaMemcpy2DAsync(s->hc_buf[b], (size_t)chunk_pos * 64, src.ring + pos * 4, (size_t)cap * 64, (size_t)n * 64, nq,
                                 cudaMemcpyDeviceToHost, s->copy_stream));
      if (src.dev_progress) {
        s->hc_consumed[b] = copied + n;
        CUDA_TRY(cudaMemcpyAsync(src.dev_progress, s->hc_consumed + b, 8, cudaMemcpyHostToDevice, s->copy_stream));
      }
      CUDA_TRY(cudaEventRecord(s->hc_ev[b], s->copy_stream));
      slot_npos[b] = n;
      slot_end[b] = copied + n;
      slot_job[b].store(jobs + 1, std::memory_order_release);
      jobs++;
      copied += n;
    }
    n_jobs.store((int64_t)jobs, std::memory_order_release);
    // the last copies: keep the release word moving until the fold threads have seen them
    if (src.tick)
      while (src.copied_done && src.copied_done->load(std::memory_order_acquire) < total && !abort.load()) {
        if (!src.tick()) throw std::runtime_error("the linked session failed");
        nap();
      }
  } catch (const std::exception& e) {
    err = e.what();
    abort.store(true);
  }
  for (auto& t : threads) t.join();
  if (err.empty() && sink_failed.load()) err = "writing a ciphertext file failed";
  if (err.empty() && abort.load()) err = "host chain thread failed";
  if (!err.empty()) {
    // unblock the persistent kernel before reporting: with a ring its workers park behind the progress
    // word nobody advances any more; "everything consumed" lets it run to completion
    if (src.dev_progress) {
      s->hc_consumed[0] = ~0ull >> 1;
      cudaMemcpyAsync(src.dev_progress, s->hc_consumed, 8, cudaMemcpyHostToDevice, s->copy_stream);
      cudaStreamSynchronize(s->copy_stream);
      cudaStreamSynchronize(s->stream);
    }
    throw std::runtime_error(err);
  }
  for (uint32_t i = 0; i < B; i++) memcpy(commits + (size_t)i * 16, h.data() + (size_t)i * 16, 16);
  if (stats) {
    const double total_s = std::max(1e-9, secs(std::chrono::steady_clock::now() - t_run0));
    stats->fold_busy = *std::max_element(busy.begin(), busy.end()) / total_s;
    stats->drain_wait_ready = wait_ready / total_s;
    stats->drain_wait_slot = wait_slot / total_s;
  }
}

}  // namespace

extern "C" {

gsv_session* gsv_session_create(const gsv_program* p, const gsv_session_options* opt) {
  if (!p || !opt) {
    fail(GSV_ERR_INVALID, "null argument");
    return nullptr;
  }
  try {
    ensure_device(opt->device);
    auto s = std::make_unique<gsv_session>();
    s->prog = p;
    s->device = opt->device;
    s->B = opt->n_instances;
    s->ct_mode = opt->ct_mode;
    if (s->B == 0) throw std::runtime_error("n_instances must be > 0");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
    s->sm_count = prop.multiProcessorCount;
    // spatial sharing: several sessions (e.g. software-pipelined cut-and-choose batches) each run their
    // persistent grid on a slice of the SMs
    if (opt->sm_limit) s->sm_count = std::max(2, std::min<int>(s->sm_count, (int)opt->sm_limit));
    const gsv::Program& g = p->prog;
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    uint32_t slots = std::max<uint32_t>(g.max_task_slots, 4);
    slots = (slots + 3) & ~3u;
    s->slots_per_worker = slots;
    s->NT = opt->worker_threads ? opt->worker_threads : 256;
    if (s->NT != 64 && s->NT != 128 && s->NT != 256 && s->NT != 512)
      throw std::runtime_error("worker_threads must be 64/128/256/512");
    // commitment consumers: one chain warp per CHAIN_INST instances, packed a few warps per SMSP
    // onto dedicated trailing CTAs (SMs) of the persistent grid
    const uint32_t chain_warps_total = (s->ct_mode == GSV_CT_COMMIT || s->ct_mode == GSV_CT_KEEP) ? (s->B + CHAIN_INST - 1) / CHAIN_INST : 0;
    uint32_t n_chain = 0;  // chain warps per chain CTA
    s->n_chain_ctas = 0;
    if (chain_warps_total) {
      // both persistent kernels run 512-thread CTAs: at most 15 chain warps + the governor warp
      uint32_t per_cta = 15;
      if (const char* e = getenv("GSV_CHAIN_WARPS_PER_SM")) per_cta = std::max(1, std::min(15, atoi(e)));  // + 1 governor warp
      n_chain = std::min(per_cta, chain_warps_total);
      s->n_chain_ctas = (chain_warps_total + n_chain - 1) / n_chain;
      if (s->n_chain_ctas * 2 > (uint32_t)s->sm_count)
        throw std::runtime_error("too many instances for one GPU (commitment warps would take over half the SMs)");
    }
    if (s->ct_mode == GSV_CT_COMMIT_HOST) {
      if (!gsv::host_chain_available()) throw std::runtime_error("GSV_CT_COMMIT_HOST needs a host CPU with AES-NI");
      n_chain = 1;  // one publisher warp on one trailing CTA
      s->n_chain_ctas = 1;
    }
    {
      // host fold threads (GSV_CT_COMMIT_HOST, linked evaluators): one per quad of chains up to half the hardware
      // threads; callers that run several sessions or ranks per host pass their share explicitly
      unsigned hw = std::thread::hardware_concurrency();
      if (hw == 0) hw = 4;
      s->host_threads = opt->host_threads ? opt->host_threads : std::max(1u, hw / 2);
      if (const char* e = getenv("GSV_HOST_CHAIN_THREADS")) s->host_threads = std::max(1, atoi(e));
    }
    uint32_t n_workers = std::min<uint32_t>(ENGINE_MAX_THREADS / s->NT, 15);  // worker w syncs on named barrier w + 1 (ids 1..15)
    if (n_workers == 0) throw std::runtime_error("worker_threads too large");
    // largest G (power of two dividing B, <= 8) whose label working set fits next to the tables
    auto smem_for = [&](uint32_t G, uint32_t nw, bool eval) {
      size_t lab = (size_t)slots * G;
      return (size_t)AES_TABLE_BYTES + nw * lab * 16 + (((eval ? nw * lab : 0) + nw * 8 + 15) & ~(size_t)15) + (size_t)nw * GATE_RING * 16 +
             (size_t)nw * PROF_WORDS * 8 + (size_t)nw * GATE_CHUNKS * 8;  // + one mbarrier per ring slot
    };
    // execution mode: lane mode (warp = 32 instances) for batches that fill warps, the levelised
    // shared-memory mode otherwise
    if (opt->exec_mode > 2) throw std::runtime_error("exec_mode must be 0 (auto), 1 (levelised) or 2 (lane)");
    s->lane_mode = opt->exec_mode == 2 || (opt->exec_mode == 0 && opt->group == 0 && s->B >= 128);
    if (!g.has_levelised) {
      if (opt->exec_mode == 1) throw std::runtime_error("program was planned lane-only; it has no levelised form");
      s->lane_mode = true;
    }
    uint32_t G = opt->group;
    if (s->lane_mode) {
      G = 32;
      s->NT = 32;
      n_workers = LANE_WARPS;
      s->B_pad = (s->B + 31) / 32 * 32;
      s->n_groups = s->B_pad / 32;
      s->scratch_stride = std::max<uint32_t>(g.max_task_seq_slots, 4);
    } else {
      if (G == 0) {
        G = 2;
        while (G > 1 && (s->B % G != 0)) G >>= 1;
      }
      if (G != 1 && G != 2 && G != 4 && G != 8) throw std::runtime_error("group must be 1/2/4/8");
      if (s->B % G) throw std::runtime_error("n_instances must be a multiple of group");
      while (n_workers > 1 && smem_for(G, n_workers, true) > smem_max) n_workers--;
      if (smem_for(G, n_workers, true) > smem_max)
        throw std::runtime_error("task working set does not fit shared memory; lower group or max_task_slots");
      s->B_pad = s->B;
      s->n_groups = s->B / G;
      s->smem_garble = smem_for(G, n_workers, false);
      s->smem_eval = smem_for(G, n_workers, true);
      {
        // The persistent grid's CTAs wait for each other (queue, ring limit, chain / governor CTAs): every CTA must be
        // resident.  Check statically that the shape fits an SM (registers x threads, shared memory); that the SMs
        // are not held by somebody else's kernels is the caller's side of the contract (sm_limit for sessions that
        // share a GPU).
        const int threads = (int)std::max<uint32_t>(n_workers * s->NT, 32 * (n_chain + (s->n_chain_ctas ? 1 : 0)));
        int resident = 0;
        cudaError_t e = cudaSuccess;
        switch (G) {
          case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_engine<1, HASH_BLAKE3, 1>, threads, s->smem_eval); break;
          case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_engine<2, HASH_BLAKE3, 1>, threads, s->smem_eval); break;
          case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_engine<4, HASH_BLAKE3, 1>, threads, s->smem_eval); break;
          default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_engine<8, HASH_BLAKE3, 1>, threads, s->smem_eval); break;
        }
        CUDA_TRY(e);
        if (resident < 1) throw std::runtime_error("the engine's CTA shape does not fit one SM (threads x registers / shared memory)");
      }
    }
    s->G = G;
    s->n_workers = n_workers;
    s->n_chain_warps = n_chain;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e : s->ev) CUDA_TRY(cudaEventCreate(&e));
    if (s->ct_mode == GSV_CT_COMMIT_HOST) ensure_host_chain_resources(s.get());
    upload_program(s.get());
    s->d_labels.alloc((size_t)s->B_pad * g.n_global_slots);
    s->d_delta.alloc(s->B_pad);
    CUDA_TRY(cudaMemset(s->d_delta.p, 0, (size_t)s->B_pad * 16));
    if (s->lane_mode) {
      const size_t n = (size_t)s->sm_count * n_workers * s->scratch_stride * 32;
      s->d_scratch.alloc(n);
      s->d_scratch_vals.alloc(n);
    }
    s->d_commit.alloc(s->B);
    s->d_seeds.alloc(s->B);
    {
      std::vector<uint32_t> gs(2 + g.n_inputs);
      for (size_t i = 0; i < gs.size(); i++) gs[i] = (uint32_t)i;
      s->d_gather_slots.upload(gs);
      s->d_io.alloc((size_t)s->B * std::max<size_t>({(size_t)g.n_inputs, g.output_slots.size(), (size_t)1}));
    }
    s->d_flags.alloc((size_t)g.calls.size() * s->n_groups + 1);
    CUDA_TRY(cudaMemset(s->d_flags.p, 0, s->d_flags.n * 4));
    s->d_ctrl.alloc(8);
    s->d_progress.alloc((size_t)std::max<uint32_t>(s->n_chain_ctas * n_chain, 1));
    CUDA_TRY(cudaMemset(s->d_ctrl.p, 0, 32));
    {
      // scheduler state: reverse dependency edges, per-item counters, the ready queue
      const size_t n_items = g.calls.size() * (size_t)s->n_groups;
      if (n_items >= (1ull << 30)) throw std::runtime_error("too many work items (calls x instance groups)");
      // Pipelined plans on the levelised kernel: a call's START dependencies (the producers of its inputs) are
      // released when the producer starts, the rest at completion.  Lane mode runs whole calls: every edge is
      // released at completion.
      const bool piped = g.pipelined && !s->lane_mode;
      auto reverse_edges = [&](bool start_edges, DevBuf<uint32_t>& d_off, DevBuf<uint32_t>& d_edges) {
        auto lo = [&](const gsv::Call& c) { return !piped ? 0u : start_edges ? 0u : c.n_start_deps; };
        auto hi = [&](const gsv::Call& c) { return !piped ? (start_edges ? 0u : c.n_deps) : start_edges ? c.n_start_deps : c.n_deps; };
        std::vector<uint32_t> off(g.calls.size() + 1, 0);
        for (const gsv::Call& c : g.calls)
          for (uint32_t k = lo(c); k < hi(c); k++) off[g.deps[c.dep_off + k] + 1]++;
        for (size_t i = 0; i < g.calls.size(); i++) off[i + 1] += off[i];
        std::vector<uint32_t> edges(std::max<size_t>(off.back(), 1), 0), fill(off.begin(), off.end() - 1);
        for (size_t c = 0; c < g.calls.size(); c++)
          for (uint32_t k = lo(g.calls[c]); k < hi(g.calls[c]); k++) edges[fill[g.deps[g.calls[c].dep_off + k]]++] = (uint32_t)c;
        d_off.upload(off);
        d_edges.upload(edges);
      };
      reverse_edges(false, s->d_succ_off, s->d_succ);
      reverse_edges(true, s->d_start_succ_off, s->d_start_succ);
      if (piped) {
        s->d_slot_flags.alloc((size_t)s->n_groups * g.n_global_slots);
        CUDA_TRY(cudaMemset(s->d_slot_flags.p, 0, s->d_slot_flags.n * 4));
      }
      s->d_pending.alloc(std::max<size_t>(n_items, 1));
      s->queue_log2 = 4;
      while ((1ull << s->queue_log2) < 2 * n_items) s->queue_log2++;
      s->d_queue.alloc((size_t)1 << s->queue_log2);
    }
    if (s->ct_mode != GSV_CT_NONE) {
      // GSV_CT_KEEP: the whole interleaved stream stays resident.  GSV_CT_COMMIT(_HOST): a
      // ring the chain warps drain (back-pressure through chain_progress).
      uint64_t total = std::max<uint64_t>(g.total_ct, 1);
      uint64_t max_task_ct = 1;
      for (const auto& t : g.tasks) max_task_ct = std::max<uint64_t>(max_task_ct, t.n_ct);
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
      // The ring bounds how far garbling may run ahead of the (serial, emission-ordered) chain, and
      // with it the task parallelism available inside one instance: make it as large as HBM allows.
      uint64_t budget = (uint64_t)(free_b * 0.85);
      if (opt->ct_buffer_bytes) budget = std::min<uint64_t>(budget, opt->ct_buffer_bytes);
      // the host-fold layout stores whole quads of instances
      const uint64_t B_ct = s->ct_mode == GSV_CT_COMMIT_HOST ? (uint64_t)((s->B + 3) / 4) * 4 : s->B;
      uint64_t ring = std::max<uint64_t>(budget / (B_ct * 16), 1);
      if (opt->ct_ring_log2) ring = std::min<uint64_t>(ring, 1ull << opt->ct_ring_log2);
      if (s->ct_mode == GSV_CT_KEEP || s->ct_mode == GSV_CT_KEEP_RAW || ring >= total) {
        if (total * B_ct * 16 > (uint64_t)(free_b * 0.9)) throw std::runtime_error("ciphertext stream does not fit in HBM; use GSV_CT_COMMIT");
        s->ct_ring = 0;
        s->d_ct.alloc((size_t)total * B_ct);
      } else {
        if (ring < 2 * max_task_ct) throw std::runtime_error("ciphertext ring smaller than two tasks; fewer instances needed");
        s->ct_ring = ring;
        s->d_ct.alloc((size_t)ring * B_ct);
        // ring governor: parking buckets of ring/64 ciphertexts
        s->park_q = std::max<uint64_t>(ring / 64, 1);
        s->d_park_head.alloc((size_t)(total / s->park_q + 2));
        s->d_park_next.alloc(std::max<size_t>(g.calls.size() * (size_t)s->n_groups, 1));
      }
    }
    s->d_limit.alloc(1);
    if (getenv("GSV_PROFILE")) {
      s->d_prof.alloc(PROF_WORDS);
      CUDA_TRY(cudaMemset(s->d_prof.p, 0, PROF_WORDS * 8));
    }
    return s.release();
  } catch (const std::exception& e) {
    fail(std::string(e.what()).find("no CUDA device") != std::string::npos ? GSV_ERR_NO_DEVICE : GSV_ERR_CUDA, e.what());
    return nullptr;
  }
}

int gsv_session_link(gsv_session* garbler, gsv_session* evaluator, uint64_t ring_bytes) {
  if (!garbler || !evaluator || garbler == evaluator) return fail(GSV_ERR_INVALID, "need two distinct sessions");
  if (garbler->prog != evaluator->prog || garbler->B != evaluator->B)
    return fail(GSV_ERR_INVALID, "linked sessions must share the program and the batch size");
  if (garbler->link || evaluator->link) return fail(GSV_ERR_INVALID, "session already linked");
  if (!gsv::host_chain_available()) return fail(GSV_ERR_INVALID, "the evaluator hashes the stream with host AES-NI threads");
  try {
    const gsv::Program& g = garbler->prog->prog;
    const uint32_t B = garbler->B;
    const uint64_t nq = (B + 3) / 4, total = std::max<uint64_t>(g.total_ct, 1);
    auto link = std::make_shared<gsv_link>();
    link->dev_g = garbler->device;
    link->dev_e = evaluator->device;
    if (link->dev_g != link->dev_e) {
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, link->dev_g, link->dev_e));
      if (!can) throw std::runtime_error("the garbler's GPU cannot store to the evaluator's GPU (no peer access)");
      CUDA_TRY(cudaSetDevice(link->dev_g));
      cudaError_t e = cudaDeviceEnablePeerAccess(link->dev_e, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
      cudaGetLastError();
    }
    CUDA_TRY(cudaSetDevice(link->dev_e));
    uint64_t max_task_ct = 1;
    for (const auto& t : g.tasks) max_task_ct = std::max<uint64_t>(max_task_ct, t.n_ct);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)(free_b * 0.85);
    if (ring_bytes) budget = std::min<uint64_t>(budget, ring_bytes);
    uint64_t ring = std::max<uint64_t>(budget / (nq * 64), 1);
    if (ring >= total) {
      link->ring_positions = 0;
      link->cap = total;
    } else {
      if (ring < 2 * max_task_ct) throw std::runtime_error("ciphertext ring smaller than two tasks");
      link->ring_positions = link->cap = ring;
    }
    CUDA_TRY(cudaMalloc(&link->ring, (size_t)link->cap * nq * 64));
    CUDA_TRY(cudaHostAlloc((void**)&link->words, 4 * 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(link->words, 0, 4 * 64);
    unsigned long long* dv = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer((void**)&dv, link->words, 0));
    link->ready_e = dv;
    link->done_e = dv + 8;
    CUDA_TRY(cudaSetDevice(link->dev_g));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&dv, link->words, 0));
    link->ready_g = dv;
    link->release_g = dv + 16;
    // both ends schedule against a progress word: parking lists (the evaluator always, the garbler with a ring)
    const uint64_t park_ring = link->ring_positions ? link->ring_positions : total;
    for (gsv_session* s : {garbler, evaluator}) {
      CUDA_TRY(cudaSetDevice(s->device));
      s->park_q = std::max<uint64_t>(park_ring / 64, 1);
      s->d_park_head.alloc((size_t)(total / s->park_q + 2));
      s->d_park_next.alloc(std::max<size_t>(g.calls.size() * (size_t)s->n_groups, 1));
    }
    CUDA_TRY(cudaSetDevice(link->dev_e));
    ensure_host_chain_resources(evaluator);
    ensure_eval_buffers(evaluator);
    garbler->link = evaluator->link = link;
    garbler->link_garbler = true;
    evaluator->link_garbler = false;
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_CUDA, e.what());
  }
}

void gsv_session_destroy(gsv_session* s) {
  if (s) cudaSetDevice(s->device);
  delete s;
}

int gsv_garble_batch(gsv_session* s, int hasher, const uint64_t* seeds, gsv_garble_result* res) {
  if (!s || !seeds || !res) return fail(GSV_ERR_INVALID, "null argument");
  try {
    CUDA_TRY(cudaSetDevice(s->device));
    const gsv::Program& g = s->prog->prog;
    const uint32_t B = s->B;
    uint32_t launches = 0;
    s->epoch++;
    res->host_fold_busy = res->host_drain_wait_kernel = res->host_drain_wait_fold = 0.f;
    const auto t_begin = std::chrono::steady_clock::now();
    CUDA_TRY(cudaMemcpyAsync(s->d_seeds.p, seeds, (size_t)B * 8, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->d_ctrl.p, 0, 16, s->stream));
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    {
      const uint32_t n_blocks = (3 + g.n_inputs + 3) / 4;
      const size_t total = (size_t)B * n_blocks;
      k_seed_expand<0><<<(unsigned)((total + 127) / 128), 128, 0, s->stream>>>(
          s->d_seeds.p, B, s->G, g.n_inputs, g.n_global_slots, s->d_labels.p, s->d_delta.p);
      CUDA_TRY(cudaGetLastError());
      launches++;
    }
    CUDA_TRY(cudaEventRecord(s->ev[1], s->stream));
    EngineParams p = make_params(s);
    p.write_ct = (s->ct_mode != GSV_CT_NONE) ? 1u : 0u;
    p.n_chain_warps = s->n_chain_warps;
    p.n_chain_ctas = s->n_chain_ctas;
    if (getenv("GSV_DEBUG_NO_CHAIN")) p.n_chain_ctas = 0;  // profiling aid: ciphertexts written, not folded
    if (s->n_chain_warps) CUDA_TRY(cudaMemsetAsync(s->d_progress.p, 0, s->d_progress.n * 8, s->stream));
    if (s->ct_mode == GSV_CT_COMMIT_HOST) {
      *reinterpret_cast<volatile unsigned long long*>(s->hc_ready) = 0;
      CUDA_TRY(cudaMemsetAsync(s->d_progress.p, 0, 8, s->stream));
    }
    const bool linked = s->link && s->link_garbler;
    if (linked) {
      // ciphertexts go to the ring in the evaluator's memory; the publisher warp announces the finished frontier,
      // the evaluator's host thread returns ring space through the release word
      gsv_link& L = *s->link;
      volatile unsigned long long* w = L.words;
      w[0] = w[8] = w[16] = 0;
      std::atomic_thread_fence(std::memory_order_seq_cst);
      L.failed.store(false);
      L.runs_started.fetch_add(1);
      if (getenv("GSV_LINK_DEBUG")) fprintf(stderr, "[link] garbler: launching\n");
      p.ct = L.ring;
      p.ct_ring = L.ring_positions;
      p.ct_pos_stride = 4;
      p.ct_quad_stride = L.cap * 4;
      p.ct_qshift = 2;
      p.write_ct = 1;
      p.host_chain = 1;
      p.host_ready = L.ready_g;
      p.flow_control = L.ring_positions ? 1u : 0u;
      p.limit_add = p.free_until = L.ring_positions;
      p.n_progress = 0;
      p.ext_progress = L.release_g;
      p.ct_sys = 1;
      p.n_chain_warps = 1;
      p.n_chain_ctas = 1;
    }
    launch_engine<0>(s, hasher, p);
    launches += 2;  // k_sched_init + the persistent engine kernel
    if (s->d_slot_flags.n) launches++;  // k_mark_slots (pipelined plan)
    CUDA_TRY(cudaEventRecord(s->ev[2], s->stream));
    std::vector<uint8_t> host_commits;
    double host_ms = 0.0;
    if (s->ct_mode == GSV_CT_COMMIT_HOST) {
      host_commits.resize((size_t)B * 16);
      HostChainStats hs;
      DrainSource src;
      src.ring = s->d_ct.p;
      src.cap = s->d_ct.n / ((size_t)((B + 3) / 4) * 4);
      src.ring_positions = s->ct_ring;
      src.ready = s->hc_ready;
      src.watch = s->stream;
      src.dev_progress = s->d_progress.p;
      run_host_chain(s, src, host_commits.data(), &hs);  // returns when the last chain is folded
      res->host_fold_busy = (float)hs.fold_busy;
      res->host_drain_wait_kernel = (float)hs.drain_wait_ready;
      res->host_drain_wait_fold = (float)hs.drain_wait_slot;
      host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
    // the chain commitment is folded inside k_engine by the chain warps (no separate launch)
    CUDA_TRY(cudaEventRecord(s->ev[3], s->stream));
    // ---- results
    if (res->delta) CUDA_TRY(cudaMemcpyAsync(res->delta, s->d_delta.p, (size_t)B * 16, cudaMemcpyDeviceToHost, s->stream));
    if (res->ct_commit && (s->ct_mode == GSV_CT_COMMIT || s->ct_mode == GSV_CT_KEEP))
      CUDA_TRY(cudaMemcpyAsync(res->ct_commit, s->d_commit.p, (size_t)B * 16, cudaMemcpyDeviceToHost, s->stream));
    if (res->ct_commit && s->ct_mode == GSV_CT_COMMIT_HOST) memcpy(res->ct_commit, host_commits.data(), host_commits.size());
    if (linked) {
      // wait for the kernel; once the evaluator has given up, report instead of waiting for ever
      std::chrono::steady_clock::time_point failed_at{};
      for (;;) {
        const cudaError_t q = cudaStreamQuery(s->stream);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) {
          s->link->failed.store(true);
          throw std::runtime_error(std::string("linked garbling kernel failed: ") + cudaGetErrorString(q));
        }
        if (s->link->failed.load()) {
          const auto now = std::chrono::steady_clock::now();
          if (failed_at == std::chrono::steady_clock::time_point{}) failed_at = now;
          if (now - failed_at > std::chrono::seconds(10)) {
            uint32_t sc[4] = {0, 0, 0, 0};
            cudaStream_t aux;
            cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking);
            cudaMemcpyAsync(sc, s->d_ctrl.p + 4, 16, cudaMemcpyDeviceToHost, aux);
            cudaStreamSynchronize(aux);
            cudaStreamDestroy(aux);
            volatile unsigned long long* w = s->link->words;
            throw std::runtime_error("linked garbling kernel still running after the evaluator failed: published " + std::to_string(w[0]) +
                                     ", release word " + std::to_string(w[16]) + ", queue head " + std::to_string(sc[0]) + " tail " +
                                     std::to_string(sc[1]) + ", items completed " + std::to_string(sc[2]) + " of " +
                                     std::to_string(g.calls.size() * (size_t)s->n_groups) + ", buckets released " + std::to_string(sc[3]));
          }
        }
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    }
    // labels of n consecutive entries of a device slot list -> host (staged through d_io)
    auto gather = [&](const uint32_t* d_slots, size_t n, uint8_t* host_out) {
      if (!host_out || n == 0) return;
      const size_t total = (size_t)B * n;
      k_gather_slots<0><<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
          s->d_labels.p, nullptr, d_slots, (uint32_t)n, B, s->G, g.n_global_slots, s->d_io.p, nullptr);
      CUDA_TRY(cudaGetLastError());
      launches++;
      CUDA_TRY(cudaMemcpyAsync(host_out, s->d_io.p, total * 16, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
    };
    gather(s->d_gather_slots.p, 1, res->false_label0);
    gather(s->d_gather_slots.p + 1, 1, res->true_label0);
    gather(s->d_gather_slots.p + 2, g.n_inputs, res->input_label0);
    gather(s->d_output_slots.p, g.output_slots.size(), res->output_label0);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaEventElapsedTime(&res->ms_seed, s->ev[0], s->ev[1]);
    cudaEventElapsedTime(&res->ms_garble, s->ev[1], s->ev[2]);
    cudaEventElapsedTime(&res->ms_commit, s->ev[2], s->ev[3]);
    cudaEventElapsedTime(&res->ms_total, s->ev[0], s->ev[3]);
    if (s->ct_mode == GSV_CT_COMMIT_HOST && host_ms > res->ms_total) {
      // the step ends when the last chain is folded: the drain / fold tail behind the kernel counts
      res->ms_total = (float)host_ms;
    }
    if (s->ct_mode == GSV_CT_COMMIT_HOST) res->ms_commit = res->ms_total - res->ms_seed - res->ms_garble;
    if (s->d_prof.p) {
      unsigned long long pc[PROF_WORDS];
      CUDA_TRY(cudaMemcpy(pc, s->d_prof.p, sizeof(pc), cudaMemcpyDeviceToHost));
      CUDA_TRY(cudaMemset(s->d_prof.p, 0, sizeof(pc)));
      const double tot = (double)std::max<unsigned long long>(pc[PROF_TOTAL], 1);
      fprintf(stderr,
              "[prof] worker cycles: wait %.1f%% gather %.1f%% aes-levels %.1f%% free-levels %.1f%% scatter %.1f%% complete %.1f%% | "
              "items %llu aes-levels %llu (%.0f cyc) free-levels %llu (%.0f cyc) passes/level %.2f gather %.0f scatter %.0f complete %.0f cyc/item\n",
              100 * pc[PROF_WAIT] / tot, 100 * pc[PROF_GATHER] / tot, 100 * pc[PROF_AES_LEVELS] / tot,
              100 * pc[PROF_FREE_LEVELS] / tot, 100 * pc[PROF_SCATTER] / tot, 100 * pc[PROF_COMPLETE] / tot,
              pc[PROF_N_ITEMS], pc[PROF_N_AES_LEVELS], (double)pc[PROF_AES_LEVELS] / std::max<unsigned long long>(pc[PROF_N_AES_LEVELS], 1),
              pc[PROF_N_FREE_LEVELS], (double)pc[PROF_FREE_LEVELS] / std::max<unsigned long long>(pc[PROF_N_FREE_LEVELS], 1),
              (double)pc[PROF_N_PASSES] / std::max<unsigned long long>(pc[PROF_N_AES_LEVELS] + pc[PROF_N_FREE_LEVELS], 1),
              (double)pc[PROF_GATHER] / std::max<unsigned long long>(pc[PROF_N_ITEMS], 1),
              (double)pc[PROF_SCATTER] / std::max<unsigned long long>(pc[PROF_N_ITEMS], 1),
              (double)pc[PROF_COMPLETE] / std::max<unsigned long long>(pc[PROF_N_ITEMS], 1));
    }
    res->n_ciphertexts = g.total_ct;
    res->n_launches = launches;
    s->ct_valid = (s->ct_mode != GSV_CT_NONE) && !linked;
    return GSV_OK;
  } catch (const std::exception& e) {
    if (s->link) s->link->failed.store(true);
    return fail(GSV_ERR_CUDA, e.what());
  }
}

int gsv_session_set_ciphertext_files(gsv_session* s, const int* fds) {
  if (!s) return fail(GSV_ERR_INVALID, "null argument");
  if (!fds) {
    s->ct_sink_fds.clear();
    return GSV_OK;
  }
  if (s->ct_mode != GSV_CT_COMMIT_HOST && !(s->link && !s->link_garbler))
    return fail(GSV_ERR_INVALID, "ciphertext files are written by the host drain: GSV_CT_COMMIT_HOST sessions and linked evaluators");
  s->ct_sink_fds.assign(fds, fds + s->B);
  return GSV_OK;
}

int gsv_session_expand_seeds(gsv_session* s, const uint64_t* seeds, gsv_garble_result* res) {
  if (!s || !seeds || !res) return fail(GSV_ERR_INVALID, "null argument");
  try {
    CUDA_TRY(cudaSetDevice(s->device));
    const gsv::Program& g = s->prog->prog;
    const uint32_t B = s->B;
    CUDA_TRY(cudaMemcpyAsync(s->d_seeds.p, seeds, (size_t)B * 8, cudaMemcpyHostToDevice, s->stream));
    const uint32_t n_blocks = (3 + g.n_inputs + 3) / 4;
    const size_t total = (size_t)B * n_blocks;
    k_seed_expand<0><<<(unsigned)((total + 127) / 128), 128, 0, s->stream>>>(s->d_seeds.p, B, s->G, g.n_inputs, g.n_global_slots,
                                                                              s->d_labels.p, s->d_delta.p);
    CUDA_TRY(cudaGetLastError());
    uint32_t launches = 1;
    if (res->delta) CUDA_TRY(cudaMemcpyAsync(res->delta, s->d_delta.p, (size_t)B * 16, cudaMemcpyDeviceToHost, s->stream));
    auto gather = [&](uint32_t first, uint32_t n, uint8_t* host_out) {
      if (!host_out || n == 0) return;
      const size_t cnt = (size_t)B * n;
      k_gather_slots<0><<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(s->d_labels.p, nullptr, s->d_gather_slots.p + first, n, B,
                                                                              s->G, g.n_global_slots, s->d_io.p, nullptr);
      CUDA_TRY(cudaGetLastError());
      launches++;
      CUDA_TRY(cudaMemcpyAsync(host_out, s->d_io.p, cnt * 16, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
    };
    gather(0, 1, res->false_label0);
    gather(1, 1, res->true_label0);
    gather(2, g.n_inputs, res->input_label0);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    res->n_launches = launches;
    res->n_ciphertexts = g.total_ct;
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_CUDA, e.what());
  }
}

int gsv_session_read_ciphertexts(gsv_session* s, uint32_t instance, uint64_t first, uint64_t count, uint8_t* out) {
  if (!s || !out) return fail(GSV_ERR_INVALID, "null argument");
  try {
    CUDA_TRY(cudaSetDevice(s->device));
    const gsv::Program& g = s->prog->prog;
    // GSV_CT_COMMIT_HOST lays its buffer out for the host drain (not [position][instance]): never readable here
    if (!s->ct_valid || s->ct_mode == GSV_CT_NONE || s->ct_mode == GSV_CT_COMMIT_HOST || s->ct_ring)
      return fail(GSV_ERR_INVALID, "no ciphertext stream kept (use GSV_CT_KEEP)");
    if (instance >= s->B || first + count > g.total_ct) return fail(GSV_ERR_INVALID, "range out of bounds");
    if (count == 0) return GSV_OK;
    if (s->d_stage.n < count) s->d_stage.alloc(count);
    k_ct_extract<0><<<(unsigned)((count + 255) / 256), 256, 0, s->stream>>>(s->d_ct.p, s->B, instance, first, count, s->d_stage.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, s->d_stage.p, count * 16, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_CUDA, e.what());
  }
}

int gsv_evaluate_batch(gsv_session* s, int hasher, gsv_evaluate_io* io) {
  if (!s || !io || !io->true_label || !io->false_label) return fail(GSV_ERR_INVALID, "null argument");
  try {
    CUDA_TRY(cudaSetDevice(s->device));
    const gsv::Program& g = s->prog->prog;
    const uint32_t B = s->B;
    const uint32_t n_in = g.n_inputs, n_out = (uint32_t)g.output_slots.size();
    if (n_in && (!io->input_active || !io->input_bits)) return fail(GSV_ERR_INVALID, "missing inputs");
    uint32_t launches = 0;
    uint64_t ct_avail = g.total_ct;
    bool host_fed = false;     // ciphertexts fed from host streams through a ring while the kernel runs
    uint64_t fed_ring = 0;
    const bool linked = s->link && !s->link_garbler;
    if (linked && getenv("GSV_LINK_DEBUG")) fprintf(stderr, "[link] evaluator: call entered\n");
    if (linked) {
      if (io->ct_streams) return fail(GSV_ERR_INVALID, "a linked evaluator takes its ciphertexts from the garbler session");
    } else if (io->ct_streams) {
      // FileSource-style host streams (ciphertext_source.rs:35-106).  Small: upload and interleave, the kernel
      // reads a resident [position][B] buffer.  Large (or io->ct_ring_log2 set): the streams are FED through a
      // ring while the kernel runs (below), so 47.7 GB-per-instance verifier streams never have to fit in HBM.
      ct_avail = std::min<uint64_t>(io->ct_stream_len, g.total_ct);
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
      const uint64_t resident_bytes = std::max<uint64_t>(g.total_ct, 1) * B * 16;
      host_fed = io->ct_ring_log2 != 0 || (s->d_ct.n < (size_t)std::max<uint64_t>(g.total_ct, 1) * B && resident_bytes > free_b / 2);
      if (!host_fed) {
        if (s->d_ct.n < (size_t)std::max<uint64_t>(g.total_ct, 1) * B) s->d_ct.alloc((size_t)std::max<uint64_t>(g.total_ct, 1) * B);
        if (ct_avail) {
          if (s->d_stage.n < ct_avail) s->d_stage.alloc(ct_avail);
          for (uint32_t i = 0; i < B; i++) {
            CUDA_TRY(cudaMemcpyAsync(s->d_stage.p, io->ct_streams[i], ct_avail * 16, cudaMemcpyHostToDevice, s->stream));
            k_ct_insert<0><<<(unsigned)((ct_avail + 255) / 256), 256, 0, s->stream>>>(s->d_ct.p, B, i, 0, ct_avail, s->d_stage.p);
            CUDA_TRY(cudaGetLastError());
            launches++;
          }
        }
      } else {
        uint64_t max_task_ct = 1;
        for (const auto& t : g.tasks) max_task_ct = std::max<uint64_t>(max_task_ct, t.n_ct);
        fed_ring = io->ct_ring_log2 ? (1ull << io->ct_ring_log2) : std::max<uint64_t>((uint64_t)(free_b * 0.4) / ((uint64_t)B * 16), 1);
        fed_ring = std::min<uint64_t>(fed_ring, std::max<uint64_t>(g.total_ct, 1));
        if (fed_ring < 2 * max_task_ct && fed_ring < g.total_ct) throw std::runtime_error("ciphertext ring smaller than two tasks");
        if (s->d_ct.n < (size_t)fed_ring * B) s->d_ct.alloc((size_t)fed_ring * B);
        ensure_host_chain_resources(s);
        if (s->d_park_head.n == 0) {
          s->park_q = std::max<uint64_t>(fed_ring / 64, 1);
          s->d_park_head.alloc((size_t)(std::max<uint64_t>(g.total_ct, 1) / s->park_q + 2));
          s->d_park_next.alloc(std::max<size_t>(g.calls.size() * (size_t)s->n_groups, 1));
        }
      }
      ct_avail = io->ct_stream_len;
    } else if (!s->ct_valid || s->ct_ring || s->ct_mode == GSV_CT_COMMIT_HOST) {
      return fail(GSV_ERR_INVALID, "no ciphertext stream in the session (garble with GSV_CT_KEEP first)");
    }
    ensure_eval_buffers(s);  // no-op after the first call (and done at link time for linked evaluators)
    // inputs -> device
    DevBuf<uint4>&d_true = s->d_ev_true, &d_false = s->d_ev_false, &d_in = s->d_ev_in;
    DevBuf<uint8_t>& d_bits = s->d_ev_bits;
    CUDA_TRY(cudaMemcpyAsync(d_true.p, io->true_label, (size_t)B * 16, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(d_false.p, io->false_label, (size_t)B * 16, cudaMemcpyHostToDevice, s->stream));
    if (n_in) {
      CUDA_TRY(cudaMemcpyAsync(d_in.p, io->input_active, (size_t)B * n_in * 16, cudaMemcpyHostToDevice, s->stream));
      CUDA_TRY(cudaMemcpyAsync(d_bits.p, io->input_bits, (size_t)B * n_in, cudaMemcpyHostToDevice, s->stream));
    }
    s->epoch++;
    CUDA_TRY(cudaMemsetAsync(s->d_ctrl.p, 0, 16, s->stream));
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    {
      const size_t total = (size_t)B * (n_in + 2);
      k_scatter_inputs<0><<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
          d_in.p, d_bits.p, d_true.p, d_false.p, n_in, B, s->G, g.n_global_slots, s->d_labels.p, s->d_vals.p);
      CUDA_TRY(cudaGetLastError());
      launches++;
    }
    CUDA_TRY(cudaEventRecord(s->ev[1], s->stream));
    EngineParams p = make_params(s);
    p.ct_capacity = ct_avail;
    p.flow_control = 0;
    p.ct_ring = 0;
    std::vector<uint8_t> link_commits;
    if (linked) {
      // the ring is filled by the garbler's kernel; admit items against what it has published, publish what
      // this kernel has consumed
      gsv_link& L = *s->link;
      const auto t_wait = std::chrono::steady_clock::now();
      double start_limit_s = 60.0;
      if (const char* e = getenv("GSV_LINK_TIMEOUT_S")) start_limit_s = std::max(1.0, atof(e));
      while (L.runs_started.load() <= L.runs_consumed.load()) {  // the garbler resets the progress words first
        if (L.failed.load() || std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wait).count() > start_limit_s)
          return fail(GSV_ERR_INVALID, "the linked garbler did not start (call gsv_garble_batch concurrently)");
        std::this_thread::sleep_for(std::chrono::microseconds(100));
      }
      L.runs_consumed.fetch_add(1);
      if (getenv("GSV_LINK_DEBUG")) fprintf(stderr, "[link] evaluator: garbler has started\n");
      p.ct = L.ring;
      p.ct_ring = L.ring_positions;
      p.ct_pos_stride = 4;
      p.ct_quad_stride = L.cap * 4;
      p.ct_qshift = 2;
      p.flow_control = 1;
      p.limit_add = p.free_until = 0;
      p.n_progress = 0;
      p.ext_progress = L.ready_e;
      p.host_chain = 1;
      p.host_ready = L.done_e;
      p.ct_sys = 1;
      p.n_chain_warps = 1;
      p.n_chain_ctas = 1;
    }
    if (host_fed) {
      // instance-major ring [instance][position]; the feeder below publishes how much of the stream has landed in
      // hc_ready[0], the kernel's publisher warp what it has consumed in hc_ready[4] (mapped host memory)
      volatile unsigned long long* w = s->hc_ready;
      w[0] = w[4] = 0;
      std::atomic_thread_fence(std::memory_order_seq_cst);
      p.ct = s->d_ct.p;
      p.ct_ring = fed_ring >= g.total_ct ? 0 : fed_ring;
      p.ct_pos_stride = 1;
      p.ct_quad_stride = fed_ring;
      p.ct_qshift = 0;
      p.flow_control = 1;
      p.limit_add = p.free_until = 0;
      p.n_progress = 0;
      p.ext_progress = s->hc_ready_dev;
      p.host_chain = 1;
      p.host_ready = s->hc_ready_dev + 4;
      p.n_chain_warps = 1;
      p.n_chain_ctas = 1;
    }
    launch_engine<1>(s, hasher, p);
    launches += 2;
    if (s->d_slot_flags.n) launches++;  // k_mark_slots (pipelined plan)
    CUDA_TRY(cudaEventRecord(s->ev[2], s->stream));
    std::vector<uint8_t> fed_commits;
    if (host_fed) {
      // ---- FileSource: feed the ring and hash the streams on host threads at the same time
      const uint64_t avail_total = std::min<uint64_t>(io->ct_stream_len, g.total_ct);
      volatile unsigned long long* w = s->hc_ready;
      fed_commits.assign((size_t)B * 16, 0);
      std::vector<std::thread> folders;
      if (io->ct_commit && gsv::host_chain_available()) {
        const uint32_t T = std::max<uint32_t>(1, std::min<uint32_t>(s->host_threads, (B + 3) / 4));
        for (uint32_t t = 0; t < T; t++)
          folders.emplace_back([&, t]() {
            const uint32_t i0 = (uint32_t)((uint64_t)B * t / T), i1 = (uint32_t)((uint64_t)B * (t + 1) / T);
            if (i1 > i0) gsv::host_chain_fold_streams(fed_commits.data() + (size_t)i0 * 16, io->ct_streams + i0, 0, avail_total, i1 - i0);
          });
      }
      std::string err;
      try {
        const uint64_t ring = fed_ring;
        const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(ring / 4, (8u << 20) / 16));
        uint64_t fed = 0;
        auto last_move = std::chrono::steady_clock::now();
        while (fed < avail_total) {
          const uint64_t done = p.ct_ring ? w[4] : 0;
          const uint64_t room_end = p.ct_ring ? done + ring : avail_total;
          uint64_t n = std::min<uint64_t>({chunk, avail_total - fed, room_end > fed ? room_end - fed : 0, ring - fed % ring});
          if (n == 0) {
            const cudaError_t q = cudaStreamQuery(s->stream);
            if (q != cudaErrorNotReady) throw std::runtime_error(q == cudaSuccess ? "evaluate kernel ended before the stream was consumed"
                                                                                  : cudaGetErrorString(q));
            if (std::chrono::steady_clock::now() - last_move > std::chrono::seconds(120)) throw std::runtime_error("ciphertext feed stalled");
            std::this_thread::sleep_for(std::chrono::microseconds(50));
            continue;
          }
          last_move = std::chrono::steady_clock::now();
          for (uint32_t i = 0; i < B; i++)
            CUDA_TRY(cudaMemcpyAsync(s->d_ct.p + (size_t)i * ring + fed % ring, io->ct_streams[i] + fed * 16, n * 16,
                                     cudaMemcpyHostToDevice, s->copy_stream));
          CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
          fed += n;
          std::atomic_thread_fence(std::memory_order_seq_cst);
          w[0] = fed;
        }
        // a short stream: let the kernel run into the exhaustion check instead of waiting for more
        if (avail_total < g.total_ct) w[0] = ~0ull >> 1;
      } catch (const std::exception& e) {
        err = e.what();
        w[0] = ~0ull >> 1;
      }
      for (auto& t : folders) t.join();
      if (!err.empty()) {
        cudaStreamSynchronize(s->stream);
        throw std::runtime_error(err);
      }
    }
    if (linked) {
      // hash what arrives (the reference's evaluator hashes the channel while evaluating,
      // examples/groth16_garble.rs:204-216): drain the ring to the host fold threads; a ring position is
      // released to the garbler once it has been drained AND consumed by the evaluate kernel
      gsv_link& L = *s->link;
      std::atomic<uint64_t> copied_done{0};
      volatile unsigned long long* w = L.words;
      DrainSource src;
      src.ring = L.ring;
      src.cap = L.cap;
      src.ring_positions = L.ring_positions;
      src.ready = w;
      src.watch = s->stream;
      src.copied_done = &copied_done;
      // watchdog: a stream that stops moving for GSV_LINK_TIMEOUT_S (default 60 s) is reported, not waited for
      double stall_limit_s = 60.0;
      if (const char* e = getenv("GSV_LINK_TIMEOUT_S")) stall_limit_s = std::max(1.0, atof(e));
      auto last_move = std::chrono::steady_clock::now();
      uint64_t last_sum = 0;
      std::string stall_msg;
      const bool dbg = getenv("GSV_LINK_DEBUG") != nullptr;
      auto last_dbg = std::chrono::steady_clock::now();
      if (dbg) fprintf(stderr, "[link] evaluator: kernel launched, draining\n");
      src.tick = [&]() {
        const uint64_t cd = copied_done.load(std::memory_order_acquire), done = w[8], rdy = w[0];
        if (dbg && std::chrono::steady_clock::now() - last_dbg > std::chrono::seconds(1)) {
          last_dbg = std::chrono::steady_clock::now();
          fprintf(stderr, "[link] published %llu consumed %llu drained %llu released %llu\n", (unsigned long long)rdy,
                  (unsigned long long)done, (unsigned long long)cd, (unsigned long long)w[16]);
        }
        const uint64_t rel = std::min<uint64_t>(cd, done);
        if (rel > w[16]) {
          std::atomic_thread_fence(std::memory_order_seq_cst);
          w[16] = rel;
        }
        const auto now = std::chrono::steady_clock::now();
        if (cd + done + rdy != last_sum) {
          last_sum = cd + done + rdy;
          last_move = now;
        } else if (std::chrono::duration<double>(now - last_move).count() > stall_limit_s) {
          stall_msg = "linked stream stalled: garbler published " + std::to_string(rdy) + ", evaluate kernel consumed " +
                      std::to_string(done) + ", drained " + std::to_string(cd) + " of " + std::to_string(g.total_ct) + " ciphertexts";
          return false;
        }
        return !L.failed.load();
      };
      link_commits.resize((size_t)B * 16);
      try {
        run_host_chain(s, src, link_commits.data(), nullptr);
      } catch (const std::exception& e) {
        uint32_t sc[4] = {0, 0, 0, 0};
        cudaMemcpyAsync(sc, s->d_ctrl.p + 4, 16, cudaMemcpyDeviceToHost, s->copy_stream);
        cudaStreamSynchronize(s->copy_stream);
        L.failed.store(true);
        w[16] = ~0ull >> 1;  // let the garbler run to completion
        w[0] = ~0ull >> 1;   // ... and this session's own kernel (it then reads whatever is in the ring)
        const auto t0 = std::chrono::steady_clock::now();
        while (cudaStreamQuery(s->stream) == cudaErrorNotReady && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(10))
          std::this_thread::sleep_for(std::chrono::milliseconds(1));
        throw std::runtime_error((stall_msg.empty() ? std::string(e.what()) : stall_msg) + "; evaluate kernel: queue head " +
                                 std::to_string(sc[0]) + " tail " + std::to_string(sc[1]) + ", items completed " + std::to_string(sc[2]) +
                                 " of " + std::to_string(g.calls.size() * (size_t)s->n_groups) + ", buckets released " + std::to_string(sc[3]) +
                                 (cudaStreamQuery(s->stream) == cudaErrorNotReady ? " (kernel still running)" : ""));
      }
      // keep returning ring space until the kernel has consumed the tail
      while (cudaStreamQuery(s->stream) == cudaErrorNotReady) {
        if (!src.tick()) {
          uint32_t sc[4] = {0, 0, 0, 0};
          cudaMemcpyAsync(sc, s->d_ctrl.p + 4, 16, cudaMemcpyDeviceToHost, s->copy_stream);
          cudaStreamSynchronize(s->copy_stream);
          L.failed.store(true);
          w[16] = ~0ull >> 1;
          throw std::runtime_error(stall_msg + "; evaluate kernel still running: queue head " + std::to_string(sc[0]) + " tail " +
                                   std::to_string(sc[1]) + ", items completed " + std::to_string(sc[2]) + " of " +
                                   std::to_string(g.calls.size() * (size_t)s->n_groups) + ", buckets released " + std::to_string(sc[3]));
        }
        std::this_thread::sleep_for(std::chrono::microseconds(50));
      }
    }
    // the evaluator's own chain hash over what it consumed (FileSource hashes while reading)
    const uint64_t used = std::min<uint64_t>(ct_avail, g.total_ct);
    if (io->ct_commit && linked) {
      memcpy(io->ct_commit, link_commits.data(), link_commits.size());
    } else if (io->ct_commit && host_fed) {
      memcpy(io->ct_commit, fed_commits.data(), fed_commits.size());
    } else if (io->ct_commit) {
    k_chain<0><<<(B + 31) / 32, 32, AES_TABLE_BYTES, s->stream>>>(s->d_ct.p, used, B, s->d_commit.p);
    CUDA_TRY(cudaGetLastError());
    launches++;
    }
    CUDA_TRY(cudaEventRecord(s->ev[3], s->stream));
    uint32_t ctrl[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(ctrl, s->d_ctrl.p, 16, cudaMemcpyDeviceToHost, s->stream));
    if (io->ct_commit && !linked && !host_fed) CUDA_TRY(cudaMemcpyAsync(io->ct_commit, s->d_commit.p, (size_t)B * 16, cudaMemcpyDeviceToHost, s->stream));
    if (n_out && (io->output_active || io->output_bits)) {
      const size_t total = (size_t)B * n_out;
      k_gather_slots<0><<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
          s->d_labels.p, s->d_vals.p, s->d_output_slots.p, n_out, B, s->G, g.n_global_slots, s->d_io.p, s->d_io_bits.p);
      CUDA_TRY(cudaGetLastError());
      launches++;
      if (io->output_active) CUDA_TRY(cudaMemcpyAsync(io->output_active, s->d_io.p, total * 16, cudaMemcpyDeviceToHost, s->stream));
      if (io->output_bits) CUDA_TRY(cudaMemcpyAsync(io->output_bits, s->d_io_bits.p, total, cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaEventElapsedTime(&io->ms_evaluate, s->ev[1], s->ev[2]);
    cudaEventElapsedTime(&io->ms_commit, s->ev[2], s->ev[3]);
    cudaEventElapsedTime(&io->ms_total, s->ev[0], s->ev[3]);
    io->n_launches = launches;
    if (ctrl[1]) return fail(GSV_ERR_CT_EXHAUSTED, "Ciphertext source exhausted");
    return GSV_OK;
  } catch (const std::exception& e) {
    if (s->link) {
      s->link->failed.store(true);
      reinterpret_cast<volatile unsigned long long*>(s->link->words)[16] = ~0ull >> 1;
    }
    return fail(GSV_ERR_CUDA, e.what());
  }
}

int gsv_execute_batch(gsv_session* s, const uint8_t* input_bits, uint32_t n_exec, uint8_t* output_bits, float* ms) {
  if (!s || !input_bits || !output_bits) return fail(GSV_ERR_INVALID, "null argument");
  try {
    CUDA_TRY(cudaSetDevice(s->device));
    const gsv::Program& g = s->prog->prog;
    const uint32_t B = s->B, n_in = g.n_inputs, n_out = (uint32_t)g.output_slots.size();
    if (n_exec == 0 || n_exec > 128u * B) return fail(GSV_ERR_INVALID, "a session of B instances executes at most 128 * B inputs at once");
    if (s->lane_mode) return fail(GSV_ERR_INVALID, "execute mode runs on the levelised kernel (exec_mode 1)");
    ensure_eval_buffers(s);
    // bit-slice: execution e -> instance e / 128, bit e % 128 of the wire's 16-byte slot
    std::vector<uint32_t> packed((size_t)B * n_in * 4, 0u);
    for (uint32_t e = 0; e < n_exec; e++) {
      const uint32_t inst = e >> 7, w = (e >> 5) & 3u, bit = e & 31u;
      const uint8_t* row = input_bits + (size_t)e * n_in;
      uint32_t* dst = packed.data() + (size_t)inst * n_in * 4 + w;
      for (uint32_t j = 0; j < n_in; j++) dst[4 * j] |= (uint32_t)(row[j] & 1u) << bit;
    }
    std::vector<uint32_t> ones((size_t)B * 4, 0xFFFFFFFFu);
    if (n_in) CUDA_TRY(cudaMemcpyAsync(s->d_ev_in.p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_ev_true.p, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->d_ev_false.p, 0, (size_t)B * 16, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->d_ev_bits.p, 0, std::max<size_t>((size_t)B * n_in, 1), s->stream));
    s->epoch++;
    CUDA_TRY(cudaMemsetAsync(s->d_ctrl.p, 0, 16, s->stream));
    {
      const size_t total = (size_t)B * (n_in + 2);
      k_scatter_inputs<0><<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
          s->d_ev_in.p, s->d_ev_bits.p, s->d_ev_true.p, s->d_ev_false.p, n_in, B, s->G, g.n_global_slots, s->d_labels.p, s->d_vals.p);
      CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(s->ev[1], s->stream));
    EngineParams p = make_params(s);
    p.flow_control = 0;
    p.ct_ring = 0;
    p.host_chain = 0;
    launch_engine<2>(s, GSV_HASH_AES, p);
    CUDA_TRY(cudaEventRecord(s->ev[2], s->stream));
    std::vector<uint32_t> out((size_t)B * std::max<uint32_t>(n_out, 1) * 4, 0u);
    if (n_out) {
      const size_t total = (size_t)B * n_out;
      k_gather_slots<0><<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
          s->d_labels.p, nullptr, s->d_output_slots.p, n_out, B, s->G, g.n_global_slots, s->d_io.p, nullptr);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(out.data(), s->d_io.p, total * 16, cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (ms) cudaEventElapsedTime(ms, s->ev[1], s->ev[2]);
    for (uint32_t e = 0; e < n_exec; e++) {
      const uint32_t inst = e >> 7, w = (e >> 5) & 3u, bit = e & 31u;
      // (outputs wired straight to a constant are gathered from slots 0 / 1 like any other slot)
      for (uint32_t k = 0; k < n_out; k++)
        output_bits[(size_t)e * n_out + k] = (uint8_t)((out[((size_t)inst * n_out + k) * 4 + w] >> bit) & 1u);
    }
    s->ct_valid = false;  // the label slots now hold plaintext slices
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(GSV_ERR_CUDA, e.what());
  }
}

int gsv_commit_labels(int device, const uint8_t* labels, uint64_t n, uint8_t* out) {
  if (!labels || !out) return fail(GSV_ERR_INVALID, "null argument");
  try {
    ensure_device(device);
    if (n == 0) return GSV_OK;
    DevBuf<uint4> d_in, d_out;
    d_in.alloc(n);
    d_out.alloc(n);
    CUDA_TRY(cudaMemcpy(d_in.p, labels, n * 16, cudaMemcpyHostToDevice));
    k_commit_labels<0><<<(unsigned)std::min<uint64_t>((n + 255) / 256, 296), 256, AES_TABLE_BYTES>>>(d_in.p, n, d_out.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, d_out.p, n * 16, cudaMemcpyDeviceToHost));
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(std::string(e.what()).find("no CUDA device") != std::string::npos ? GSV_ERR_NO_DEVICE : GSV_ERR_CUDA, e.what());
  }
}

int gsv_hash_blocks(int device, int hasher, const uint8_t* x, const uint64_t* gid, uint64_t n, uint8_t* out) {
  if (!x || !gid || !out) return fail(GSV_ERR_INVALID, "null argument");
  try {
    ensure_device(device);
    if (n == 0) return GSV_OK;
    DevBuf<uint4> d_in, d_out;
    DevBuf<unsigned long long> d_gid;
    d_in.alloc(n);
    d_out.alloc(n);
    d_gid.alloc(n);
    CUDA_TRY(cudaMemcpy(d_in.p, x, n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_gid.p, gid, n * 8, cudaMemcpyHostToDevice));
    const unsigned hb_grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 296);
    if (hasher == GSV_HASH_AES) k_hash_blocks<HASH_AES><<<hb_grid, 256, AES_TABLE_BYTES>>>(d_in.p, d_gid.p, n, d_out.p);
    else k_hash_blocks<HASH_BLAKE3><<<hb_grid, 256, AES_TABLE_BYTES>>>(d_in.p, d_gid.p, n, d_out.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, d_out.p, n * 16, cudaMemcpyDeviceToHost));
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(std::string(e.what()).find("no CUDA device") != std::string::npos ? GSV_ERR_NO_DEVICE : GSV_ERR_CUDA, e.what());
  }
}

int gsv_bench_hash_latency(int device, int hasher, uint32_t warps_per_sm, uint64_t n, double* cycles_per_hash) {
  if (!cycles_per_hash || warps_per_sm == 0 || warps_per_sm > 32 || n == 0) return fail(GSV_ERR_INVALID, "bad argument");
  try {
    ensure_device(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    DevBuf<unsigned long long> out;
    out.alloc(2);
    if (hasher == GSV_HASH_AES) k_hash_latency<HASH_AES><<<prop.multiProcessorCount, 32 * warps_per_sm, AES_TABLE_BYTES>>>(n, out.p);
    else k_hash_latency<HASH_BLAKE3><<<prop.multiProcessorCount, 32 * warps_per_sm, AES_TABLE_BYTES>>>(n, out.p);
    CUDA_TRY(cudaGetLastError());
    unsigned long long cyc = 0;
    CUDA_TRY(cudaMemcpy(&cyc, out.p, 8, cudaMemcpyDeviceToHost));
    *cycles_per_hash = (double)cyc / (double)n;
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(std::string(e.what()).find("no CUDA device") != std::string::npos ? GSV_ERR_NO_DEVICE : GSV_ERR_CUDA, e.what());
  }
}

int gsv_bench_hash(int device, int hasher, uint64_t n_blocks, int iters, double* blocks_per_s) {
  if (!blocks_per_s) return fail(GSV_ERR_INVALID, "null argument");
  try {
    ensure_device(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const unsigned grid = (unsigned)prop.multiProcessorCount * 2, block = 512;
    const unsigned long long threads = (unsigned long long)grid * block;
    unsigned long long per_thread = std::max<unsigned long long>(1, n_blocks / (2 * threads));
    DevBuf<uint4> sink;
    sink.alloc(1);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < std::max(iters, 1) + 1; it++) {
      CUDA_TRY(cudaEventRecord(e0));
      if (hasher == GSV_HASH_AES) k_bench_hash<HASH_AES><<<grid, block, AES_TABLE_BYTES>>>(per_thread, sink.p);
      else k_bench_hash<HASH_BLAKE3><<<grid, block, AES_TABLE_BYTES>>>(per_thread, sink.p);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(e1));
      CUDA_TRY(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *blocks_per_s = (double)(2 * per_thread * threads) / (best * 1e-3);
    return GSV_OK;
  } catch (const std::exception& e) {
    return fail(std::string(e.what()).find("no CUDA device") != std::string::npos ? GSV_ERR_NO_DEVICE : GSV_ERR_CUDA, e.what());
  }
}

}  // extern "C"

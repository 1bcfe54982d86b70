// planner.cpp -- cuts the template DAG into shared-memory-sized tasks, levelises them and
// emits the call list with global slots and dependencies.  See program.h.
#include <algorithm>
#include <iterator>
#include <atomic>
#include <memory>
#include <thread>
#include <cassert>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>

#include "program.h"

namespace gsv {

namespace {

constexpr uint32_t UNSET = 0xFFFFFFFEu;

// Levelise + slot-pack one flat SSA stream (ids: 0/1 consts, [2, 2+n_in) inputs, then defs).
Task compile_flat(const FlatStream& fs, const std::string& key, const PlanOptions& opt) {
  Task t;
  t.key = key;
  t.n_in = fs.n_inputs;
  const size_t ng = fs.type.size();
  t.n_gates_total = ng;
  const uint32_t nw = fs.n_wires;
  const uint32_t first_def = WIRE_MIN + fs.n_inputs;

  // ---- ASAP levels (inputs / constants at level 0)
  std::vector<uint32_t> wlevel(nw, 0);       // level at which a wire becomes available
  std::vector<uint32_t> asap(ng, 0), sched(ng, 0);
  std::vector<uint32_t> ct_off(ng, 0);
  uint32_t depth = 0;
  uint64_t n_ct = 0, n_live = 0;
  for (size_t g = 0; g < ng; g++) {
    if (fs.c[g] == WIRE_DEAD) continue;
    uint32_t l = 1 + std::max(wlevel[fs.a[g]], wlevel[fs.b[g]]);
    asap[g] = l;
    wlevel[fs.c[g]] = l;
    depth = std::max(depth, l);
    n_live++;
    if (!is_free(fs.type[g])) ct_off[g] = (uint32_t)n_ct++;
  }
  t.n_ct = n_ct;
  t.n_live = n_live;
  t.n_levels = depth;

  if (!opt.build_levelised) goto lane_form;
  {
  // ---- pipelining analysis: outputs at their ASAP level, the rest as late as that allows
  {
    std::vector<uint32_t> req(nw, depth);
    for (uint32_t o : fs.outputs)
      if (o != WIRE_DEAD && o >= first_def) req[o] = wlevel[o];
    for (size_t gi = ng; gi-- > 0;) {
      if (fs.c[gi] == WIRE_DEAD) continue;
      const uint32_t l = std::max(req[fs.c[gi]], asap[gi]);
      req[fs.a[gi]] = std::min(req[fs.a[gi]], l - 1);
      req[fs.b[gi]] = std::min(req[fs.b[gi]], l - 1);
    }
    t.pipe_in_need.assign(fs.n_inputs, depth + 1);
    for (uint32_t i = 0; i < fs.n_inputs; i++) t.pipe_in_need[i] = req[WIRE_MIN + i] + 1;
    for (uint32_t o : fs.outputs)
      if (o != WIRE_DEAD && o >= first_def) t.pipe_out_ready.push_back(wlevel[o]);
    t.pipe_depth = depth;
  }
  // ---- ALAP levels: as late as the consumers allow; sinks (unread wires) at `depth`.  Outputs: at `depth` too,
  // or -- when calls are pipelined (opt.pipeline) -- at their EARLIEST level, so that a consumer call can start on
  // them while this task is still running.
  if (opt.alap) {
    std::vector<uint32_t> req(nw, depth);  // latest level at which the wire must be available
    if (opt.pipeline)
      for (uint32_t o : fs.outputs)
        if (o != WIRE_DEAD && o >= first_def) req[o] = std::min(req[o], wlevel[o]);
    for (size_t gi = ng; gi-- > 0;) {
      if (fs.c[gi] == WIRE_DEAD) continue;
      uint32_t l = req[fs.c[gi]];
      sched[gi] = l;
      if (l < asap[gi]) throw std::logic_error("ALAP below ASAP in " + key);
      req[fs.a[gi]] = std::min(req[fs.a[gi]], l - 1);
      req[fs.b[gi]] = std::min(req[fs.b[gi]], l - 1);
    }
  } else {
    sched = asap;
  }

  // ---- order: by level, non-free first, then emission order
  std::vector<uint32_t> order;
  order.reserve(n_live);
  for (size_t g = 0; g < ng; g++)
    if (fs.c[g] != WIRE_DEAD) order.push_back((uint32_t)g);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    if (sched[x] != sched[y]) return sched[x] < sched[y];
    return (int)is_free(fs.type[x]) < (int)is_free(fs.type[y]);
  });

  // ---- last use per wire (in scheduled levels); outputs are pinned to the end
  const uint32_t PIN = depth + 1;
  std::vector<uint32_t> last_use(nw, 0), def_level(nw, 0);
  std::vector<uint8_t> is_read(nw, 0);
  for (uint32_t g : order) {
    last_use[fs.a[g]] = std::max(last_use[fs.a[g]], sched[g]);
    last_use[fs.b[g]] = std::max(last_use[fs.b[g]], sched[g]);
    is_read[fs.a[g]] = is_read[fs.b[g]] = 1;
    def_level[fs.c[g]] = sched[g];
    last_use[fs.c[g]] = std::max(last_use[fs.c[g]], sched[g]);
  }
  for (size_t j = 0; j < fs.outputs.size(); j++) {
    uint32_t o = fs.outputs[j];
    if (o != WIRE_DEAD && o >= first_def) last_use[o] = PIN;
  }

  // ---- interval colouring.  A slot whose wire is last read at level l is reusable by a gate
  // at level >= l+1 (readers and writers of one level run concurrently between two barriers).
  std::vector<uint32_t> slot(nw, UNSET);
  slot[0] = 0;
  slot[1] = 1;
  uint32_t next_slot = 2;
  std::vector<uint32_t> free_slots;
  std::vector<std::vector<uint32_t>> release(depth + 3);  // release[l]: wires whose last use is l
  t.in_slot.assign(fs.n_inputs, 0xFFFF);
  for (uint32_t i = 0; i < fs.n_inputs; i++) {
    uint32_t w = WIRE_MIN + i;
    if (!is_read[w]) continue;  // never read inside the task: no gather, no slot
    slot[w] = next_slot++;
    t.in_slot[i] = (uint16_t)slot[w];
    release[last_use[w]].push_back(w);
  }
  // Device levels are capped at LEVEL_WIDTH_MAX gates (a wider level becomes consecutive
  // sub-levels: extra barriers only) and the first record of each carries its width, so the
  // kernel streams the gate list through a small shared-memory ring without a level table.
  t.level_off.clear();
  t.gates.reserve(order.size());
  size_t oi = 0;
  for (uint32_t l = 1; l <= depth; l++) {
    for (uint32_t w : release[l - 1]) free_slots.push_back(slot[w]);
    size_t begin = oi;
    while (oi < order.size() && sched[order[oi]] == l) oi++;
    t.max_width = std::max<uint32_t>(t.max_width, (uint32_t)(oi - begin));
    for (size_t k = begin; k < oi; k++) {
      if ((k - begin) % LEVEL_WIDTH_MAX == 0) t.level_off.push_back((uint32_t)t.gates.size());
      uint32_t g = order[k];
      uint32_t c = fs.c[g];
      uint32_t s;
      if (!free_slots.empty()) {
        s = free_slots.back();
        free_slots.pop_back();
      } else {
        s = next_slot++;
      }
      slot[c] = s;
      if (last_use[c] != PIN) release[std::max(last_use[c], l)].push_back(c);
      if (slot[fs.a[g]] == UNSET || slot[fs.b[g]] == UNSET) throw std::logic_error("unslotted read in " + key);
      if (next_slot > 0xFFFF) throw std::length_error("task needs more than 65535 slots: " + key);
      DevGate dg;
      dg.a = (uint16_t)slot[fs.a[g]];
      dg.b = (uint16_t)slot[fs.b[g]];
      dg.c = (uint16_t)s;
      dg.type = fs.type[g];
      dg.flags = is_free(fs.type[g]) ? 0 : 1;
      dg.gid_off = g;
      dg.ct_off = ct_off[g];
      t.gates.push_back(dg);
    }
  }
  t.level_off.push_back((uint32_t)t.gates.size());
  t.n_levels = (uint32_t)t.level_off.size() - 1;
  {
    // ---- windows (call pipelining): inputs are gathered at the start of the window of W device levels that first
    // reads them, produced outputs are published at the end of the window that completes them.  Without
    // pipelining there is ONE window: everything gathered up front, everything published at the end.
    const uint32_t W = opt.pipeline ? std::max<uint32_t>(opt.window_levels, 1) : 0xFFFFFFFFu;
    t.window_levels = W;
    const uint32_t n_win = t.n_levels == 0 ? 1 : (opt.pipeline ? (t.n_levels + W - 1) / W : 1);
    std::vector<uint32_t> dev_level(t.gates.size(), 0);  // device level of every record
    for (uint32_t l = 0; l < t.n_levels; l++)
      for (uint32_t k = t.level_off[l]; k < t.level_off[l + 1]; k++) dev_level[k] = l;
    // first device level that reads each wire / device level that writes it (order[k] is the gate of record k)
    std::vector<uint32_t> first_read(nw, 0xFFFFFFFFu), written_at(nw, 0);
    for (size_t k = 0; k < t.gates.size(); k++) {
      const uint32_t g = order[k];
      first_read[fs.a[g]] = std::min(first_read[fs.a[g]], dev_level[k]);
      first_read[fs.b[g]] = std::min(first_read[fs.b[g]], dev_level[k]);
      written_at[fs.c[g]] = dev_level[k];
    }
    std::vector<uint32_t> in_need(fs.n_inputs, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < fs.n_inputs; i++)
      if (t.in_slot[i] != 0xFFFF) in_need[i] = first_read[WIRE_MIN + i];
    t.win_in_off.assign(n_win + 1, 0);
    t.win_out_off.assign(n_win + 1, 0);
    std::vector<std::vector<uint16_t>> ins(n_win), outs(n_win);
    for (uint32_t i = 0; i < fs.n_inputs; i++)
      if (t.in_slot[i] != 0xFFFF) ins[opt.pipeline && in_need[i] != 0xFFFFFFFFu ? in_need[i] / W : 0].push_back((uint16_t)i);
    if (fs.n_inputs > 0xFFFF) throw std::length_error("task with more than 65535 inputs: " + key);
    t.out_ready_level.clear();
    {
      uint32_t kk = 0;
      for (size_t j = 0; j < fs.outputs.size(); j++) {
        const uint32_t o = fs.outputs[j];
        if (o == WIRE_DEAD || o < first_def) continue;
        // device level of the gate that writes this output
        uint32_t lv = t.n_levels ? t.n_levels - 1 : 0;
        if (opt.pipeline) lv = written_at[o];
        t.out_ready_level.push_back(lv);
        outs[opt.pipeline ? lv / W : 0].push_back((uint16_t)kk);
        kk++;
      }
      if (kk > 0xFFFF) throw std::length_error("task with more than 65535 outputs: " + key);
    }
    for (uint32_t w = 0; w < n_win; w++) {
      t.win_in_off[w + 1] = t.win_in_off[w] + (uint32_t)ins[w].size();
      t.win_out_off[w + 1] = t.win_out_off[w] + (uint32_t)outs[w].size();
      t.win_in.insert(t.win_in.end(), ins[w].begin(), ins[w].end());
      t.win_out.insert(t.win_out.end(), outs[w].begin(), outs[w].end());
    }
    t.in_need_level = in_need;
  }
  // level header, carried by the level's first record: width - 1 in flags bits 1-7 and the number of
  // non-free gates (they come first) in the top byte of ct_off (a task has < 2^24 ciphertexts)
  if (n_ct >= (1u << 24)) throw std::length_error("task with 2^24 or more ciphertexts: " + key);
  for (uint32_t l = 0; l < t.n_levels; l++) {
    uint32_t nf = 0;
    for (uint32_t k = t.level_off[l]; k < t.level_off[l + 1] && !is_free(t.gates[k].type); k++) nf++;
    DevGate& h = t.gates[t.level_off[l]];
    h.flags |= (uint8_t)((t.level_off[l + 1] - t.level_off[l] - 1) << 1);
    h.ct_off |= nf << 24;
  }
  t.n_slots = next_slot;
  for (size_t j = 0; j < fs.outputs.size(); j++) {
    uint32_t o = fs.outputs[j];
    if (o == WIRE_DEAD || o < first_def) continue;
    t.out_slot.push_back((uint16_t)slot[o]);
  }
  }
lane_form:
  if (!opt.build_levelised) {
    t.n_levels = 0;
    t.n_slots = 2;
    t.level_off.assign(1, 0);
    t.in_slot.assign(fs.n_inputs, 0xFFFF);
    t.win_in_off.assign(2, 0);
    t.win_out_off.assign(2, 0);
  }
  // ---- produced outputs
  for (size_t j = 0; j < fs.outputs.size(); j++) {
    uint32_t o = fs.outputs[j];
    if (o == WIRE_DEAD || o < first_def) continue;  // dead / constant / passthrough input
    t.out_pos.push_back((uint32_t)j);
  }
  t.n_out = (uint32_t)t.out_pos.size();
  if (!opt.build_levelised) t.out_slot.assign(t.n_out, 0);

  // ---- lane-mode form: emission order, slots freed right after the last read (LIFO reuse)
  {
    std::vector<uint64_t> last(nw, 0);  // 1 + index of the last live gate reading the wire
    const uint64_t NEVER = ~0ull;
    for (size_t g = 0; g < ng; g++) {
      if (fs.c[g] == WIRE_DEAD) continue;
      last[fs.a[g]] = g + 1;
      last[fs.b[g]] = g + 1;
    }
    for (size_t j = 0; j < fs.outputs.size(); j++) {
      uint32_t o = fs.outputs[j];
      if (o != WIRE_DEAD && o >= first_def) last[o] = NEVER;
    }
    std::vector<uint32_t> sl(nw, UNSET);
    sl[0] = 0;
    sl[1] = 1;
    uint32_t next = 2;
    std::vector<uint32_t> freed;
    t.seq_in_slot.assign(fs.n_inputs, 0xFFFF);
    for (uint32_t i = 0; i < fs.n_inputs; i++) {
      uint32_t w = WIRE_MIN + i;
      if (!last[w]) continue;
      sl[w] = next++;
      t.seq_in_slot[i] = (uint16_t)sl[w];
    }
    t.seq_gates.reserve(n_live);
    for (size_t g = 0; g < ng; g++) {
      if (fs.c[g] == WIRE_DEAD) continue;
      const uint32_t a = fs.a[g], b = fs.b[g], c = fs.c[g];
      DevGate dg;
      dg.a = (uint16_t)sl[a];
      dg.b = (uint16_t)sl[b];
      if (a >= WIRE_MIN && last[a] == g + 1) freed.push_back(sl[a]);
      if (b >= WIRE_MIN && b != a && last[b] == g + 1) freed.push_back(sl[b]);
      uint32_t s;
      if (!freed.empty()) {
        s = freed.back();
        freed.pop_back();
      } else {
        s = next++;
      }
      if (next > 0xFFFF) throw std::length_error("task needs more than 65535 scratch slots: " + key);
      sl[c] = s;
      if (last[c] == 0) freed.push_back(s);  // written, never read, not an output
      dg.c = (uint16_t)s;
      dg.type = fs.type[g];
      dg.flags = is_free(fs.type[g]) ? 0 : 1;
      dg.gid_off = (uint32_t)g;
      dg.ct_off = ct_off[g];
      t.seq_gates.push_back(dg);
    }
    t.n_seq_slots = next;
    for (size_t j = 0; j < fs.outputs.size(); j++) {
      uint32_t o = fs.outputs[j];
      if (o == WIRE_DEAD || o < first_def) continue;
      t.seq_out_slot.push_back((uint16_t)sl[o]);
    }
    // lane-only plans: out_slot only serves as the "same produced wire" key of emit_call
    if (!opt.build_levelised) t.out_slot = t.seq_out_slot;
  }
  return t;
}

struct Planner {
  const Builder& b;
  PlanOptions opt;
  Program prog;
  // per builder template: -2 unknown, -1 structural, >=0 task index
  std::vector<int64_t> kind;
  // loose-gate runs of structural templates: (template, first item) -> task index
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> loose_task;
  std::vector<int64_t> producer;  // global wire -> producing call (-1: circuit input / constant)
  std::vector<uint64_t> mult;     // occurrences of each template in the flattened circuit
  uint32_t next_global = 0;
  uint64_t gid = 0, ct = 0;
  // task bodies compiled ahead of classify() by precompile(), on all host cores
  std::unordered_map<uint32_t, std::unique_ptr<Task>> compiled;
  std::unordered_set<uint32_t> too_big;  // bodies whose working set overflows the 16-bit slot ids

  // the size / sharing rules of classify() that need no compilation: true = structural
  bool structural_by_rule(uint32_t ti) const {
    const Template& t = b.tmpl(ti);
    const bool can_split = !t.calls.empty();
    if (t.total_gates > opt.max_task_gates && can_split) return true;
    return can_split && t.total_gates > opt.small_task_gates && mult[ti] < opt.min_shared_calls;
  }

  // Flattening + levelising + colouring a task body is independent of every other body, and is where
  // planning time goes.  Walk the template DAG top-down in waves: compile this wave's task candidates
  // in parallel, descend into the ones that turn out structural.  classify() then only looks up.
  void precompile(uint32_t root) {
    std::vector<uint8_t> seen(b.n_templates(), 0);
    std::vector<uint32_t> wave{root};
    seen[root] = 1;
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    while (!wave.empty()) {
      std::vector<uint32_t> cand;
      for (uint32_t ti : wave)
        if (!structural_by_rule(ti) && b.tmpl(ti).total_gates > 0) cand.push_back(ti);
      std::vector<std::unique_ptr<Task>> out(cand.size());
      std::vector<std::string> errs(cand.size());
      std::vector<uint8_t> oversize(cand.size(), 0);
      std::atomic<size_t> next{0};
      auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < cand.size();) {
          try {
            const Template& t = b.tmpl(cand[i]);
            FlatStream fs = flatten(b, cand[i], std::max<uint64_t>(opt.max_task_gates, t.total_gates) + 1);
            out[i] = std::make_unique<Task>(compile_flat(fs, t.key, opt));
          } catch (const std::length_error&) {
            out[i] = nullptr;  // working set beyond the 16-bit slot ids: cannot be a task
            oversize[i] = 1;
          } catch (const std::exception& e) {
            errs[i] = e.what();
          }
        }
      };
      std::vector<std::thread> th;
      for (unsigned k = 1; k < std::min<size_t>(nt, cand.size()); k++) th.emplace_back(work);
      work();
      for (auto& x : th) x.join();
      for (size_t i = 0; i < cand.size(); i++) {
        if (!errs[i].empty()) throw std::runtime_error(errs[i]);
        if (oversize[i]) too_big.insert(cand[i]);
        else compiled[cand[i]] = std::move(out[i]);
      }
      std::vector<uint32_t> nextw;
      for (uint32_t ti : wave) {
        const Template& t = b.tmpl(ti);
        bool structural = structural_by_rule(ti) || (too_big.count(ti) && !t.calls.empty());
        auto it = compiled.find(ti);
        if (!structural && it != compiled.end() && opt.build_levelised && it->second->n_slots > opt.max_task_slots &&
            !t.calls.empty())
          structural = true;
        if (!structural) continue;
        for (const CallRec& c : t.calls)
          if (!seen[c.tmpl]) {
            seen[c.tmpl] = 1;
            nextw.push_back(c.tmpl);
          }
      }
      wave.swap(nextw);
    }
  }

  Planner(const Builder& b_, const PlanOptions& o, uint32_t root) : b(b_), opt(o), kind(b_.n_templates(), -2) {
    // children are created before their parents, so a descending sweep propagates multiplicities
    mult.assign(b.n_templates(), 0);
    mult[root] = 1;
    for (uint32_t ti = (uint32_t)b.n_templates(); ti-- > 0;) {
      if (!mult[ti]) continue;
      for (const CallRec& c : b.tmpl(ti).calls) mult[c.tmpl] += mult[ti];
    }
    // a circuit that fits one task needs no sharing analysis
    if (b.tmpl(root).total_gates <= opt.max_task_gates) opt.min_shared_calls = 0;
  }

  int64_t classify(uint32_t ti) {
    if (kind[ti] != -2) return kind[ti];
    const Template& t = b.tmpl(ti);
    bool can_split = !t.calls.empty();
    if (t.total_gates > opt.max_task_gates && can_split) return kind[ti] = -1;
    // Sharing-aware cut: a component that occurs only a few times (e.g. a multiplier by one
    // vk-specific constant) but is built from widely shared children (bigint::add) is expanded into
    // those children, so the device program holds each distinct gate list once.
    if (can_split && t.total_gates > opt.small_task_gates && mult[ti] < opt.min_shared_calls) return kind[ti] = -1;
    if (t.total_gates == 0) {
      // pure re-wiring component (e.g. add_constant(0)): an empty task
      Task e;
      e.key = t.key;
      e.n_in = t.n_in;
      e.in_slot.assign(t.n_in, 0xFFFF);
      e.n_slots = 2;
      e.level_off.assign(1, 0);
      e.seq_in_slot.assign(t.n_in, 0xFFFF);
      e.n_seq_slots = 2;
      e.win_in_off.assign(2, 0);   // one empty window
      e.win_out_off.assign(2, 0);
      prog.tasks.push_back(std::move(e));
      return kind[ti] = (int64_t)prog.tasks.size() - 1;
    }
    if (too_big.count(ti) && can_split) return kind[ti] = -1;
    Task task;
    auto pre = compiled.find(ti);
    if (pre != compiled.end()) {
      task = std::move(*pre->second);
      compiled.erase(pre);
    } else {
      FlatStream fs = flatten(b, ti, std::max<uint64_t>(opt.max_task_gates, t.total_gates) + 1);
      task = compile_flat(fs, t.key, opt);
    }
    if (opt.build_levelised && task.n_slots > opt.max_task_slots && can_split) return kind[ti] = -1;
    // an unsplittable body above the slot budget stays a task: it runs in lane mode only (the
    // session refuses the levelised mode when max_task_slots does not fit shared memory)
    prog.tasks.push_back(std::move(task));
    return kind[ti] = (int64_t)prog.tasks.size() - 1;
  }

  void emit_call(uint32_t task_idx, const std::vector<uint32_t>& in_global, std::vector<uint32_t>& out_global) {
    const Task& task = prog.tasks[task_idx];
    Call c;
    c.task = task_idx;
    c.gid_base = gid;
    c.ct_base = ct;
    c.in_off = (uint32_t)prog.call_slots.size();
    std::vector<uint32_t> deps;
    for (uint32_t i = 0; i < task.n_in; i++) {
      uint32_t w = in_global[i];
      // inputs the task never reads are not gathered; keep the slot list dense anyway
      prog.call_slots.push_back(w == WIRE_DEAD ? 0u : w);
      if ((task.in_slot[i] != 0xFFFF || task.seq_in_slot[i] != 0xFFFF) && w != WIRE_DEAD && producer[w] >= 0) deps.push_back((uint32_t)producer[w]);
    }
    c.out_off = (uint32_t)prog.call_slots.size();
    uint32_t call_idx = (uint32_t)prog.calls.size();
    out_global.assign(task.n_out, UNSET);
    // duplicate produced outputs (same slot) share one global wire
    std::unordered_map<uint16_t, uint32_t> by_slot;
    for (uint32_t k = 0; k < task.n_out; k++) {
      auto it = by_slot.find(task.out_slot[k]);
      uint32_t w;
      if (it != by_slot.end()) w = it->second;
      else {
        w = next_global++;
        producer.push_back((int64_t)call_idx);
        by_slot.emplace(task.out_slot[k], w);
      }
      out_global[k] = w;
      prog.call_slots.push_back(w);
    }
    std::sort(deps.begin(), deps.end());
    deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
    c.dep_off = (uint32_t)prog.deps.size();
    c.n_deps = (uint32_t)deps.size();
    c.n_start_deps = opt.pipeline ? c.n_deps : 0;  // producers: start dependencies when calls are pipelined
    prog.deps.insert(prog.deps.end(), deps.begin(), deps.end());
    prog.max_call_deps = std::max(prog.max_call_deps, c.n_deps);
    prog.calls.push_back(c);
    gid += task.n_gates_total;
    ct += task.n_ct;
  }

  // Expands a structural template; returns the global wire of every callee output position.
  std::vector<uint32_t> expand(uint32_t ti, const std::vector<uint32_t>& in_global) {
    const Template& t = b.tmpl(ti);
    std::vector<uint32_t> l2g(t.n_wires, UNSET);
    l2g[0] = 0;
    l2g[1] = 1;
    for (uint32_t i = 0; i < t.n_in; i++) l2g[WIRE_MIN + i] = in_global[i];

    // last item that reads each local wire (for loose-run outputs)
    std::vector<uint32_t> last_read;
    bool has_gates = !t.gates.empty();
    if (has_gates) {
      last_read.assign(t.n_wires, 0);
      for (uint32_t k = 0; k < t.items.size(); k++) {
        const Item& it = t.items[k];
        if (!it.is_call) {
          const GateRec& g = t.gates[it.idx];
          last_read[g.a] = k + 1;
          last_read[g.b] = k + 1;
        } else {
          const CallRec& c = t.calls[it.idx];
          uint32_t n_in = b.tmpl(c.tmpl).n_in;
          for (uint32_t i = 0; i < n_in; i++) {
            Wire w = t.call_wires[c.in_off + i];
            if (w != WIRE_DEAD) last_read[w] = k + 1;
          }
        }
      }
      for (Wire o : t.outs)
        if (o != WIRE_DEAD) last_read[o] = (uint32_t)t.items.size() + 1;
    }

    size_t k = 0;
    while (k < t.items.size()) {
      const Item& it = t.items[k];
      if (!it.is_call) {
        // ---- a run of loose gates [k, e): wrap it into an anonymous task
        size_t e = k;
        while (e < t.items.size() && !t.items[e].is_call) e++;
        auto key = std::make_pair(ti, (uint32_t)k);
        // build the run's flat stream (also needed to bind wires on a cache hit)
        FlatStream fs;
        std::unordered_map<Wire, uint32_t> ext;  // external local wire -> flat input id
        std::vector<Wire> ext_order;
        std::unordered_map<Wire, uint32_t> def;  // local wire -> flat id (latest def)
        auto rd = [&](Wire w) -> uint32_t {
          if (w < WIRE_MIN) return w;
          auto d = def.find(w);
          if (d != def.end()) return d->second;
          auto x = ext.find(w);
          if (x != ext.end()) return x->second;
          uint32_t id = WIRE_MIN + (uint32_t)ext_order.size();
          ext.emplace(w, id);
          ext_order.push_back(w);
          return id;
        };
        // two passes: first discover the external inputs so that def ids start after them
        for (size_t q = k; q < e; q++) {
          const GateRec& g = t.gates[t.items[q].idx];
          if (g.a >= WIRE_MIN && !def.count(g.a)) rd(g.a);
          if (g.b >= WIRE_MIN && !def.count(g.b)) rd(g.b);
          if (g.c != WIRE_DEAD) def[g.c] = 0;
        }
        def.clear();
        fs.n_inputs = (uint32_t)ext_order.size();
        uint32_t next = WIRE_MIN + fs.n_inputs;
        std::vector<std::pair<Wire, uint32_t>> defs_in_order;
        for (size_t q = k; q < e; q++) {
          const GateRec& g = t.gates[t.items[q].idx];
          uint32_t fa = rd(g.a), fb = rd(g.b);
          uint32_t fc = WIRE_DEAD;
          if (g.c != WIRE_DEAD) {
            fc = next++;
            def[g.c] = fc;
          }
          fs.type.push_back(g.type);
          fs.a.push_back(fa);
          fs.b.push_back(fb);
          fs.c.push_back(fc);
        }
        fs.n_wires = next;
        std::vector<Wire> out_local;
        for (auto& d : def)
          if (last_read[d.first] > e) out_local.push_back(d.first);
        std::sort(out_local.begin(), out_local.end());
        for (Wire w : out_local) fs.outputs.push_back(def[w]);
        uint32_t task_idx;
        auto lt = loose_task.find(key);
        if (lt != loose_task.end()) task_idx = lt->second;
        else {
          Task task = compile_flat(fs, t.key + "@run" + std::to_string(k), opt);
          prog.tasks.push_back(std::move(task));
          task_idx = (uint32_t)prog.tasks.size() - 1;
          loose_task.emplace(key, task_idx);
        }
        std::vector<uint32_t> ins(ext_order.size());
        for (size_t i = 0; i < ext_order.size(); i++) {
          ins[i] = l2g[ext_order[i]];
          if (ins[i] == UNSET) throw std::logic_error("loose run reads unset wire in " + t.key);
        }
        std::vector<uint32_t> outs;
        emit_call(task_idx, ins, outs);
        const Task& task = prog.tasks[task_idx];
        for (uint32_t q = 0; q < task.n_out; q++) l2g[out_local[task.out_pos[q]]] = outs[q];
        k = e;
        continue;
      }
      const CallRec& c = t.calls[it.idx];
      const Template& ch = b.tmpl(c.tmpl);
      std::vector<uint32_t> ins(ch.n_in);
      for (uint32_t i = 0; i < ch.n_in; i++) {
        Wire w = t.call_wires[c.in_off + i];
        ins[i] = (w == WIRE_DEAD) ? WIRE_DEAD : l2g[w];
        if (ins[i] == UNSET) throw std::logic_error("call passes unset wire in " + t.key);
      }
      int64_t kd = classify(c.tmpl);
      if (kd >= 0) {
        std::vector<uint32_t> outs;
        emit_call((uint32_t)kd, ins, outs);
        const Task& task = prog.tasks[(size_t)kd];
        for (uint32_t q = 0; q < task.n_out; q++) {
          Wire p = t.call_wires[c.out_off + task.out_pos[q]];
          if (p == WIRE_DEAD || p < WIRE_MIN) throw std::logic_error("produced output bound to const/dead");
          l2g[p] = outs[q];
        }
      } else {
        std::vector<uint32_t> outs = expand(c.tmpl, ins);
        for (size_t j = 0; j < outs.size(); j++) {
          Wire p = t.call_wires[c.out_off + j];
          if (p == WIRE_DEAD || p < WIRE_MIN) continue;
          if (l2g[p] == UNSET) l2g[p] = outs[j];
        }
      }
      k++;
    }
    std::vector<uint32_t> res(t.outs.size());
    for (size_t j = 0; j < t.outs.size(); j++) {
      Wire o = t.outs[j];
      res[j] = (o == WIRE_DEAD) ? WIRE_DEAD : l2g[o];
      if (res[j] == UNSET) throw std::logic_error("unset output in " + t.key);
    }
    return res;
  }
};

}  // namespace

Task compile_task(const Builder& b, uint32_t tmpl, const PlanOptions& opt) {
  FlatStream fs = flatten(b, tmpl);
  return compile_flat(fs, b.tmpl(tmpl).key, opt);
}

Program plan_program(const Builder& b, uint32_t root, const PlanOptions& opt) {
  Planner p(b, opt, root);
  const Template& rt = b.tmpl(root);
  p.prog.n_inputs = rt.n_in;
  p.next_global = WIRE_MIN + rt.n_in;
  p.producer.assign(p.next_global, -1);
  std::vector<uint32_t> ins(rt.n_in);
  for (uint32_t i = 0; i < rt.n_in; i++) ins[i] = WIRE_MIN + i;
  std::vector<uint32_t> outs;
  p.precompile(root);
  int64_t kd = p.classify(root);
  if (kd >= 0) {
    // the whole circuit is one task
    std::vector<uint32_t> produced;
    p.emit_call((uint32_t)kd, ins, produced);
    const Task& task = p.prog.tasks[(size_t)kd];
    outs.assign(rt.outs.size(), UNSET);
    for (size_t j = 0; j < rt.outs.size(); j++) {
      Wire o = rt.outs[j];
      if (o == WIRE_DEAD) outs[j] = WIRE_DEAD;
      else if (o < WIRE_MIN) outs[j] = o;
      else if (o - WIRE_MIN < rt.n_in) outs[j] = ins[o - WIRE_MIN];
    }
    for (uint32_t q = 0; q < task.n_out; q++) outs[task.out_pos[q]] = produced[q];
  } else {
    outs = p.expand(root, ins);
  }
  Program& prog = p.prog;
  prog.output_slots = outs;
  prog.n_global_slots = p.next_global;
  if (opt.reuse_distance > 0 && !prog.calls.empty()) {
    // ---- global slot recycling.  Inter-task wires are SSA so far; pack them into recycled slots,
    // one contiguous BLOCK per call (its produced wires), freed after the block's last reader.
    // A call that takes over a block lists the readers of the previous occupant as WAR
    // dependencies.  Blocks are only recycled `reuse_distance` calls after their last reader, so
    // those dependencies are almost always satisfied long before the new owner starts.
    const uint32_t n_calls = (uint32_t)prog.calls.size();
    const uint32_t first_wire = WIRE_MIN + prog.n_inputs;
    const uint32_t n_w = p.next_global;
    std::vector<uint8_t> pinned(n_w, 0);
    for (uint32_t o : outs)
      if (o != WIRE_DEAD && o < n_w) pinned[o] = 1;
    std::vector<std::vector<uint32_t>> readers(n_calls);  // producer call -> reader calls (ascending)
    std::vector<uint32_t> block_last(n_calls, 0);
    std::vector<uint8_t> block_pinned(n_calls, 0);
    for (uint32_t c = 0; c < n_calls; c++) {
      const Call& call = prog.calls[c];
      const Task& task = prog.tasks[call.task];
      block_last[c] = std::max(block_last[c], c);
      for (uint32_t i = 0; i < task.n_in; i++) {
        if (task.seq_in_slot[i] == 0xFFFF) continue;
        uint32_t w = prog.call_slots[call.in_off + i];
        if (w < first_wire) continue;
        uint32_t pc = (uint32_t)p.producer[w];
        if (readers[pc].empty() || readers[pc].back() != c) readers[pc].push_back(c);
        block_last[pc] = std::max(block_last[pc], c);
      }
      for (uint32_t k = 0; k < task.n_out; k++)
        if (pinned[prog.call_slots[call.out_off + k]]) block_pinned[c] = 1;
    }
    // Which free block a call takes over matters: the WAR edges it inherits (the old occupant's readers)
    // are FALSE dependencies, and with first-fit they nearly doubled the verifier's critical path (two
    // independent sub-circuits emitted one after the other got chained through their slots).  So the
    // choice is depth-aware: finish[] is every call's earliest completion over the edges chosen so far
    // (in levels; gates for lane-only plans), and a call only takes a block whose readers have finished
    // by the time its own inputs are there -- the WAR edges then never delay anything.  If no such block
    // is free, fresh slots are used until max_global_slots, then the block that delays it least.
    struct FreeBlock { uint32_t base, producer, free_at; };
    std::map<uint32_t, std::vector<FreeBlock>> waiting;          // size class -> FIFO by free_at
    std::map<uint32_t, size_t> waiting_head;
    std::map<uint32_t, std::multimap<uint64_t, FreeBlock>> ready;  // size class -> readers' finish time -> block
    std::vector<std::vector<uint32_t>> release_at(n_calls + 1);
    std::vector<uint32_t> slot_of(n_w, UNSET), block_base(n_calls, 0), block_size(n_calls, 0);
    for (uint32_t w = 0; w < first_wire; w++) slot_of[w] = w;
    uint32_t next_slot = first_wire;
    std::vector<std::vector<uint32_t>> war(n_calls);
    // finish[] / begin[]: earliest completion / start of every call over the edges chosen so far.  Pipelined plans: a
    // consumer starts one window after its producers START and cannot finish before one window after they finish.
    std::vector<uint64_t> finish(n_calls, 0);
    // when every inter-task wire is available (in levels): a pipelined consumer needs input i only at the start of
    // the window that first reads it, a pipelined producer publishes output k at the end of the window that
    // completes it (+ 2 levels of flag-polling latency); without pipelining: at the call's end
    std::vector<uint32_t> ready_t(n_w, 0);
    auto call_start = [&](const Call& call, const Task& task) {
      uint64_t S = 0;
      for (uint32_t i = 0; i < task.n_in; i++) {
        if (task.in_slot[i] == 0xFFFF && task.seq_in_slot[i] == 0xFFFF) continue;
        const uint32_t w = prog.call_slots[call.in_off + i];
        if (w < first_wire) continue;
        uint64_t need = 0;
        if (opt.pipeline && i < task.in_need_level.size() && task.in_need_level[i] != 0xFFFFFFFFu)
          need = task.in_need_level[i] / task.window_levels * (uint64_t)task.window_levels;
        const uint64_t avail = ready_t[w] + (opt.pipeline ? 2 : 0);
        if (avail > need) S = std::max(S, avail - need);
      }
      return S;
    };
    auto publish = [&](const Call& call, const Task& task, uint64_t S) {
      for (uint32_t k = 0; k < task.n_out; k++) {
        uint64_t r = std::max<uint64_t>(1, opt.build_levelised ? task.n_levels : task.n_gates_total);
        if (opt.pipeline && k < task.out_ready_level.size())
          r = std::min<uint64_t>(r, (task.out_ready_level[k] / task.window_levels + 1) * (uint64_t)task.window_levels);
        ready_t[prog.call_slots[call.out_off + k]] = (uint32_t)std::min<uint64_t>(S + r, 0xFFFFFFFFu);
      }
    };
    auto weight = [&](uint32_t c) -> uint64_t {
      const Task& t = prog.tasks[prog.calls[c].task];
      return std::max<uint64_t>(1, opt.build_levelised ? t.n_levels : t.n_gates_total);
    };
    auto readers_finish = [&](uint32_t prev) {
      uint64_t f = finish[prev];
      for (uint32_t r : readers[prev]) f = std::max(f, finish[r]);
      return f;
    };
    for (uint32_t c = 0; c < n_calls; c++) {
      for (uint32_t pc : release_at[c]) waiting[block_size[pc]].push_back(FreeBlock{block_base[pc], pc, c});
      const Call& call = prog.calls[c];
      const Task& task = prog.tasks[call.task];
      uint64_t start = call_start(call, task);  // earliest start over the RAW edges
      // unique produced wires of this call, in output order
      std::vector<uint32_t> uniq;
      for (uint32_t k = 0; k < task.n_out; k++) {
        uint32_t w = prog.call_slots[call.out_off + k];
        if (slot_of[w] == UNSET) {
          slot_of[w] = 0;  // mark
          uniq.push_back(w);
        }
      }
      const uint32_t size = (uint32_t)uniq.size();
      if (size == 0) {
        finish[c] = start + weight(c);
        continue;
      }
      // blocks become eligible `reuse_distance` calls after their last reader
      {
        auto& q = waiting[size];
        size_t& head = waiting_head[size];
        auto& rd = ready[size];
        while (head < q.size() && q[head].free_at + opt.reuse_distance <= c) {
          rd.emplace(readers_finish(q[head].producer), q[head]);
          head++;
        }
      }
      uint32_t base;
      auto& rd = ready[size];
      auto it = rd.upper_bound(start);  // first block whose readers finish AFTER `start`
      bool take = false;
      if (it != rd.begin()) {           // the latest-finishing block that still costs nothing
        --it;
        take = true;
      } else if (!rd.empty() && (uint64_t)next_slot + size > opt.max_global_slots) {
        it = rd.begin();                // out of fresh slots: the block that delays this call least
        take = true;
      }
      if (take) {
        base = it->second.base;
        const uint32_t prev = it->second.producer;
        start = std::max(start, it->first);
        war[c] = readers[prev];
        if (war[c].empty()) war[c].push_back(prev);  // never read: at least wait for the old writer
        rd.erase(it);
      } else {
        base = next_slot;
        next_slot += size;
      }
      finish[c] = start + weight(c);
      publish(call, task, start);
      for (uint32_t k = 0; k < size; k++) slot_of[uniq[k]] = base + k;
      block_base[c] = base;
      block_size[c] = size;
      // high-fan-out blocks (e.g. the affine G1 points every line evaluation reads) are never recycled:
      // their reader list would become the next owner's dependency list
      if (!block_pinned[c] && readers[c].size() <= 32) release_at[std::min(n_calls, block_last[c] + 1)].push_back(c);
    }
    // rewrite wires -> slots, merge WAR dependencies
    for (uint32_t& w : prog.call_slots) w = slot_of[w] == UNSET ? 0u : slot_of[w];
    for (uint32_t& o : prog.output_slots)
      if (o != WIRE_DEAD) o = slot_of[o];
    std::vector<uint32_t> deps;
    deps.reserve(prog.deps.size() + n_calls);
    prog.max_call_deps = 0;
    for (uint32_t c = 0; c < n_calls; c++) {
      Call& call = prog.calls[c];
      std::vector<uint32_t> d(prog.deps.begin() + call.dep_off, prog.deps.begin() + call.dep_off + call.n_deps);
      std::vector<uint32_t> w = war[c];
      std::sort(w.begin(), w.end());
      w.erase(std::unique(w.begin(), w.end()), w.end());
      if (opt.pipeline) {
        // start dependencies (producers) first, then done dependencies (readers of the slots taken over); a call
        // that is both must have completed
        std::vector<uint32_t> st;
        std::set_difference(d.begin(), d.end(), w.begin(), w.end(), std::back_inserter(st));
        call.n_start_deps = (uint32_t)st.size();
        d = st;
        d.insert(d.end(), w.begin(), w.end());
      } else {
        d.insert(d.end(), w.begin(), w.end());
        std::sort(d.begin(), d.end());
        d.erase(std::unique(d.begin(), d.end()), d.end());
        call.n_start_deps = 0;
      }
      call.dep_off = (uint32_t)deps.size();
      call.n_deps = (uint32_t)d.size();
      deps.insert(deps.end(), d.begin(), d.end());
      prog.max_call_deps = std::max(prog.max_call_deps, call.n_deps);
    }
    prog.deps.swap(deps);
    prog.n_global_slots = next_slot;
  }
  prog.total_gates = p.gid;
  prog.total_ct = p.ct;
  prog.total_live = rt.total_live;
  for (int i = 0; i < 11; i++) prog.type_count[i] = rt.type_count[i];
  for (const Task& t : prog.tasks) {
    prog.max_task_slots = std::max(prog.max_task_slots, t.n_slots);
    prog.max_task_in = std::max(prog.max_task_in, t.n_in);
    prog.max_task_seq_slots = std::max(prog.max_task_seq_slots, t.n_seq_slots);
    prog.has_levelised = opt.build_levelised;
  }
  prog.pipelined = opt.pipeline && opt.build_levelised;
  if (prog.total_gates != rt.total_gates || prog.total_ct != rt.total_ct)
    throw std::logic_error("planner lost gates: " + std::to_string(prog.total_gates) + " vs " +
                           std::to_string(rt.total_gates));
  return prog;
}

}  // namespace gsv

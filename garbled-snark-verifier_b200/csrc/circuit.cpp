// circuit.cpp -- two-pass (credits, execution) circuit recorder.  See circuit.h.
#include "circuit.h"

#include <cstring>

#include <algorithm>
#include <cassert>
#include <stdexcept>

namespace gsv {

Builder::Builder() {}
Builder::~Builder() {}

bool Builder::in_meta() const { return !frames_.empty() && frames_.back().meta; }

Wire Builder::issue_wire() {
  Frame& f = frames_.back();
  if (f.meta) {
    // ComponentMetaBuilder::issue_wire, component_meta.rs:267-279
    f.credits.push_back(0);
    return WIRE_MIN + (Wire)f.credits.size() - 1;
  }
  // StreamingContext::issue_wire_with_credit, streaming_mode.rs:259-270
  if (f.cursor >= f.stack.size()) throw std::logic_error("credit stack exhausted in " + f.t->key);
  uint32_t credit = f.stack[f.cursor++];
  if (credit == 0) return WIRE_DEAD;  // Storage::allocate(credits = 0) -> UNREACHABLE
  return f.t->n_wires++;
}

Wires Builder::issue_wires(size_t n) {
  Wires w(n);
  for (size_t i = 0; i < n; i++) w[i] = issue_wire();
  return w;
}

void Builder::add_gate(uint8_t type, Wire a, Wire b, Wire c) {
  Frame& f = frames_.back();
  if (a == WIRE_DEAD || b == WIRE_DEAD) throw std::logic_error("gate reads an unreachable wire");
  if (f.meta) {
    // ComponentMetaBuilder::add_gate, component_meta.rs:281-296: one credit per read
    if (a >= WIRE_MIN) f.credits[a - WIRE_MIN]++;
    if (b >= WIRE_MIN) f.credits[b - WIRE_MIN]++;
    return;
  }
  Template& t = *f.t;
  t.items.push_back(Item{0, (uint32_t)t.gates.size()});
  t.gates.push_back(GateRec{a, b, c, type});
}

const Builder::CreditsTemplate& Builder::credits_for(const std::string& full_key, size_t n_in,
                                                     const Body& body) {
  auto it = credits_.find(full_key);
  if (it != credits_.end()) return it->second;
  // Child template construction, streaming_mode.rs:175-210: metadata pass over the body with
  // fresh mock inputs [MIN, MIN + n_in); grandchildren stay opaque.
  Frame f;
  f.meta = true;
  f.n_in = (uint32_t)n_in;
  f.credits.assign(n_in, 0);
  frames_.push_back(std::move(f));
  Wires ins(n_in);
  for (size_t i = 0; i < n_in; i++) ins[i] = WIRE_MIN + (Wire)i;
  Wires outs = body(*this, ins);
  Frame done = std::move(frames_.back());
  frames_.pop_back();
  CreditsTemplate ct;
  ct.internal.assign(done.credits.begin() + n_in, done.credits.end());
  ct.out_kind.reserve(outs.size());
  for (Wire o : outs) {
    // ComponentMetaBuilder::build, component_meta.rs:117-152
    if (o == WIRE_FALSE || o == WIRE_TRUE) ct.out_kind.push_back(-1);
    else if (o - WIRE_MIN < n_in) ct.out_kind.push_back(-2 - (int64_t)(o - WIRE_MIN));
    else ct.out_kind.push_back((int64_t)(o - WIRE_MIN - n_in));
  }
  return credits_.emplace(full_key, std::move(ct)).first->second;
}

void Builder::finalize_totals(Template& t) {
  t.total_gates = t.gates.size();
  for (const GateRec& g : t.gates) {
    t.type_count[g.type]++;
    if (g.c != WIRE_DEAD) {
      t.total_live++;
      if (!is_free(g.type)) t.total_ct++;
    }
  }
  for (const CallRec& c : t.calls) {
    const Template& ch = *templates_[c.tmpl];
    t.total_gates += ch.total_gates;
    t.total_ct += ch.total_ct;
    t.total_live += ch.total_live;
    for (int i = 0; i < 11; i++) t.type_count[i] += ch.type_count[i];
  }
}

uint32_t Builder::instantiate(const std::string& memo_key, size_t n_in,
                              const std::vector<uint32_t>& out_credits, const Body& body) {
  // full_key is memo_key up to '#'
  std::string full_key = memo_key.substr(0, memo_key.find('#'));
  const CreditsTemplate& ct = credits_for(full_key, n_in, body);
  if (ct.out_kind.size() != out_credits.size())
    throw std::logic_error("arity mismatch in " + full_key);
  // ComponentMetaTemplate::to_instance, component_meta.rs:179-222
  std::vector<uint32_t> stack = ct.internal;
  for (size_t j = 0; j < out_credits.size(); j++)
    if (ct.out_kind[j] >= 0) stack[(size_t)ct.out_kind[j]] += out_credits[j];

  auto t = std::make_unique<Template>();
  t->key = memo_key;
  t->n_in = (uint32_t)n_in;
  t->n_wires = WIRE_MIN + (uint32_t)n_in;
  Frame f;
  f.meta = false;
  f.t = t.get();
  f.stack = std::move(stack);
  frames_.push_back(std::move(f));
  Wires ins(n_in);
  for (size_t i = 0; i < n_in; i++) ins[i] = WIRE_MIN + (Wire)i;
  Wires outs = body(*this, ins);
  {
    Frame& top = frames_.back();
    if (top.cursor != top.stack.size())
      throw std::logic_error("credit stack not drained in " + memo_key);
  }
  frames_.pop_back();
  if (outs.size() != out_credits.size()) throw std::logic_error("body arity mismatch in " + memo_key);
  t->outs = std::move(outs);
  finalize_totals(*t);
  uint32_t idx = (uint32_t)templates_.size();
  templates_.push_back(std::move(t));
  memo_[memo_key] = idx;
  return idx;
}

static std::string liveness_mask(const std::vector<uint32_t>& out_credits) {
  // only zero / non-zero matters: a wire is unreachable iff own reads + parent credits == 0
  std::string m((out_credits.size() + 3) / 4, '0');
  for (size_t j = 0; j < out_credits.size(); j++)
    if (out_credits[j]) m[j / 4] = (char)(m[j / 4] + (1 << (j % 4)));  // '0'..'?' nibble chars
  return m;
}

Wires Builder::component(const std::string& key, const Wires& inputs, size_t arity, const Body& body) {
  if (frames_.empty()) throw std::logic_error("component() outside build_root");
  if (frames_.back().meta) {
    // ComponentMetaBuilder::with_named_child, component_meta.rs:284-301
    Frame& f = frames_.back();
    for (Wire w : inputs)
      if (w >= WIRE_MIN && w != WIRE_DEAD) f.credits[w - WIRE_MIN]++;
    Wires outs(arity);
    for (size_t i = 0; i < arity; i++) outs[i] = issue_wire();
    return outs;
  }
  // StreamingMode::with_named_child (execution pass), streaming_mode.rs:150-247
  std::vector<uint32_t> out_credits(arity);
  {
    Frame& f = frames_.back();
    if (f.cursor + arity > f.stack.size()) throw std::logic_error("credit stack exhausted at call " + key);
    for (size_t i = 0; i < arity; i++) out_credits[i] = f.stack[f.cursor++];
  }
  std::string full_key = key + "|a" + std::to_string(arity) + "|n" + std::to_string(inputs.size());
  std::string memo_key = full_key + "#" + liveness_mask(out_credits);
  uint32_t idx;
  auto it = memo_.find(memo_key);
  if (it != memo_.end()) idx = it->second;
  else idx = instantiate(memo_key, inputs.size(), out_credits, body);

  const Template& callee = *templates_[idx];
  Frame& f = frames_.back();
  Template& t = *f.t;
  CallRec c;
  c.tmpl = idx;
  c.in_off = (uint32_t)t.call_wires.size();
  t.call_wires.insert(t.call_wires.end(), inputs.begin(), inputs.end());
  c.out_off = (uint32_t)t.call_wires.size();
  Wires outs(arity);
  std::unordered_map<Wire, Wire> seen;  // callee-internal -> caller-local (duplicate outputs)
  for (size_t j = 0; j < arity; j++) {
    Wire o = callee.outs[j];
    if (o == WIRE_DEAD || o < WIRE_MIN) outs[j] = o;
    else if (o - WIRE_MIN < callee.n_in) outs[j] = inputs[o - WIRE_MIN];
    else {
      auto s = seen.find(o);
      if (s != seen.end()) outs[j] = s->second;
      else {
        outs[j] = t.n_wires++;
        seen.emplace(o, outs[j]);
      }
    }
  }
  t.call_wires.insert(t.call_wires.end(), outs.begin(), outs.end());
  t.items.push_back(Item{1, (uint32_t)t.calls.size()});
  t.calls.push_back(c);
  return outs;
}

uint32_t Builder::build_root(const std::string& name, size_t n_inputs, const Body& body) {
  if (!frames_.empty()) throw std::logic_error("build_root is not re-entrant");
  std::string full_key = "root::" + name + "|n" + std::to_string(n_inputs);
  // root metadata pass (mod.rs:260-265) -- also tells us the arity
  Frame guard;  // sentinel so credits_for can push/pop
  (void)guard;
  const CreditsTemplate& ct = credits_for(full_key, n_inputs, body);
  std::vector<uint32_t> out_credits(ct.out_kind.size(), 1);  // streaming_mode.rs:89-90
  std::string memo_key = full_key + "#root";
  return instantiate(memo_key, n_inputs, out_credits, body);
}

// ---- flat expansion ---------------------------------------------------------------------
namespace {
struct Flattener {
  const Builder& b;
  FlatStream& out;
  uint64_t max_gates;
  uint32_t next_id;

  Wires expand(uint32_t ti, const Wires& in_global) {
    const Template& t = b.tmpl(ti);
    constexpr uint32_t UNSET = 0xFFFFFFFEu;
    std::vector<uint32_t> l2g(t.n_wires, UNSET);
    l2g[0] = 0;
    l2g[1] = 1;
    for (uint32_t i = 0; i < t.n_in; i++) l2g[WIRE_MIN + i] = in_global[i];
    for (const Item& it : t.items) {
      if (!it.is_call) {
        const GateRec& g = t.gates[it.idx];
        if (out.type.size() >= max_gates) throw std::length_error("flat stream too large");
        uint32_t ga = l2g[g.a], gb = l2g[g.b];
        if (ga == UNSET || gb == UNSET) throw std::logic_error("read of unset wire in " + t.key);
        uint32_t gc = WIRE_DEAD;
        if (g.c != WIRE_DEAD) {
          gc = next_id++;
          l2g[g.c] = gc;  // SSA: an in-place overwrite gets a fresh id
        }
        out.type.push_back(g.type);
        out.a.push_back(ga);
        out.b.push_back(gb);
        out.c.push_back(gc);
      } else {
        const CallRec& c = t.calls[it.idx];
        const Template& ch = b.tmpl(c.tmpl);
        Wires ins(ch.n_in);
        for (uint32_t i = 0; i < ch.n_in; i++) {
          Wire w = t.call_wires[c.in_off + i];
          ins[i] = (w == WIRE_DEAD) ? WIRE_DEAD : l2g[w];
          if (ins[i] == UNSET) throw std::logic_error("call passes unset wire in " + t.key);
        }
        Wires outs = expand(c.tmpl, ins);
        for (size_t j = 0; j < outs.size(); j++) {
          Wire p = t.call_wires[c.out_off + j];
          if (p == WIRE_DEAD || p < WIRE_MIN) continue;
          if (l2g[p] == UNSET) l2g[p] = outs[j];
        }
      }
    }
    Wires res(t.outs.size());
    for (size_t j = 0; j < t.outs.size(); j++) {
      Wire o = t.outs[j];
      res[j] = (o == WIRE_DEAD) ? WIRE_DEAD : l2g[o];
    }
    return res;
  }
};
}  // namespace

namespace {
struct Executor {
  const Builder& b;
  std::vector<uint8_t> arena;  // per-instance wire values, stack allocated
  uint64_t gates = 0;
  static inline uint8_t eval(uint8_t t, uint8_t x, uint8_t y) {
    if (t < 8) return (uint8_t)((((x ^ (t >> 2)) & (y ^ (t >> 1))) ^ t) & 1);
    if (t == XOR) return x ^ y;
    if (t == XNOR) return (uint8_t)(x ^ y ^ 1);
    return (uint8_t)(x ^ 1);
  }
  // runs template `ti` whose frame starts at arena[base] (inputs already written at base+2..)
  void run(uint32_t ti, size_t base) {
    const Template& t = b.tmpl(ti);
    if (arena.size() < base + t.n_wires) arena.resize(std::max(arena.size() * 2, base + t.n_wires));
    arena[base] = 0;
    arena[base + 1] = 1;
    for (const Item& it : t.items) {
      if (!it.is_call) {
        const GateRec& g = t.gates[it.idx];
        gates++;
        if (g.c == WIRE_DEAD) continue;
        arena[base + g.c] = eval(g.type, arena[base + g.a], arena[base + g.b]);
      } else {
        const CallRec& c = t.calls[it.idx];
        const Template& ch = b.tmpl(c.tmpl);
        const size_t cb = base + t.n_wires;
        if (arena.size() < cb + ch.n_wires) arena.resize(std::max(arena.size() * 2, cb + ch.n_wires));
        for (uint32_t i = 0; i < ch.n_in; i++) {
          Wire w = t.call_wires[c.in_off + i];
          arena[cb + WIRE_MIN + i] = (w == WIRE_DEAD) ? 0 : arena[base + w];
        }
        run(c.tmpl, cb);
        for (size_t j = 0; j < ch.outs.size(); j++) {
          Wire p = t.call_wires[c.out_off + j];
          Wire o = ch.outs[j];
          if (p == WIRE_DEAD || p < WIRE_MIN || o == WIRE_DEAD) continue;
          if (o >= WIRE_MIN + ch.n_in) arena[base + p] = arena[cb + o];  // produced output (passthroughs alias)
        }
      }
    }
  }
};
}  // namespace

namespace {
// dependency depth of every wire (all gates / non-free gates only), same walk as Executor
struct DepthWalker {
  const Builder& b;
  std::vector<uint64_t> arena;  // hi 32 bits: depth counting every gate, lo 32: non-free gates only
  void run(uint32_t ti, size_t base) {
    const Template& t = b.tmpl(ti);
    if (arena.size() < base + t.n_wires) arena.resize(std::max(arena.size() * 2, base + t.n_wires));
    arena[base] = arena[base + 1] = 0;
    for (const Item& it : t.items) {
      if (!it.is_call) {
        const GateRec& g = t.gates[it.idx];
        if (g.c == WIRE_DEAD) continue;
        const uint64_t x = arena[base + g.a], y = arena[base + g.b];
        const uint64_t all = std::max(x >> 32, y >> 32) + 1;
        const uint64_t nf = std::max(x & 0xFFFFFFFFu, y & 0xFFFFFFFFu) + (is_free(g.type) ? 0 : 1);
        arena[base + g.c] = (all << 32) | nf;
      } else {
        const CallRec& c = t.calls[it.idx];
        const Template& ch = b.tmpl(c.tmpl);
        const size_t cb = base + t.n_wires;
        if (arena.size() < cb + ch.n_wires) arena.resize(std::max(arena.size() * 2, cb + ch.n_wires));
        for (uint32_t i = 0; i < ch.n_in; i++) {
          Wire w = t.call_wires[c.in_off + i];
          arena[cb + WIRE_MIN + i] = (w == WIRE_DEAD) ? 0 : arena[base + w];
        }
        run(c.tmpl, cb);
        for (size_t j = 0; j < ch.outs.size(); j++) {
          Wire p = t.call_wires[c.out_off + j];
          Wire o = ch.outs[j];
          if (p == WIRE_DEAD || p < WIRE_MIN || o == WIRE_DEAD) continue;
          if (o >= WIRE_MIN + ch.n_in) arena[base + p] = arena[cb + o];
        }
      }
    }
  }
};
}  // namespace

void circuit_depth(const Builder& b, uint32_t root, uint64_t* depth_all, uint64_t* depth_nonfree) {
  const Template& t = b.tmpl(root);
  DepthWalker w{b, std::vector<uint64_t>(std::max<size_t>(1 << 20, t.n_wires), 0)};
  w.run(root, 0);
  uint64_t a = 0, n = 0;
  for (Wire o : t.outs)
    if (o != WIRE_DEAD) {
      a = std::max(a, w.arena[o] >> 32);
      n = std::max<uint64_t>(n, w.arena[o] & 0xFFFFFFFFu);
    }
  if (depth_all) *depth_all = a;
  if (depth_nonfree) *depth_nonfree = n;
}

std::vector<uint8_t> execute(const Builder& b, uint32_t root, const std::vector<uint8_t>& input_bits,
                             uint64_t* gates_executed) {
  const Template& t = b.tmpl(root);
  if (input_bits.size() != t.n_in) throw std::invalid_argument("execute: wrong number of input bits");
  Executor ex{b, std::vector<uint8_t>(1 << 20), 0};
  if (ex.arena.size() < t.n_wires) ex.arena.resize(t.n_wires);
  for (uint32_t i = 0; i < t.n_in; i++) ex.arena[WIRE_MIN + i] = input_bits[i] ? 1 : 0;
  ex.run(root, 0);
  std::vector<uint8_t> out(t.outs.size());
  for (size_t j = 0; j < t.outs.size(); j++) out[j] = (t.outs[j] == WIRE_DEAD) ? 0 : ex.arena[t.outs[j]];
  if (gates_executed) *gates_executed = ex.gates;
  return out;
}

FlatStream flatten(const Builder& b, uint32_t root, uint64_t max_gates) {
  FlatStream fs;
  const Template& t = b.tmpl(root);
  if (t.total_gates > max_gates) throw std::length_error("flat stream too large");
  fs.n_inputs = t.n_in;
  fs.type.reserve(t.total_gates);
  fs.a.reserve(t.total_gates);
  fs.b.reserve(t.total_gates);
  fs.c.reserve(t.total_gates);
  Flattener f{b, fs, max_gates, WIRE_MIN + t.n_in};
  Wires ins(t.n_in);
  for (uint32_t i = 0; i < t.n_in; i++) ins[i] = WIRE_MIN + i;
  fs.outputs = f.expand(root, ins);
  fs.n_wires = f.next_id;
  return fs;
}

// The memoised template DAG as flat arrays (layout documented at gsv_program_export_templates).
bool export_templates(const Builder& b, uint64_t sizes[6], uint32_t* tmpl, uint32_t* gates, uint32_t* calls,
                      uint32_t* items, uint32_t* call_wires, uint32_t* outs) {
  uint64_t ng = 0, nc = 0, ni = 0, nw = 0, no = 0;
  for (size_t t = 0; t < b.n_templates(); t++) {
    const Template& T = b.tmpl((uint32_t)t);
    if (tmpl) {
      uint32_t* r = tmpl + 12 * t;
      r[0] = T.n_in; r[1] = T.n_wires; r[2] = (uint32_t)ng; r[3] = (uint32_t)T.gates.size();
      r[4] = (uint32_t)nc; r[5] = (uint32_t)T.calls.size(); r[6] = (uint32_t)ni; r[7] = (uint32_t)T.items.size();
      r[8] = (uint32_t)nw; r[9] = (uint32_t)T.call_wires.size(); r[10] = (uint32_t)no; r[11] = (uint32_t)T.outs.size();
    }
    if (gates)
      for (size_t k = 0; k < T.gates.size(); k++) {
        uint32_t* g = gates + 4 * (ng + k);
        g[0] = T.gates[k].a; g[1] = T.gates[k].b; g[2] = T.gates[k].c; g[3] = T.gates[k].type;
      }
    if (calls)
      for (size_t k = 0; k < T.calls.size(); k++) {
        uint32_t* c = calls + 3 * (nc + k);
        c[0] = T.calls[k].tmpl; c[1] = T.calls[k].in_off; c[2] = T.calls[k].out_off;
      }
    if (items)
      for (size_t k = 0; k < T.items.size(); k++) items[ni + k] = (T.items[k].is_call ? 0x80000000u : 0u) | T.items[k].idx;
    if (call_wires && !T.call_wires.empty()) memcpy(call_wires + nw, T.call_wires.data(), T.call_wires.size() * 4);
    if (outs && !T.outs.empty()) memcpy(outs + no, T.outs.data(), T.outs.size() * 4);
    ng += T.gates.size(); nc += T.calls.size(); ni += T.items.size(); nw += T.call_wires.size(); no += T.outs.size();
  }
  if (ng >= (1ull << 32) || nw >= (1ull << 32) || ni >= (1ull << 32)) return false;
  sizes[0] = 12 * b.n_templates(); sizes[1] = 4 * ng; sizes[2] = 3 * nc; sizes[3] = ni; sizes[4] = nw; sizes[5] = no;
  return true;
}

}  // namespace gsv

// circuit.h -- host-side circuit recorder: the C++ mirror of the reference's streaming builder.
//
// The reference never materialises the circuit: gadget closures run twice (metadata pass ->
// per-wire fan-out "credits", then the execution pass) and hand gates one at a time to a
// CircuitMode (src/circuit/mod.rs:253-301, src/circuit/streaming_mode.rs:134-247,
// src/circuit/component_meta.rs).  The B200 engine needs the topology up front, once, so this
// recorder runs the same two passes but *records* instead of evaluating:
//
//   * `Builder::component()` == `CircuitContext::with_named_child` (circuit_context_trait.rs:12-27):
//     pops the output credits, instantiates the credits template (built by a metadata pass over
//     the body, children opaque: component_meta.rs:284-301), and runs the body.
//   * `Builder::issue_wire()` == `StreamingContext::issue_wire_with_credit`: a wire whose credit
//     is 0 becomes WIRE_DEAD (Storage::allocate -> UNREACHABLE, src/storage.rs:119-133), so the
//     gate that writes it consumes a gate index but produces no label and no ciphertext
//     (garble_mode.rs:192-197).
//   * bodies are memoised per (component key, output-liveness mask) into `Template`s with
//     template-local wire ids, so a 11 G-gate verifier is a small DAG of templates, not a flat
//     145 GB gate list.  Gate order inside a template is the emission order of the reference.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace gsv {

using Wire = uint32_t;
using Wires = std::vector<Wire>;
constexpr Wire WIRE_FALSE = 0;  // FALSE_WIRE, circuit_context_trait.rs:2
constexpr Wire WIRE_TRUE = 1;   // TRUE_WIRE
constexpr Wire WIRE_MIN = 2;    // WireId::MIN, src/core/wire.rs:7
constexpr Wire WIRE_DEAD = 0xFFFFFFFFu;  // WireId::UNREACHABLE

// src/core/gate_type.rs:1-15
enum GateType : uint8_t { AND = 0, NAND, NIMP, IMP, NCIMP, CIMP, NOR, OR, XOR, XNOR, NOT };
inline bool is_free(uint8_t t) { return t >= XOR; }

struct GateRec {
  Wire a, b, c;
  uint8_t type;
};

struct CallRec {
  uint32_t tmpl;     // callee template index
  uint32_t in_off;   // into Template::call_wires: n_in callee-input bindings (caller-local ids)
  uint32_t out_off;  // into Template::call_wires: n_out caller-local ids (or consts / DEAD)
};

struct Item {
  uint32_t is_call : 1;
  uint32_t idx : 31;
};

// One memoised component body in template-local numbering:
// ids 0/1 = constants, [2, 2+n_in) = inputs by position, then internals in issue order.
struct Template {
  std::string key;
  uint32_t n_in = 0;
  uint32_t n_wires = 2;  // exclusive upper bound of local ids
  std::vector<GateRec> gates;
  std::vector<CallRec> calls;
  std::vector<Item> items;  // emission order
  Wires call_wires;
  Wires outs;  // local ids; may be inputs, constants or WIRE_DEAD
  // totals over the flattened body
  uint64_t total_gates = 0;    // every add_gate, dead ones included (== gate-index advance)
  uint64_t total_ct = 0;       // live non-free gates (== ciphertexts)
  uint64_t total_live = 0;     // live gates
  uint64_t total_internal = 0; // internal wires over the flattened body (SSA id budget)
  uint64_t type_count[11] = {0};
};

class Builder;
using Body = std::function<Wires(Builder&, const Wires&)>;

class Builder {
 public:
  Builder();
  ~Builder();

  // CircuitContext::issue_wire
  Wire issue_wire();
  Wires issue_wires(size_t n);
  // CircuitContext::add_gate
  void add_gate(uint8_t type, Wire a, Wire b, Wire c);
  // CircuitContext::with_named_child.  `key` plays the role of generate_component_key's
  // (name, params, arity, input_len) tuple (component_key.rs:15-39); arity and input length are
  // appended here.  `memo_extra` disambiguates bodies that the reference keys identically but
  // whose gate lists differ (never the case inside one circuit; kept for safety).
  Wires component(const std::string& key, const Wires& inputs, size_t arity, const Body& body);

  // CircuitBuilder::run_streaming: root metadata pass, then the execution pass; every root
  // output gets one credit (streaming_mode.rs:89-90).  Returns the root template index.
  uint32_t build_root(const std::string& name, size_t n_inputs, const Body& body);

  const Template& tmpl(uint32_t i) const { return *templates_[i]; }
  size_t n_templates() const { return templates_.size(); }
  bool in_meta() const;

 private:
  struct CreditsTemplate {
    std::vector<uint32_t> internal;  // credits of internal wires in issue order
    std::vector<int64_t> out_kind;   // >=0: internal index, -1: constant, <=-2: input (-2-pos)
  };
  struct Frame {
    bool meta;
    // meta pass
    std::vector<uint32_t> credits;  // index = id - WIRE_MIN (inputs first)
    uint32_t n_in = 0;
    // execution pass
    Template* t = nullptr;
    std::vector<uint32_t> stack;  // credits in issue order
    size_t cursor = 0;
  };
  const CreditsTemplate& credits_for(const std::string& full_key, size_t n_in, const Body& body);
  uint32_t instantiate(const std::string& full_key, size_t n_in, const std::vector<uint32_t>& out_credits,
                       const Body& body);
  void finalize_totals(Template& t);

  std::vector<Frame> frames_;
  std::vector<std::unique_ptr<Template>> templates_;
  std::unordered_map<std::string, CreditsTemplate> credits_;
  std::unordered_map<std::string, uint32_t> memo_;
};

// ---- flat expansion (feeds the oracle and small-circuit tests) ------------------------------
struct FlatStream {
  std::vector<uint8_t> type;
  std::vector<uint32_t> a, b, c;  // SSA ids: 0,1 consts, 2.. inputs, then issue order; c may be DEAD
  Wires outputs;
  uint32_t n_inputs = 0;
  uint32_t n_wires = 0;
};
// Expands template `root` depth-first in emission order.  Throws std::length_error if the
// stream would exceed `max_gates`.
FlatStream flatten(const Builder& b, uint32_t root, uint64_t max_gates = (1ull << 31));

// ExecuteMode (src/circuit/modes/execute_mode.rs): plain boolean evaluation of the recorded circuit,
// walking the template DAG without materialising the flat stream.  Host-side topology self-check
// (the reference uses the same mode for its gadget tests and pre-checks); NOT a garbling path.
std::vector<uint8_t> execute(const Builder& b, uint32_t root, const std::vector<uint8_t>& input_bits,
                             uint64_t* gates_executed = nullptr);


// The memoised template DAG as flat arrays (layout: gsv_program_export_templates in gsv_cuda.h).  Pass null
// arrays to query the six sizes (in words).  false when an array would exceed 2^32 words.
bool export_templates(const Builder& b, uint64_t sizes[6], uint32_t* tmpl, uint32_t* gates, uint32_t* calls,
                      uint32_t* items, uint32_t* call_wires, uint32_t* outs);

// Gate-level dependency depth of the circuit outputs: counting every live gate, and counting only the
// non-free (AND-family, one dependent hash each) gates -- the floor of any schedule's critical path.
void circuit_depth(const Builder& b, uint32_t root, uint64_t* depth_all, uint64_t* depth_nonfree);

}  // namespace gsv

// circuitgen.cpp -- host-only workload generator for the CPU oracle (TEST / BASELINE INFRASTRUCTURE).
//
// The reference's circuits exist only as Rust gadget closures (src/gadgets/**); the CPU oracle needs the
// same gate streams without loading the CUDA engine (bench.py --impl reference must run none of the
// product's kernels or runtime).  This library links ONLY the host-side gadget restatement and recorder
// (csrc/circuit.cpp, gadgets*.cpp, bn254_host.cpp, circuits.cpp: no CUDA, no planner, no scheduler) and
// exports the recorded template DAG in the layout gsvo_garble_templates walks.
#include <cstdint>
#include <memory>
#include <string>

#include "circuit.h"
#include "gadgets.h"

namespace {
struct Gen {
  gsv::Builder b;
  uint32_t root = 0;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* gsvgen_last_error(void) { return g_err.c_str(); }

void* gsvgen_build(const char* circuit) {
  try {
    auto g = std::make_unique<Gen>();
    g->root = gsv::build_named_circuit(g->b, circuit ? circuit : "");
    return g.release();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void gsvgen_destroy(void* h) { delete static_cast<Gen*>(h); }

// sizes[6] in words; pass null arrays to query (see gsv_program_export_templates)
int gsvgen_export(void* h, uint64_t sizes[6], uint32_t* root, uint32_t* tmpl, uint32_t* gates, uint32_t* calls,
                  uint32_t* items, uint32_t* call_wires, uint32_t* outs) {
  Gen* g = static_cast<Gen*>(h);
  if (!g || !sizes) return -1;
  if (!gsv::export_templates(g->b, sizes, tmpl, gates, calls, items, call_wires, outs)) return -1;
  if (root) *root = g->root;
  return 0;
}
// per template: gates / ciphertexts over its flattened body (for choosing bounded samples)
int gsvgen_totals(void* h, uint64_t* total_gates, uint64_t* total_ct) {
  Gen* g = static_cast<Gen*>(h);
  if (!g) return -1;
  for (size_t t = 0; t < g->b.n_templates(); t++) {
    if (total_gates) total_gates[t] = g->b.tmpl((uint32_t)t).total_gates;
    if (total_ct) total_ct[t] = g->b.tmpl((uint32_t)t).total_ct;
  }
  return 0;
}

}  // extern "C"

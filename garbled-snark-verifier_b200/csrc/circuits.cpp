// circuits.cpp -- the named workload circuits (gsv_program_build and the host-only generator library).
#include <stdexcept>
#include <string>

#include "gadgets.h"

namespace gsv {

uint32_t build_named_circuit(Builder& b, const std::string& c) {
  if (c == "fq12_mul") return build_fq12_mul(b);
  if (c == "fq6_mul") return build_fq6_mul(b);
  if (c == "fq2_mul") return build_fq2_mul(b);
  if (c == "fq_mul") return build_fq_mul(b);
  if (c == "fq_add") return build_fq_add(b);
  if (c == "fq_expr") return build_fq_expr(b);
  if (c == "gate_zoo") return build_gate_zoo(b);
  if (c.rfind("bn_mul", 0) == 0) return build_bn_mul(b, (size_t)std::stoul(c.substr(6)));
  if (c == "fq_inverse") return build_fq_inverse(b);
  if (c == "fq_sqrt") return build_fq_sqrt(b);
  if (c == "fq2_sqrt") return build_fq2_sqrt(b);
  if (c == "g1_add") return build_g1_add(b);
  if (c == "g1_msm1")
    return build_g1_msm1(b, host::g1_to_affine(host::g1_mul(host::g1_from_affine(host::g1_generator()), U256(0xC0FFEE))));
  if (c == "fq12_square") return build_fq12_square(b);
  if (c == "fq12_cyclotomic_square") return build_fq12_cyclotomic_square(b);
  if (c == "fq12_inverse") return build_fq12_inverse(b);
  if (c.rfind("fq12_frobenius", 0) == 0) return build_fq12_frobenius(b, (size_t)std::stoul(c.substr(14)));
  if (c == "g2_double_step") return build_g2_double_step(b);
  if (c == "g2_add_step") return build_g2_add_step(b);
  if (c == "g2_mul_by_char") return build_g2_mul_by_char(b);
  if (c == "ell") return build_ell(b);
  if (c == "ell_const") return build_ell_const(b);
  if (c == "g1_to_affine") return build_g1_to_affine(b);
  if (c == "decompress_g1") return build_decompress_g1(b);
  if (c == "final_exponentiation") return build_final_exponentiation(b);
  if (c == "miller_loop_groth16" || c == "groth16_verify_compressed" || c == "groth16_verify") {
    host::VerifyingKey vk;
    host::Proof pr;
    host::synthetic_groth16(7, U256(424242), vk, pr);
    if (c == "miller_loop_groth16") return build_miller_loop_groth16(b, host::g2_neg(vk.gamma_g2), host::g2_neg(vk.delta_g2));
    if (c == "groth16_verify") return build_groth16_verify(b, vk, 1);
    return build_groth16_verify_compressed(b, vk, 1);
  }
  throw std::invalid_argument("unknown circuit: " + c);
}

}  // namespace gsv

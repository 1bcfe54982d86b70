"""Root functions (ctx, input wires) -> output wires of the named circuits, over the emission_model gadgets."""
import emission_model as em

N = em.N


def _miller(x, w):
    sc = em.synthetic_vk_scalars(7)
    neg = lambda q: (q[0], em._f2neg(q[1]))
    k1, k2 = neg(em._g2_mul(em.G2_GENERATOR, sc["gamma"])), neg(em._g2_mul(em.G2_GENERATOR, sc["delta"]))
    return em.miller_loop_groth16(x, w[:3 * N], w[3 * N:6 * N], w[6 * N:9 * N], k1, k2, w[9 * N:])


ROOTS = {
    "fq_add": (2 * N, lambda x, w: em.fq_add(x, w[:N], w[N:])),
    "fq_mul": (2 * N, lambda x, w: em.fq_mul(x, w[:N], w[N:])),
    "fq2_mul": (4 * N, lambda x, w: sum(em.fq2_mul(x, [w[:N], w[N:2 * N]], [w[2 * N:3 * N], w[3 * N:]]), [])),
    "fq6_mul": (12 * N, lambda x, w: em._flat6(em.fq6_mul(x, em._fq6_of(w[:6 * N]), em._fq6_of(w[6 * N:])))),
    "fq12_mul": (24 * N, lambda x, w: em.fq12_mul(x, w[:12 * N], w[12 * N:])),
    "fq_inverse": (N, lambda x, w: em.fq_inverse_montgomery(x, w)),
    "g1_add": (6 * N, lambda x, w: em.g1_add(x, w[:3 * N], w[3 * N:])),
    "fq12_square": (12 * N, lambda x, w: em.fq12_square(x, w)),
    "fq12_cyclotomic_square": (12 * N, lambda x, w: em.fq12_cyclotomic_square(x, w)),
    "fq12_inverse": (12 * N, lambda x, w: em.fq12_inverse(x, w)),
    "fq12_frobenius1": (12 * N, lambda x, w: em.fq12_frobenius(x, w, 1)),
    "g2_double_step": (6 * N, lambda x, w: em.g2_double_step(x, w)),
    "g2_add_step": (12 * N, lambda x, w: em.g2_add_step(x, w[:6 * N], w[6 * N:])),
    "ell": (21 * N, lambda x, w: em.ell(x, w[:12 * N], w[12 * N:18 * N], w[18 * N:])),
    "ell_const": (15 * N, lambda x, w: em.ell_by_constant(x, w[:12 * N], ((3, 5), (7, 11), (13, 17)), w[12 * N:])),
    "g1_to_affine": (3 * N, lambda x, w: em.g1_to_affine(x, w)),
    "fq_sqrt": (N, lambda x, w: em.fq_sqrt(x, w)),
    "decompress_g1": (N + 1, lambda x, w: em.decompress_g1(x, w[:N], w[N])),
    "final_exponentiation": (12 * N, lambda x, w: em.final_exponentiation(x, w)),
    "miller_loop_groth16": (15 * N, _miller),
    "fq2_sqrt": (2 * N, lambda x, w: sum(em.fq2_sqrt_general(x, [w[:N], w[N:]]), [])),
    "g1_msm1": (N, lambda x, w: em.g1_msm_const(x, [w], [em._g1_mul_affine(0xC0FFEE) + (1,)])),
    "groth16_verify_compressed": (N + 2 * (N + 1) + 2 * N + 1, lambda x, w: em.groth16_verify_compressed(x, w, 1, em.synthetic_vk(7))),
    "groth16_verify": (9 * N, lambda x, w: em.groth16_verify_uncompressed(x, w, 1, em.synthetic_vk(7))),
}

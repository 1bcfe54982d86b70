//! Prints, from the REAL crate, the quantities the engine's CPU oracle is pinned on (SURVEY.md Appendix E and
//! tests/golden/garble_vectors.json of the gsv-b200 repository).  The engine's image has no Rust toolchain, so
//! two rows could only be pinned against independent restatements there: the seed -> delta / label derivation
//! (rand_chacha 0.3.1 / rand 0.8.5 / rand_core 0.6.4) and the gate stream of Fq12::mul_montgomery.  One
//! `cargo test --release golden_dump -- --nocapture > dump.jsonl` wherever the reference builds settles both:
//! `python tests/golden/check_reference_dump.py dump.jsonl` compares the lines with the committed fixtures.
use garbled_snark_verifier::{
    AESAccumulatingHash, AesNiHasher, Blake3Hasher, GarbleMode, GarbledWire, GateHasher, S, WireId,
    circuit::{CiphertextHandler, CircuitBuilder, CircuitInput, CircuitMode, EncodeInput, StreamingResult, WiresObject},
    gadgets::bn254::fq12::Fq12,
};
use rand::{Rng, SeedableRng};
use rand_chacha::ChaCha20Rng;

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{x:02x}")).collect()
}

/// Two Fq12 operands as 6 096 input wires (a then b, `to_wires_vec` order: tests/fq12_mul_e2e.rs:41-52).
#[derive(Clone)]
struct TwoFq12;
#[derive(Clone)]
struct TwoFq12Wires {
    a: Fq12,
    b: Fq12,
}
impl CircuitInput for TwoFq12 {
    type WireRepr = TwoFq12Wires;
    fn allocate(&self, mut issue: impl FnMut() -> WireId) -> Self::WireRepr {
        TwoFq12Wires { a: Fq12::new(&mut issue), b: Fq12::new(issue) }
    }
    fn collect_wire_ids(repr: &Self::WireRepr) -> Vec<WireId> {
        let mut ids = repr.a.to_wires_vec();
        ids.extend(repr.b.to_wires_vec());
        ids
    }
}
impl<H: GateHasher, CTH: CiphertextHandler> EncodeInput<GarbleMode<H, CTH>> for TwoFq12 {
    fn encode(&self, repr: &Self::WireRepr, cache: &mut GarbleMode<H, CTH>) {
        // one label draw per input wire, in wire order (the label values do not depend on the plaintext)
        for w in Self::collect_wire_ids(repr) {
            let gw = cache.issue_garbled_wire();
            cache.feed_wire(w, gw);
        }
    }
}

fn dump_fq12<H: GateHasher>(name: &str, seed: u64) {
    let res: StreamingResult<GarbleMode<H, AESAccumulatingHash>, _, Vec<GarbledWire>> =
        CircuitBuilder::streaming_garbling(TwoFq12, 20_000, seed, AESAccumulatingHash::default(), |ctx, w: &TwoFq12Wires| {
            Fq12::mul_montgomery(ctx, &w.a, &w.b).to_wires_vec()
        });
    println!(
        "{{\"kind\":\"garble\",\"circuit\":\"fq12_mul\",\"hasher\":\"{name}\",\"seed\":{seed},\"ct_commit\":\"{}\",\
         \"false_label0\":\"{}\",\"true_label0\":\"{}\",\"first_input_label0\":\"{}\",\"first_output_label0\":\"{}\",\"n_gates\":{}}}",
        hex(&res.ciphertext_handler_result),
        hex(&res.false_wire_constant.select(false).to_bytes()),
        hex(&res.true_wire_constant.select(false).to_bytes()),
        hex(&res.input_wire_values[0].select(false).to_bytes()),
        hex(&res.output_value[0].select(false).to_bytes()),
        res.gate_count.total_gate_count()
    );
}

#[test]
fn golden_dump() {
    // row a2: ChaCha20Rng::seed_from_u64 -> first u128 draws (delta is the first, garble_mode.rs:80-97)
    for seed in [0u64, 42, 1234] {
        let mut rng = ChaCha20Rng::seed_from_u64(seed);
        let draws: Vec<String> = (0..4).map(|_| hex(&S::from_u128(rng.r#gen::<u128>()).to_bytes())).collect();
        println!("{{\"kind\":\"rng_u128\",\"seed\":{seed},\"draws\":{draws:?}}}");
    }
    // cut-and-choose instance seeds (garbler.rs:201-203)
    let mut rng = ChaCha20Rng::seed_from_u64(1234);
    let seeds: Vec<u64> = (0..16).map(|_| rng.r#gen()).collect();
    println!("{{\"kind\":\"instance_seeds\",\"master\":1234,\"seeds\":{seeds:?}}}");
    // rows a3 / a4: the gate hashers
    let x = S::from_u128(0x0123456789abcdeffedcba9876543210);
    let [h] = AesNiHasher::default().hash_with_gate(&[x], (1usize << 32) + 5);
    println!("{{\"kind\":\"hash\",\"hasher\":\"aes\",\"x\":\"0123456789abcdeffedcba9876543210\",\"gid\":4294967301,\"value\":\"{}\"}}", hex(&h.to_bytes()));
    let [h] = Blake3Hasher::default().hash_with_gate(&[x], (1usize << 32) + 5);
    println!("{{\"kind\":\"hash\",\"hasher\":\"blake3\",\"x\":\"0123456789abcdeffedcba9876543210\",\"gid\":4294967301,\"value\":\"{}\"}}", hex(&h.to_bytes()));
    // rows a5 / a7 / a8 / a9: Fq12::mul_montgomery garbled with the chain commitment, both hashers
    for seed in [0u64, 42, 1234] {
        dump_fq12::<AesNiHasher>("aes", seed);
        dump_fq12::<Blake3Hasher>("blake3", seed);
    }
}

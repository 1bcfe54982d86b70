//! Records key-independent circuits with the REAL reference gadgets through the engine's C ABI (`GpuRecorder` is a
//! `CircuitContext`) and writes the recorded template DAGs to `target/structure/<circuit>.gsvt`.
//! `python tests/golden/check_reference_structure.py gsv-cuda/target/structure` (in the gsv-b200 repository) hashes
//! them and compares with `tests/golden/structural_hashes.json` -- the hashes of an independent Python restatement
//! that the engine's own generator already equals (tests/test_structural_hash.py), up to the whole verifier.
//! Equal hashes mean: the reference's gadgets emit exactly the gate stream (order, types, wiring, dead gates) the
//! engine was built and measured on.  Needs no GPU: recording and planning are host code.
//! The whole verifier can be added the same way: `structural_hashes.json` lists the scalars of the synthetic key
//! (`synthetic_key.scalars`: every key element is scalar * generator) behind its `groth16_verify_compressed` /
//! `groth16_verify` / `miller_loop_groth16` hashes.
use std::{fs, io::Write, path::PathBuf};

use garbled_snark_verifier::{
    WireId,
    circuit::{FromWires, WiresObject},
    gadgets::{
        bigint::BigIntWires,
        bn254::{
            final_exponentiation::final_exponentiation_montgomery, fq::Fq, fq12::Fq12, g1::G1Projective,
        },
    },
};
use gsv_cuda::{GpuRecorder, Program};

fn dump(name: &str, n_inputs: usize, n_outputs: usize, root: impl Fn(&mut GpuRecorder, &[WireId]) -> Vec<WireId>) {
    let prog = Program::record(name, n_inputs, n_outputs, root).expect("record");
    let (root_idx, arrays) = prog.export_templates().expect("export");
    let dir = PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("target").join("structure");
    fs::create_dir_all(&dir).unwrap();
    let mut f = fs::File::create(dir.join(format!("{name}.gsvt"))).unwrap();
    f.write_all(b"GSVT").unwrap();
    f.write_all(&root_idx.to_le_bytes()).unwrap();
    for a in &arrays {
        f.write_all(&(a.len() as u64).to_le_bytes()).unwrap();
    }
    for a in &arrays {
        for w in a {
            f.write_all(&w.to_le_bytes()).unwrap();
        }
    }
    println!("{name}: {} gates recorded, {} templates", prog.info.n_gates, arrays[0].len() / 12);
}

fn fq_of(w: &[WireId]) -> Fq {
    Fq(BigIntWires::from_wires(w).unwrap())
}

#[test]
fn structure_dump() {
    const N: usize = 254;
    // Fq::mul_montgomery, Fq::inverse_montgomery (the roots of the engine's "fq_mul" / "fq_inverse")
    dump("fq_mul", 2 * N, N, |c, w| Fq::mul_montgomery(c, &fq_of(&w[..N]), &fq_of(&w[N..])).to_wires_vec());
    dump("fq_inverse", N, N, |c, w| Fq::inverse_montgomery(c, &fq_of(w)).to_wires_vec());
    // Fq12 multiplication, squaring, inverse (tests/fq12_mul_e2e.rs input order: a then b, to_wires_vec order)
    dump("fq12_mul", 24 * N, 12 * N, |c, w| {
        let (a, b) = (Fq12::from_wires(&w[..12 * N]).unwrap(), Fq12::from_wires(&w[12 * N..]).unwrap());
        Fq12::mul_montgomery(c, &a, &b).to_wires_vec()
    });
    dump("fq12_square", 12 * N, 12 * N, |c, w| Fq12::square_montgomery(c, &Fq12::from_wires(w).unwrap()).to_wires_vec());
    dump("fq12_inverse", 12 * N, 12 * N, |c, w| Fq12::inverse_montgomery(c, &Fq12::from_wires(w).unwrap()).to_wires_vec());
    // projective G1 addition
    dump("g1_add", 6 * N, 3 * N, |c, w| {
        let (p, q) = (G1Projective::from_wires(&w[..3 * N]).unwrap(), G1Projective::from_wires(&w[3 * N..]).unwrap());
        G1Projective::add_montgomery(c, &p, &q).to_wires_vec()
    });
    // the final exponentiation: 3.5 G gates, a few MB as a template DAG
    dump("final_exponentiation", 12 * N, 12 * N, |c, w| {
        final_exponentiation_montgomery(c, &Fq12::from_wires(w).unwrap()).to_wires_vec()
    });
}

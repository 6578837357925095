//! Cross-check harness: runs halo2_proofs 0.3.0 itself on the Rust circuit of
//! `zkodst_backend::halo2_chip` and compares everything that reaches the transcript with libzkodst.so.
//!
//!   cargo run --release -p zkodst-xcheck -- [k] [n_compressions] [out.json]
//!
//! Steps (the reference's sequence, blake2f-circuit/benches/blake2f.rs:83-142, with the seed of
//! benchmarking/src/blake2f_circuit_bench.rs:41-44):
//!   1. `Params::<EqAffine>::new(k)`, written with `Params::write` and loaded into the library with
//!      `zk_params_load` — both sides use halo2's genuine URS;
//!   2. `MockProver::run(k, &circuit, vec![]).verify()` on the cells the library computed;
//!   3. `keygen_vk`: `format!("{:?}", vk.pinned())` vs `zk_vk_pinned_debug` (first differing offset is
//!      reported), fixed / permutation commitments vs `zk_vk_bytes`;
//!   4. `create_proof` with `XorShiftRng::from_seed(SEED)` vs `zk_create_proof` with the same seed: proof bytes;
//!   5. `verify_proof` of the library's proof under halo2's vk, and `zk_verify_proof` of halo2's proof.
//! Writes the JSON fixture tests/test_xcheck_fixture.py consumes (copy it to tests/golden/xcheck.json).
//!
//! NOT COMPILED in the build image (no Rust toolchain, no registry access).
use std::fs::File;
use std::io::Write;

use halo2_proofs::{
    dev::MockProver,
    pasta::{EqAffine, Fp},
    plonk::{create_proof, keygen_pk, keygen_vk, verify_proof, SingleVerifier},
    poly::commitment::Params,
    transcript::{Blake2bRead, Blake2bWrite, Challenge255},
};
use rand_core::{RngCore, SeedableRng};
use rand_xorshift::XorShiftRng;

use zkodst_backend::halo2_chip::Blake2fCircuit;
use zkodst_backend::{Blake2fWitness, GpuProver};

/// benchmarking/src/blake2f_circuit_bench.rs:41-44
const SEED: [u8; 16] = [0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5];

/// The synthetic records of bench.py / zk-odst_b200/inputs.py: h, m, t = consecutive XorShiftRng outputs,
/// f = i & 1, rounds = 12.
fn synthetic_inputs(n: usize) -> Vec<Blake2fWitness> {
    let mut rng = XorShiftRng::from_seed(SEED);
    (0..n)
        .map(|i| {
            let mut w = Blake2fWitness { rounds: 12, h: [0; 8], m: [0; 16], t: [0; 2], f: i & 1 == 1 };
            for x in w.h.iter_mut() { *x = rng.next_u64(); }
            for x in w.m.iter_mut() { *x = rng.next_u64(); }
            for x in w.t.iter_mut() { *x = rng.next_u64(); }
            w
        })
        .collect()
}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let k: u32 = args.get(1).map(|s| s.parse().unwrap()).unwrap_or(17);
    let n: usize = args.get(2).map(|s| s.parse().unwrap()).unwrap_or(2);
    let out = args.get(3).cloned().unwrap_or_else(|| "xcheck.json".into());
    let inputs = synthetic_inputs(n);

    // 1. params: halo2's own URS on both sides
    let params: Params<EqAffine> = Params::new(k);
    let mut params_bytes = Vec::new();
    params.write(&mut params_bytes).unwrap();
    let mut gpu = GpuProver::new(0).expect("no CUDA device");
    gpu.load_params(&params_bytes).expect("zk_params_load");
    gpu.keygen(12, n as u64).expect("zk_blake2f_keygen");

    // 2. MockProver on the library's cells
    let circuit = Blake2fCircuit::new(&mut gpu, k, 12, &inputs, vec![false; n]).expect("witness");
    let mock = MockProver::<Fp>::run(k, &circuit, vec![]).unwrap();
    let mock_ok = mock.verify().is_ok();

    // 3. keys
    let vk = keygen_vk(&params, &circuit).unwrap();
    let pinned_halo2 = format!("{:?}", vk.pinned());
    let pinned_lib = gpu.vk_pinned_debug().unwrap();
    let first_diff = pinned_halo2.bytes().zip(pinned_lib.bytes()).position(|(a, b)| a != b)
        .or(if pinned_halo2.len() != pinned_lib.len() { Some(pinned_halo2.len().min(pinned_lib.len())) } else { None });
    let pk = keygen_pk(&params, vk, &circuit).unwrap();

    // 4. proofs with the same seeded RNG
    let mut transcript = Blake2bWrite::<_, _, Challenge255<_>>::init(vec![]);
    create_proof(&params, &pk, &[circuit.clone()], &[&[]], XorShiftRng::from_seed(SEED), &mut transcript).unwrap();
    let proof_halo2: Vec<u8> = transcript.finalize();
    let proof_lib = gpu.create_proof(&inputs, SEED).unwrap();

    // 5. cross verification
    let strategy = SingleVerifier::new(&params);
    let mut reader = Blake2bRead::<_, _, Challenge255<_>>::init(&proof_lib[..]);
    let halo2_accepts_lib = verify_proof(&params, pk.get_vk(), strategy, &[&[]], &mut reader).is_ok();
    let lib_accepts_halo2 = gpu.verify_proof(&proof_halo2).unwrap_or(false);

    let vk_bytes_lib = gpu.vk_bytes().unwrap();
    let mut f = File::create(&out).unwrap();
    write!(f, "{{\n \"k\": {}, \"n_compressions\": {}, \"seed\": \"{}\",\n \"mock_prover_ok\": {},\n \
               \"pinned_equal\": {}, \"pinned_first_diff\": {},\n \"pinned_halo2_len\": {}, \"pinned_lib_len\": {},\n \
               \"vk_bytes_lib\": \"{}\",\n \"proof_equal\": {},\n \"proof_halo2\": \"{}\",\n \"proof_lib\": \"{}\",\n \
               \"halo2_accepts_lib\": {}, \"lib_accepts_halo2\": {}\n}}\n",
           k, n, hex(&SEED), mock_ok, first_diff.is_none(),
           first_diff.map(|d| d.to_string()).unwrap_or_else(|| "null".into()),
           pinned_halo2.len(), pinned_lib.len(), hex(&vk_bytes_lib), proof_halo2 == proof_lib,
           hex(&proof_halo2), hex(&proof_lib), halo2_accepts_lib, lib_accepts_halo2).unwrap();
    std::fs::write(format!("{}.pinned_halo2.txt", out), &pinned_halo2).unwrap();
    std::fs::write(format!("{}.pinned_lib.txt", out), &pinned_lib).unwrap();
    println!("mock {} pinned_equal {} proof_equal {} halo2_accepts_lib {} lib_accepts_halo2 {} -> {}",
             mock_ok, first_diff.is_none(), proof_halo2 == proof_lib, halo2_accepts_lib, lib_accepts_halo2, out);
}

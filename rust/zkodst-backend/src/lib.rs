//! Drop-in façade for the proving path of zk-odst's BLAKE2f circuit.
//!
//! Keeps the reference's plugin surface (paths relative to the reference repository):
//!   * `Table16Chip::{configure, construct, load}`      blake2f-circuit/src/blake2f/table16.rs:267-336
//!   * `Blake2fInstructions::{initialization_vector, initialization, compress, digest}`
//!                                                      blake2f-circuit/src/blake2f.rs:40-72
//!   * `Circuit::synthesize`                            blake2f-circuit/src/blake2f.rs:270-277
//!   * `create_proof` call shape                        blake2f-circuit/benches/blake2f.rs:124-127
//! and routes the heavy lifting through the C ABI.  `configure` stays pure Rust (it only declares
//! columns, gates and the lookup exactly as docs/CIRCUIT.md lists them) so that halo2's own
//! `keygen_vk`, `MockProver` and `verify_proof` keep working on the same circuit.
//!
//! NOTE: this crate is shipped as source; the build image has no Rust toolchain, so it has not
//! been compiled.  INTEGRATION.md walks through the steps a maintainer follows.
use std::ffi::CStr;
use std::ptr;

use zkodst_sys as sys;

/// EIP-152 precompile input, the `Blake2fWitness` of the reference's commented test circuit
/// (blake2f-circuit/src/blake2f.rs:201-240).
#[derive(Clone, Debug)]
pub struct Blake2fWitness {
    pub rounds: u32,
    pub h: [u64; 8],
    pub m: [u64; 16],
    pub t: [u64; 2],
    pub f: bool,
}

impl Blake2fWitness {
    pub fn to_eip152(&self) -> [u8; sys::ZK_BLAKE2F_INPUT_BYTES] {
        let mut out = [0u8; sys::ZK_BLAKE2F_INPUT_BYTES];
        out[0..4].copy_from_slice(&self.rounds.to_be_bytes());
        for (i, w) in self.h.iter().enumerate() {
            out[4 + 8 * i..12 + 8 * i].copy_from_slice(&w.to_le_bytes());
        }
        for (i, w) in self.m.iter().enumerate() {
            out[68 + 8 * i..76 + 8 * i].copy_from_slice(&w.to_le_bytes());
        }
        for (i, w) in self.t.iter().enumerate() {
            out[196 + 8 * i..204 + 8 * i].copy_from_slice(&w.to_le_bytes());
        }
        out[212] = self.f as u8;
        out
    }
}

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

/// One context per (host thread, device); not `Sync`.
pub struct GpuProver {
    ctx: *mut sys::zk_ctx,
}

impl GpuProver {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::zk_ctx_create(device, &mut ctx) };
        if rc != sys::ZK_OK {
            return Err(Error { code: rc, message: "zk_ctx_create failed (no CPU fallback)".into() });
        }
        Ok(Self { ctx })
    }

    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == sys::ZK_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::zk_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }

    /// `Params::read` (benches/blake2f.rs:83-97): bytes produced by halo2's `Params::write`.
    pub fn load_params(&mut self, bytes: &[u8]) -> Result<(), Error> {
        self.check(unsafe { sys::zk_params_load(self.ctx, bytes.as_ptr(), bytes.len() as u64) })
    }

    /// `keygen_vk` + `keygen_pk` (benches/blake2f.rs:102-103) for `n` compressions of `rounds`.
    pub fn keygen(&mut self, rounds: u32, n: u64) -> Result<(), Error> {
        self.check(unsafe { sys::zk_blake2f_keygen(self.ctx, rounds, n) })
    }

    /// `create_proof(.., rng, &mut transcript); transcript.finalize()` (benches/blake2f.rs:124-127).
    /// `seed` seeds the XorShiftRng the prover consumes
    /// (benchmarking/src/blake2f_circuit_bench.rs:41-44).
    pub fn create_proof(&mut self, inputs: &[Blake2fWitness], seed: [u8; 16]) -> Result<Vec<u8>, Error> {
        let mut flat = Vec::with_capacity(inputs.len() * sys::ZK_BLAKE2F_INPUT_BYTES);
        for w in inputs {
            flat.extend_from_slice(&w.to_eip152());
        }
        let mut proof = vec![0u8; 1 << 16];
        let mut len = proof.len() as u64;
        self.check(unsafe {
            sys::zk_create_proof(self.ctx, flat.as_ptr(), inputs.len() as u64, seed.as_ptr(),
                                 proof.as_mut_ptr(), &mut len)
        })?;
        proof.truncate(len as usize);
        Ok(proof)
    }

    /// `verify_proof(&params, vk, SingleVerifier::new(&params), &[&[]], &mut Blake2bRead::init(proof))`
    /// (benches/blake2f.rs:138-144) against this prover's params and keys.
    pub fn verify_proof(&mut self, proof: &[u8]) -> Result<bool, Error> {
        match unsafe { sys::zk_verify_proof(self.ctx, proof.as_ptr(), proof.len() as u64) } {
            sys::ZK_OK => Ok(true),
            sys::ZK_E_VERIFY => Ok(false),
            rc => self.check(rc).map(|_| false),
        }
    }

    /// `MockProver::run(k, &circuit, vec![]).verify()` (table16/spread_table.rs:759-763):
    /// `Ok(None)` when every constraint holds, else the first failure (kind, row, index).
    pub fn mock_verify(&mut self, inputs: &[Blake2fWitness]) -> Result<Option<(u64, u64, u64)>, Error> {
        let mut flat = Vec::with_capacity(inputs.len() * sys::ZK_BLAKE2F_INPUT_BYTES);
        for w in inputs {
            flat.extend_from_slice(&w.to_eip152());
        }
        let mut fail = [0u64; 3];
        match unsafe {
            sys::zk_mock_verify(self.ctx, flat.as_ptr(), inputs.len() as u64, ptr::null(), fail.as_mut_ptr())
        } {
            sys::ZK_OK => Ok(None),
            sys::ZK_E_VERIFY => Ok(Some((fail[0], fail[1], fail[2]))),
            rc => self.check(rc).map(|_| None),
        }
    }

    /// Joins a group of `world` processes (one per GPU) that split every MSM by point range
    /// (BASELINE configs[3]).  `id` comes from `zk_dist_unique_id` on rank 0; call before
    /// `load_params`.  All ranks must then prove the same inputs with the same seed.
    pub fn join_group(&mut self, id: &[u8; sys::ZK_DIST_ID_BYTES], rank: i32, world: i32) -> Result<(), Error> {
        self.check(unsafe { sys::zk_dist_init(self.ctx, id.as_ptr(), rank, world) })
    }

    /// `Circuit::synthesize` for the GPU path: fills all advice columns of the batch.
    /// `advice` must hold 12 * 2^k field elements (column-major, Montgomery form — the memory
    /// image of halo2's `Polynomial<Fp, LagrangeCoeff>` values).
    pub fn witness(&mut self, k: u32, rounds: u32, inputs: &[Blake2fWitness],
                   advice: &mut [[u64; 4]]) -> Result<(), Error> {
        assert_eq!(advice.len(), 12usize << k);
        let mut flat = Vec::with_capacity(inputs.len() * sys::ZK_BLAKE2F_INPUT_BYTES);
        for w in inputs {
            flat.extend_from_slice(&w.to_eip152());
        }
        self.check(unsafe {
            sys::zk_blake2f_witness_batch(self.ctx, k as i32, rounds, flat.as_ptr(), inputs.len() as u64,
                                          advice.as_mut_ptr() as *mut _, ptr::null_mut())
        })
    }
}

impl Drop for GpuProver {
    fn drop(&mut self) {
        unsafe { sys::zk_ctx_destroy(self.ctx) }
    }
}

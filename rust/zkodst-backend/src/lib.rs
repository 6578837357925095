//! Drop-in façade for the proving path of zk-odst's BLAKE2f circuit.
//!
//! Keeps the reference's plugin surface (paths relative to the reference repository):
//!   * `Table16Chip::{configure, construct, load}`      blake2f-circuit/src/blake2f/table16.rs:267-336
//!   * `Blake2fInstructions::{initialization_vector, initialization, compress, digest}`
//!                                                      blake2f-circuit/src/blake2f.rs:40-72
//!   * `Circuit::synthesize`                            blake2f-circuit/src/blake2f.rs:270-277
//!   * `create_proof` call shape                        blake2f-circuit/benches/blake2f.rs:124-127
//! and routes the heavy lifting through the C ABI.  Two layers:
//!   * `halo2_chip` — `Table16Chip::configure(meta: &mut ConstraintSystem<pallas::Base>)` in pure Rust (the 12
//!     advice columns, 3 table columns, the constants column, the selector-less lookup, 14 selectors and 26 gate
//!     polynomials of docs/CIRCUIT.md in declaration order) and a `Circuit` whose `synthesize` assigns the cells
//!     `zk_blake2f_witness_batch` computes, so that halo2's own `keygen_vk`, `MockProver` and `verify_proof` run
//!     on the same circuit (rust/xcheck drives them and compares with the library);
//!   * `gadget` — the streaming `Blake2fInstructions` / `Blake2f` surface over a recording layouter, the
//!     counterpart of include/zkodst.hpp.
//!
//! NOTE: this crate is shipped as source; the build image has no Rust toolchain, so it has not
//! been compiled.  INTEGRATION.md walks through the steps a maintainer follows.  The same surface
//! exists in C++ (include/zkodst.hpp), which IS compiled and run by the test suite
//! (tests/cpp/facade_test.cpp); `gadget` below follows it line for line.
use std::ffi::CStr;
use std::ptr;

/// `Table16Chip::configure` over halo2's own `ConstraintSystem` and the batch `Circuit` (see the module docs).
pub mod halo2_chip;

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

    /// The same with record chaining: `chain[j]` makes compression j continue compression j - 1
    /// (`CompressionConfig::initialize_with_state`, table16/compression.rs:1096-1111).
    pub fn keygen_chained(&mut self, rounds: u32, chain: &[bool]) -> Result<(), Error> {
        let flags: Vec<u8> = chain.iter().map(|&c| c as u8).collect();
        self.check(unsafe { sys::zk_blake2f_keygen_chained(self.ctx, rounds, flags.len() as u64, flags.as_ptr()) })
    }

    /// The `{:?}` rendering of `vk.pinned()` the library hashed into `vk.transcript_repr`.
    pub fn vk_pinned_debug(&mut self) -> Result<String, Error> {
        let mut len = 0u64;
        unsafe { sys::zk_vk_pinned_debug(self.ctx, ptr::null_mut(), &mut len) };
        let mut buf = vec![0u8; len as usize];
        self.check(unsafe { sys::zk_vk_pinned_debug(self.ctx, buf.as_mut_ptr() as *mut _, &mut len) })?;
        Ok(String::from_utf8_lossy(&buf).into_owned())
    }

    /// 12 fixed + 8 permutation commitments (32 B compressed each), then `vk.transcript_repr`.
    pub fn vk_bytes(&mut self) -> Result<Vec<u8>, Error> {
        let mut len = 0u64;
        unsafe { sys::zk_vk_bytes(self.ctx, ptr::null_mut(), &mut len) };
        let mut buf = vec![0u8; len as usize];
        self.check(unsafe { sys::zk_vk_bytes(self.ctx, buf.as_mut_ptr(), &mut len) })?;
        Ok(buf)
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

// ---- the chip / gadget surface ----------------------------------------------------------------------------
// Same shape as include/zkodst.hpp (which is compiled and tested in this repository): `synthesize` runs
// against a recording layouter — every `compress` lays out one region, i.e. one EIP-152 record — and the
// cells of all regions are assigned on the device in one call.
pub mod gadget {
    use super::{sys, Blake2fWitness, Error};

    /// The size of a message block, in 64-bit words (blake2f-circuit/src/blake2f.rs:34).
    pub const BLOCK_SIZE: usize = 16;
    /// The size of a digest, in 64-bit words (src/blake2f.rs:36).
    pub const DIGEST_SIZE: usize = 8;
    /// table16.rs:47-56
    pub const IV: [u64; 8] = [
        0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
        0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179,
    ];

    /// table16.rs:59-61 with 64-bit words; `None` = `Value::unknown()` (key generation).
    #[derive(Clone, Copy, Debug)]
    pub struct BlockWord(pub Option<u64>);
    impl Default for BlockWord {
        fn default() -> Self {
            BlockWord(Some(0)) // the zero word `finalize` pads with
        }
    }

    /// table16/compression.rs:286-525 `State`, as values, plus what BLAKE2b's F needs besides the chaining
    /// value: the byte counter of the block about to be compressed, the final flag, the round count.
    #[derive(Clone, Debug)]
    pub struct State {
        pub h: [Option<u64>; 8],
        pub t: u64,
        pub last: bool,
        pub rounds: u32,
    }

    /// Records the regions a circuit lays out: 213 bytes per compression, in layout order.
    pub struct Layouter {
        pub witnesses_known: bool,
        pub table_loaded: bool,
        pub rounds: u32,
        pub regions: u64,
        pub records: Vec<u8>,
    }
    impl Layouter {
        pub fn new(witnesses_known: bool) -> Self {
            Layouter { witnesses_known, table_loaded: false, rounds: 0, regions: 0, records: Vec::new() }
        }
    }

    /// blake2f-circuit/src/blake2f.rs:40-72
    pub trait Blake2fInstructions {
        fn initialization_vector(&self, layouter: &mut Layouter) -> Result<State, Error>;
        fn initialization(&self, layouter: &mut Layouter, init_state: &State) -> Result<State, Error>;
        fn compress(&self, layouter: &mut Layouter, initialized_state: &State,
                    input: [BlockWord; BLOCK_SIZE]) -> Result<State, Error>;
        fn digest(&self, layouter: &mut Layouter, state: &State) -> Result<[BlockWord; DIGEST_SIZE], Error>;
    }

    #[derive(Clone, Debug, Default)]
    pub struct Table16Config;

    /// blake2f-circuit/src/blake2f/table16.rs:250-384
    #[derive(Clone, Debug)]
    pub struct Table16Chip {
        config: Table16Config,
    }
    impl Table16Chip {
        /// The column / gate / lookup plan is fixed (docs/CIRCUIT.md) and lives in the library.
        pub fn configure() -> Table16Config {
            Table16Config
        }
        pub fn construct(config: Table16Config) -> Self {
            Table16Chip { config }
        }
        pub fn load(_config: Table16Config, layouter: &mut Layouter) -> Result<(), Error> {
            layouter.table_loaded = true;
            Ok(())
        }
        pub fn config(&self) -> &Table16Config {
            &self.config
        }
    }

    fn synthesis_error(message: &str) -> Error {
        Error { code: sys::ZK_E_INPUT, message: message.into() }
    }

    impl Blake2fInstructions for Table16Chip {
        fn initialization_vector(&self, _layouter: &mut Layouter) -> Result<State, Error> {
            let mut h = [None; 8];
            for i in 0..8 {
                h[i] = Some(IV[i]);
            }
            h[0] = Some(IV[0] ^ 0x0101_0040); // BLAKE2b-512, unkeyed
            Ok(State { h, t: 0, last: false, rounds: 12 })
        }
        fn initialization(&self, _layouter: &mut Layouter, init_state: &State) -> Result<State, Error> {
            Ok(init_state.clone())
        }
        fn compress(&self, layouter: &mut Layouter, st: &State, input: [BlockWord; BLOCK_SIZE]) -> Result<State, Error> {
            if !layouter.table_loaded {
                return Err(Error { code: sys::ZK_E_STATE, message: "compress before load".into() });
            }
            if layouter.regions > 0 && layouter.rounds != st.rounds {
                return Err(synthesis_error("regions of one circuit must share the round count"));
            }
            layouter.rounds = st.rounds;
            layouter.regions += 1;
            let mut out = st.clone();
            if !layouter.witnesses_known {
                out.h = [None; 8];
                return Ok(out);
            }
            let mut w = Blake2fWitness { rounds: st.rounds, h: [0; 8], m: [0; 16], t: [st.t, 0], f: st.last };
            for i in 0..8 {
                w.h[i] = st.h[i].ok_or_else(|| synthesis_error("unknown chaining value in a proving run"))?;
            }
            for i in 0..BLOCK_SIZE {
                w.m[i] = input[i].0.ok_or_else(|| synthesis_error("unknown block word in a proving run"))?;
            }
            let record = w.to_eip152();
            layouter.records.extend_from_slice(&record);
            let mut hout = [0u8; 64];
            let rc = unsafe { sys::zk_blake2f_compress(record.as_ptr(), hout.as_mut_ptr()) };
            if rc != sys::ZK_OK {
                return Err(Error { code: rc, message: "zk_blake2f_compress".into() });
            }
            for i in 0..8 {
                let mut b = [0u8; 8];
                b.copy_from_slice(&hout[8 * i..8 * i + 8]);
                out.h[i] = Some(u64::from_le_bytes(b));
            }
            Ok(out)
        }
        fn digest(&self, _layouter: &mut Layouter, state: &State) -> Result<[BlockWord; DIGEST_SIZE], Error> {
            let mut out = [BlockWord(None); DIGEST_SIZE];
            for i in 0..DIGEST_SIZE {
                out[i] = BlockWord(state.h[i]);
            }
            Ok(out)
        }
    }

    /// blake2f-circuit/src/blake2f.rs:80-181, at a granularity of one 64-bit word.  BLAKE2b flags its last
    /// block, so a full block is compressed only once more data (or `finalize`) arrives.
    pub struct Blake2f<CS: Blake2fInstructions> {
        chip: CS,
        state: State,
        cur_block: Vec<BlockWord>,
        bytes_done: u64,
    }
    impl<CS: Blake2fInstructions> Blake2f<CS> {
        pub fn new(chip: CS, layouter: &mut Layouter) -> Result<Self, Error> {
            let state = chip.initialization_vector(layouter)?;
            Ok(Blake2f { chip, state, cur_block: Vec::with_capacity(BLOCK_SIZE), bytes_done: 0 })
        }
        fn flush(&mut self, layouter: &mut Layouter, last: bool, block_bytes: u64) -> Result<(), Error> {
            let mut st = self.chip.initialization(layouter, &self.state)?;
            st.t = self.bytes_done + block_bytes;
            st.last = last;
            let mut block = [BlockWord::default(); BLOCK_SIZE];
            block[..self.cur_block.len()].copy_from_slice(&self.cur_block);
            self.state = self.chip.compress(layouter, &st, block)?;
            self.bytes_done += block_bytes;
            self.cur_block.clear();
            Ok(())
        }
        pub fn update(&mut self, layouter: &mut Layouter, data: &[BlockWord]) -> Result<(), Error> {
            for w in data {
                if self.cur_block.len() == BLOCK_SIZE {
                    self.flush(layouter, false, 128)?;
                }
                self.cur_block.push(*w);
            }
            Ok(())
        }
        /// `unused_bytes_of_last_word` (< 8) trims the byte counter when the message does not end on a word boundary.
        pub fn finalize(mut self, layouter: &mut Layouter, unused_bytes_of_last_word: u64)
                        -> Result<[BlockWord; DIGEST_SIZE], Error> {
            let words = self.cur_block.len() as u64;
            let bytes = words * 8 - if words > 0 { unused_bytes_of_last_word } else { 0 };
            self.flush(layouter, true, bytes)?;
            self.chip.digest(layouter, &self.state)
        }
        pub fn digest(chip: CS, layouter: &mut Layouter, data: &[BlockWord], unused_bytes_of_last_word: u64)
                      -> Result<[BlockWord; DIGEST_SIZE], Error> {
            let mut hasher = Self::new(chip, layouter)?;
            hasher.update(layouter, data)?;
            hasher.finalize(layouter, unused_bytes_of_last_word)
        }
    }
}

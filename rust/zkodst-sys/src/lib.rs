//! `extern "C"` declarations mirroring include/zkodst.h one to one.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct zk_ctx {
    _private: [u8; 0],
}

pub const ZK_OK: i32 = 0;
pub const ZK_E_INVALID: i32 = -1;
pub const ZK_E_CUDA: i32 = -2;
pub const ZK_E_NOMEM: i32 = -3;
pub const ZK_E_ROWS: i32 = -4;
pub const ZK_E_INPUT: i32 = -5;
pub const ZK_E_STATE: i32 = -6;
pub const ZK_E_VERIFY: i32 = -7;
pub const ZK_E_BUFFER: i32 = -8;
pub const ZK_BLAKE2F_INPUT_BYTES: usize = 213;
pub const ZK_NUM_SELECTORS: usize = 14;
pub const ZK_NUM_ADVICE: usize = 12;

extern "C" {
    pub fn zk_ctx_create(device_id: i32, out: *mut *mut zk_ctx) -> i32;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);
    pub fn zk_last_error(ctx: *const zk_ctx) -> *const c_char;
    pub fn zk_ctx_set_stream(ctx: *mut zk_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn zk_ctx_synchronize(ctx: *mut zk_ctx) -> i32;
    pub fn zk_ctx_set_blocking_sync(ctx: *mut zk_ctx, on: i32) -> i32;
    pub fn zk_ctx_launch_count(ctx: *const zk_ctx) -> u64;
    pub fn zk_blake2f_rows_per_compression(rounds: u32, rows: *mut u64) -> i32;
    pub fn zk_blake2f_min_k(rounds: u32, n_compressions: u64, k: *mut i32) -> i32;
    pub fn zk_blake2f_layout_tables(rounds: u32, copies: *mut u32, n_copies: *mut u64, selectors: *mut u8,
        constants: *mut u64, chain_rows: *mut u32) -> i32;
    pub fn zk_blake2f_witness_batch(ctx: *mut zk_ctx, k: i32, rounds: u32, inputs: *const u8,
        n_compressions: u64, advice_out: *mut c_void, digests_out: *mut u64) -> i32;
    pub fn zk_msm_vesta(ctx: *mut zk_ctx, scalars: *const c_void, bases: *const c_void, n: u64,
        on_device: i32, out_affine: *mut c_void) -> i32;
    pub fn zk_ntt_fp(ctx: *mut zk_ctx, data: *mut c_void, log_n: i32, inverse: i32, on_device: i32) -> i32;
    pub fn zk_commit_batch(ctx: *mut zk_ctx, basis: i32, scalars: *const c_void, ncols: u32, blinds: *const c_void,
        index_mask: u32, index_select: i32, on_device: i32, out_affine: *mut c_void) -> i32;
    pub fn zk_ntt_fp_batch(ctx: *mut zk_ctx, input: *const c_void, out: *mut c_void, log_n: i32, batch: u32,
        inverse: i32, on_device: i32) -> i32;
    pub fn zk_coeff_to_cosets(ctx: *mut zk_ctx, coeffs: *const c_void, ncols: u32, on_device: i32,
        out: *mut c_void) -> i32;
    pub fn zk_params_generate_substitute(ctx: *mut zk_ctx, k: i32, seed: *const u8) -> i32;
    pub fn zk_params_load(ctx: *mut zk_ctx, bytes: *const u8, len: u64) -> i32;
    pub fn zk_params_write(ctx: *mut zk_ctx, out: *mut u8, len: *mut u64) -> i32;
    pub fn zk_blake2f_keygen(ctx: *mut zk_ctx, rounds: u32, n_compressions: u64) -> i32;
    pub fn zk_blake2f_keygen_chained(ctx: *mut zk_ctx, rounds: u32, n_compressions: u64, chain: *const u8) -> i32;
    pub fn zk_vk_bytes(ctx: *mut zk_ctx, out: *mut u8, len: *mut u64) -> i32;
    pub fn zk_vk_pinned_debug(ctx: *mut zk_ctx, out: *mut c_char, len: *mut u64) -> i32;
    pub fn zk_blake2f_pinned_debug(k: i32, rounds: u32, commitments: *const c_void, out: *mut c_char, len: *mut u64) -> i32;
    pub fn zk_vk_repr_override(ctx: *mut zk_ctx, repr: *const u8) -> i32;
    pub fn zk_create_proof(ctx: *mut zk_ctx, inputs: *const u8, n_compressions: u64,
        seed: *const u8, proof_out: *mut u8, proof_len: *mut u64) -> i32;
    pub fn zk_create_proof_device_inputs(ctx: *mut zk_ctx, d_inputs: *const u8, n_compressions: u64,
        seed: *const u8, proof_out: *mut u8, proof_len: *mut u64) -> i32;
    pub fn zk_eip152_validate(input: *const u8, len: u64, rounds: *mut u32) -> i32;
    pub fn zk_blake2f_compress(input: *const u8, out: *mut u8) -> i32;
    pub fn zk_blake2b_records(msg: *const u8, len: u64, rounds: u32, records_out: *mut u8,
        n_records: *mut u64, digest_out: *mut u8) -> i32;
    pub fn zk_verify_proof(ctx: *mut zk_ctx, proof: *const u8, proof_len: u64) -> i32;
    pub fn zk_verify_proofs_batch(ctx: *mut zk_ctx, proofs: *const u8, proof_lens: *const u64, count: u64,
        seed: *const u8) -> i32;
    pub fn zk_mock_verify(ctx: *mut zk_ctx, inputs: *const u8, n_compressions: u64,
        advice_override: *const c_void, failure: *mut u64) -> i32;
    pub fn zk_dist_unique_id(out: *mut u8) -> i32;
    pub fn zk_dist_init(ctx: *mut zk_ctx, id: *const u8, rank: i32, world: i32) -> i32;
    pub fn zk_dist_info(ctx: *const zk_ctx, rank: *mut i32, world: *mut i32) -> i32;
    pub fn zk_dist_range(n_points: u64, rank: i32, world: i32, lo: *mut u64, hi: *mut u64) -> i32;
    pub fn zk_dist_column_block(rank: i32, world: i32, lo: *mut u32, hi: *mut u32, per_rank: *mut u32) -> i32;
    pub fn zk_dist_quotient_rows(n: u64, rank: i32, world: i32, row_lo: *mut u64, row_hi: *mut u64,
        segments: *mut u64, n_segments: *mut u32) -> i32;
}
pub const ZK_DIST_ID_BYTES: usize = 128;

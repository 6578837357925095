/*
 * zkodst.h — C ABI of the B200-native prover backend for zk-odst's BLAKE2f Table16 circuit.
 *
 * This is the drop-in boundary: the entry points a Rust `extern "C"` block
 * (rust/zkodst-sys) binds in place of the CPU work the reference reaches
 * through halo2_proofs 0.3.0.  Plain pointers and sizes only; every call returns an
 * int32_t status (0 = ZK_OK, negative = ZK_E_*); nothing unwinds across the boundary.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * ZK_E_CUDA.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef ZKODST_H
#define ZKODST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZK_OK 0
#define ZK_E_INVALID (-1)   /* bad argument (NULL pointer, k out of range, malformed input) */
#define ZK_E_CUDA (-2)      /* CUDA runtime / driver error, or no device */
#define ZK_E_NOMEM (-3)     /* host or device allocation failed */
#define ZK_E_ROWS (-4)      /* circuit does not fit 2^k rows (halo2 Error::NotEnoughRowsAvailable) */
#define ZK_E_INPUT (-5)     /* EIP-152 input rejected (f not in {0,1}, rounds mismatch) */
#define ZK_E_STATE (-6)     /* call order violated (e.g. prove before params/keygen) */
#define ZK_E_VERIFY (-7)    /* proof rejected / constraint not satisfied */
#define ZK_E_BUFFER (-8)    /* caller buffer too small; required size written back */

#define ZK_BLAKE2F_INPUT_BYTES 213 /* EIP-152: rounds(4 BE) h(64) m(128) t(16) f(1) */
#define ZK_FIELD_BYTES 32          /* Fp / Fq element: 4 x u64 LE limbs, Montgomery form */
#define ZK_POINT_BYTES 64          /* Vesta affine point: x, y (Montgomery Fq) */
#define ZK_NUM_ADVICE 12           /* table16.rs:281-310 + spread_table.rs:435-441 */

typedef struct zk_ctx zk_ctx;

/* ---- context ------------------------------------------------------------------------- */

/* One context per (host thread, device).  Owns all device memory it allocates and a private
 * CUDA stream.  Replaces nothing in the reference (which has no device); it is the handle
 * the Rust shim keeps inside its `Table16Chip` backend state. */
int32_t zk_ctx_create(int32_t device_id, zk_ctx** out);
void zk_ctx_destroy(zk_ctx* ctx);
/* Last error text for this context (valid until the next call on the context). */
const char* zk_last_error(const zk_ctx* ctx);
/* Run the context's work on an externally owned CUDA stream (cudaStream_t as void*);
 * NULL restores the private stream. */
int32_t zk_ctx_set_stream(zk_ctx* ctx, void* cuda_stream);
int32_t zk_ctx_synchronize(zk_ctx* ctx);
/* How the host thread waits for the device inside the library's calls: 0 = spin (lowest latency,
 * the default), 1 = sleep on a blocking event — for several contexts per CPU core (e.g. 8 GPUs x 4
 * proof streams).  Environment ZK_BLOCKING_SYNC=1 sets the default for new contexts. */
int32_t zk_ctx_set_blocking_sync(zk_ctx* ctx, int32_t on);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t zk_ctx_launch_count(const zk_ctx* ctx);
/* Milliseconds the most recent launch of the named kernel class took, measured with CUDA
 * events on the context's stream.  which: 0 = witness kernel. */
int32_t zk_ctx_last_kernel_ms(zk_ctx* ctx, int32_t which, float* ms);
int32_t zk_ctx_enable_timing(zk_ctx* ctx, int32_t on);
/* Device time (ms) and number of timed regions per kernel class since the previous report:
 * 0 witness, 1 msm (whole pipeline), 2 ntt, 3 quotient, 4 ipa generator collapse,
 * 5 msm bucket accumulation; arrays of 8. */
int32_t zk_ctx_timing_report(zk_ctx* ctx, float* ms_per_class, uint32_t* launches_per_class);

/* Integer-pipe micro-benchmark (the compute roofline of the field-arithmetic kernels;
 * SURVEY.md §6 asks for it because no INT32 peak was measured by the driver).
 * mode 0 = mad.lo.u32, 1 = mad.wide.u32, 2 = mad.lo.cc/madc.hi.cc chain, 9 = fma.rn.f64, 10 = fma.rn.f64 and
 * mad.wide.u32 interleaved one to one (do the two pipes issue side by side?): issued instructions per second
 * over the whole device.  3 = Fq Montgomery products per second, 4 = XYZZ mixed additions per second (5: with
 * the product as a real call; 6, 7, 8: at 5, 6, 8 resident blocks per SM). */
int32_t zk_bench_int_pipe(zk_ctx* ctx, int32_t mode, uint32_t iters, double* instr_per_sec);

/* ---- circuit shape ------------------------------------------------------------------- */

/* Rows one compression region occupies (292 + 392 * rounds, docs/CIRCUIT.md).  Replaces the
 * row arithmetic of compression/compression_util.rs:32-43,112-205 (dead code upstream). */
int32_t zk_blake2f_rows_per_compression(uint32_t rounds, uint64_t* rows);
/* Smallest k such that n_compressions regions plus the 2^16-row spread table plus the
 * blinding rows fit (reference default k = 17: spread_table.rs:759, benches/blake2f.rs:149). */
int32_t zk_blake2f_min_k(uint32_t rounds, uint64_t n_compressions, int32_t* k);

/* FNV-1a digests of one region's copy constraints (as LE u32 quadruples left col, left row,
 * right col, right row in `copy_advice` call order) and selector activations ([12][R] bytes),
 * so that a host-only test can compare the layout with an independent implementation. */
int32_t zk_blake2f_layout_hash(uint32_t rounds, uint64_t* copies_hash, uint64_t* selectors_hash,
                               uint64_t* n_copies);

/* The layout of one region as tables, for a host-language `Circuit::synthesize` over halo2's own `Layouter`
 * (rust/zkodst-backend/src/halo2_chip.rs): the copy constraints in `copy_advice` call order as quadruples
 * (left advice column, left row, right advice column, right row; rows relative to the region start; *n_copies:
 * in = capacity in quadruples, out = count, copies NULL = size query), the selector activations
 * ([ZK_NUM_SELECTORS][rows] bytes), the constants fixed column ([rows] u64) and the rows of the word cells record
 * chaining connects (chain_rows[0..8) = h_i in advice column 1, chain_rows[8..16) = h'_i in advice column 0).
 * Any output pointer may be NULL.  Replaces the row arithmetic of compression/compression_util.rs:32-43,112-205. */
#define ZK_NUM_SELECTORS 14
int32_t zk_blake2f_layout_tables(uint32_t rounds, uint32_t* copies, uint64_t* n_copies, uint8_t* selectors,
                                 uint64_t* constants, uint32_t chain_rows[16]);

/* ---- EIP-152 wire format and the multi-block hashing driver (host-only helpers) ------------------
 * zk_eip152_validate: the precompile's input checks — length must be 213, f must be 0 or 1
 * (ZK_E_INPUT otherwise); writes the big-endian round count.
 * zk_blake2f_compress: F on one record (the function the circuit proves), 64-byte h' out.
 * zk_blake2b_records: replaces the streaming gadget `Blake2f::{new, update, finalize, digest}`
 * (blake2f-circuit/src/blake2f.rs:88-181): the unkeyed BLAKE2b-512 of `msg` as the chain of
 * ceil(len / 128) (at least 1) EIP-152 records — chaining value, block, byte counter, final flag —
 * ready for zk_blake2f_witness_batch / zk_create_proof; rounds = 12 for real BLAKE2b.
 * *n_records: in = capacity of records_out in records, out = records needed (ZK_E_BUFFER if short;
 * records_out and digest_out NULL = size query).  digest_out: the 64-byte hash, or NULL. */
int32_t zk_eip152_validate(const uint8_t* input, uint64_t len, uint32_t* rounds);
int32_t zk_blake2f_compress(const uint8_t input[ZK_BLAKE2F_INPUT_BYTES], uint8_t out[64]);
int32_t zk_blake2b_records(const uint8_t* msg, uint64_t len, uint32_t rounds, uint8_t* records_out,
                           uint64_t* n_records, uint8_t digest_out[64]);

/* ---- K1: batched witness generation ------------------------------------------------------
 * Replaces `Circuit::synthesize` for the BLAKE2f circuit: Table16Chip::compress
 * (blake2f-circuit/src/blake2f/table16.rs:361-373, `todo!()` upstream),
 * CompressionConfig::{initialize_with_iv, compress, digest} (table16/compression.rs:1078-1149),
 * SpreadVar::with_lookup (table16/spread_table.rs:257-285) and AssignedBits::assign_bits
 * (table16.rs:136-166), i.e. every `region.assign_advice` the chip performs.
 *
 * inputs: n_compressions x 213-byte EIP-152 records, all with the same `rounds`.
 * advice: ZK_NUM_ADVICE columns x 2^k rows x 32 bytes, column-major by halo2 advice column
 *         index, each cell a Montgomery-form pallas::Base — the in-memory image of halo2's
 *         `Vec<Polynomial<Fp, LagrangeCoeff>>` before blinding.  Compression j occupies rows
 *         [j*R, (j+1)*R).  All other cells are zero.
 * digests: n_compressions x 8 u64 (the F outputs h'), or NULL.
 *
 * Host-buffer form: copies inputs H2D, runs the kernels, copies results D2H. */
int32_t zk_blake2f_witness_batch(zk_ctx* ctx, int32_t k, uint32_t rounds,
                                 const uint8_t* inputs, uint64_t n_compressions,
                                 void* advice_out, uint64_t* digests_out);
/* Device-buffer form: all three pointers are device memory on the context's device; runs
 * asynchronously on the context's stream.  d_advice must be 32-byte aligned (cells are written with
 * 256-bit stores; any cudaMalloc pointer is): ZK_E_INVALID otherwise. */
int32_t zk_blake2f_witness_batch_device(zk_ctx* ctx, int32_t k, uint32_t rounds,
                                        const uint8_t* d_inputs, uint64_t n_compressions,
                                        void* d_advice, uint64_t* d_digests);

/* ---- K2/K3: multi-scalar multiplication over Vesta --------------------------------------
 * Replaces halo2_proofs 0.3.0 `best_multiexp` as reached from `Params::commit_lagrange` /
 * `Params::commit` (blake2f-circuit/benches/blake2f.rs:125 `create_proof`; :85 `Params`).
 * scalars: n Montgomery-form Fp (32 B each); bases: n affine points (x, y Montgomery Fq,
 * 64 B each, identity = all zero); out_affine: 64 B.  on_device != 0: both arrays are device
 * pointers on the context's device. */
int32_t zk_msm_vesta(zk_ctx* ctx, const void* scalars, const void* bases, uint64_t n,
                     int32_t on_device, void* out_affine);

/* ---- K4/K5: NTT over Fp -----------------------------------------------------------------
 * Replaces halo2_proofs 0.3.0 `best_fft` behind `EvaluationDomain::lagrange_to_coeff` /
 * `coeff_to_lagrange` (reached from `create_proof`, blake2f-circuit/benches/blake2f.rs:125).
 * data: 2^log_n Montgomery-form Fp, natural order, transformed in place.
 * inverse = 0: coefficients -> evaluations over <omega_n>; 1: evaluations -> coefficients
 * (scaled by 1/n).  omega_n = ROOT_OF_UNITY^(2^(32 - log_n)). */
int32_t zk_ntt_fp(zk_ctx* ctx, void* data, int32_t log_n, int32_t inverse, int32_t on_device);

/* ---- the prover's own batched forms of K2-K5, over the context's params and keys ---------------
 * These are the pipelines create_proof drives (not the stand-alone zk_msm_vesta / zk_ntt_fp above), exported
 * so that each can be compared with a CPU implementation on its own.
 *
 * zk_commit_batch replaces `Params::commit` (basis 0: bases g, coefficient form) and
 * `Params::commit_lagrange` (basis 1: bases g_lagrange) of halo2_proofs 0.3.0 for `ncols` columns at once
 * (create_proof commits the 12 advice columns, the 5 grand products, the 3 pieces of h this way;
 * blake2f-circuit/benches/blake2f.rs:125): out[m] = sum_t scalars[m][t] * base_t + blinds[m] * W.
 * scalars: ncols x 2^k Montgomery Fp, column after column (host, or device when on_device != 0);
 * blinds: ncols Montgomery Fp (host); out_affine: ncols x 64 B (host).  index_mask != 0 (one bit) restricts the sum
 * to the terms t with ((t & index_mask) != 0) == (index_select != 0): the support of the vectors L and R of
 * an inner-product round (`best_multiexp(&p_prime[half..], &g_prime[..half])` in poly/commitment/prover.rs).
 *
 * zk_ntt_fp_batch: `batch` transforms of size 2^log_n in one sequence of launches, transform b reading
 * in + b * 2^log_n and writing out + b * 2^log_n elements (in != out); same conventions as zk_ntt_fp.
 *
 * zk_coeff_to_cosets replaces `EvaluationDomain::coeff_to_extended` for `ncols` polynomials of the keys'
 * domain: the quotient lives on three cosets c_j <omega_n>, c_j = zeta * omega_4n^j (j = 0, 1, 2) of halo2's
 * extended domain of 4n points, so out[col][j][i] = extended(col)[4 i + j]; out: ncols x 3 x 2^k elements. */
int32_t zk_commit_batch(zk_ctx* ctx, int32_t basis, const void* scalars, uint32_t ncols, const void* blinds,
                        uint32_t index_mask, int32_t index_select, int32_t on_device, void* out_affine);
int32_t zk_ntt_fp_batch(zk_ctx* ctx, const void* in, void* out, int32_t log_n, uint32_t batch, int32_t inverse,
                        int32_t on_device);
int32_t zk_coeff_to_cosets(zk_ctx* ctx, const void* coeffs, uint32_t ncols, int32_t on_device, void* out);

/* ---- params, keys, proofs ---------------------------------------------------------------
 * Replace the reference's prove sequence (blake2f-circuit/benches/blake2f.rs:83-142):
 *   Params::<EqAffine>::new / read / write  ->  zk_params_generate_substitute / _load / _write
 *   keygen_vk + keygen_pk                    ->  zk_blake2f_keygen
 *   create_proof(.., rng, &mut transcript); transcript.finalize()  ->  zk_create_proof
 * All state lives inside the context (device memory); one params set and one key set at a time. */

/* Substitute URS with known discrete logs (benchmark / parity use only; Params::new's
 * hash-to-curve cannot be reproduced offline): g_i = [s_i] G with s_i from XorShiftRng(seed). */
int32_t zk_params_generate_substitute(zk_ctx* ctx, int32_t k, const uint8_t seed[16]);
/* halo2 params file format (Params::write): k u32 LE | n x g | n x g_lagrange | w | u, 32 B
 * compressed points. */
int32_t zk_params_load(zk_ctx* ctx, const uint8_t* bytes, uint64_t len);
int32_t zk_params_write(zk_ctx* ctx, uint8_t* out, uint64_t* len);
/* Keys for a circuit of n_compressions regions of `rounds` rounds at the params' k. */
int32_t zk_blake2f_keygen(zk_ctx* ctx, uint32_t rounds, uint64_t n_compressions);
/* The same with record chaining, the in-circuit form of `CompressionConfig::initialize_with_state`
 * (blake2f-circuit/src/blake2f/table16/compression.rs:1096-1111; the streaming gadget's
 * `Blake2f::update`, src/blake2f.rs:101-140): chain[j] != 0 (j >= 1; chain[0] must be 0) copy-constrains the
 * eight h word cells of compression j to the output words h' of compression j - 1, so a proof binds the
 * whole chain of a multi-block hash (zk_blake2b_records emits such chains; records of different messages
 * start with chain[j] = 0).  chain: n_compressions bytes, or NULL for independent compressions.  The
 * chain pattern is part of the verifying key. */
int32_t zk_blake2f_keygen_chained(zk_ctx* ctx, uint32_t rounds, uint64_t n_compressions, const uint8_t* chain);
/* 12 fixed + 8 permutation commitments (32 B compressed each) followed by vk.transcript_repr. */
int32_t zk_vk_bytes(zk_ctx* ctx, uint8_t* out, uint64_t* len);
/* vk.transcript_repr is derived as halo2_proofs 0.3.0 `VerifyingKey::from_parts` does (reached from `keygen_vk`,
 * blake2f-circuit/benches/blake2f.rs:102): BLAKE2b-512 ("Halo2-Verify-Key") of the length-prefixed Rust `{:?}`
 * rendering of `vk.pinned()`, reduced with from_uniform_bytes.  zk_vk_pinned_debug returns that string (not
 * NUL-terminated; *len: in = capacity, out = length, ZK_E_BUFFER if short) so that it can be compared character by
 * character with `format!("{:?}", vk.pinned())` of the Rust circuit (rust/xcheck) — the rendering is restated from
 * the published halo2 sources and has not been run against them in this image. */
int32_t zk_vk_pinned_debug(zk_ctx* ctx, char* out, uint64_t* len);
/* Host-only form for commitments computed elsewhere (e.g. by halo2's own keygen_vk): `commitments` = the 12 fixed
 * then the 8 permutation commitments as 64-byte affine points (Montgomery x, y). */
int32_t zk_blake2f_pinned_debug(int32_t k, uint32_t rounds, const void* commitments, char* out, uint64_t* len);
/* Inject a `vk.transcript_repr` obtained from halo2 itself (32 B canonical LE) in place of the derived one. */
int32_t zk_vk_repr_override(zk_ctx* ctx, const uint8_t repr[32]);
/* inputs: n_compressions x 213 B (host).  seed: 16-byte XorShiftRng seed (the reference harness
 * seeds its prover RNG the same way, benchmarking/src/blake2f_circuit_bench.rs:41-44).
 * proof_out / proof_len: caller buffer and its capacity; the proof length is written back
 * (ZK_E_BUFFER if too small). */
int32_t zk_create_proof(zk_ctx* ctx, const uint8_t* inputs, uint64_t n_compressions,
                        const uint8_t seed[16], uint8_t* proof_out, uint64_t* proof_len);

/* Same with the EIP-152 records already resident in device memory (record validation then
 * happens in the witness kernel and is reported as ZK_E_INPUT). */
int32_t zk_create_proof_device_inputs(zk_ctx* ctx, const uint8_t* d_inputs, uint64_t n_compressions,
                                      const uint8_t seed[16], uint8_t* proof_out, uint64_t* proof_len);

/* verify_proof with the SingleVerifier strategy (blake2f-circuit/benches/blake2f.rs:138-144:
 * `verify_proof(&params, pk.get_vk(), SingleVerifier::new(&params), &[&[]], &mut Blake2bRead)`)
 * against the context's params and keys.  ZK_OK = accepted; ZK_E_VERIFY = rejected (reason in
 * zk_last_error: malformed encoding, truncated / trailing bytes, final MSM not the identity). */
int32_t zk_verify_proof(zk_ctx* ctx, const uint8_t* proof, uint64_t proof_len);
/* Batch verification (halo2_proofs 0.3.0 `plonk::verifier::BatchVerifier`; the reference verifies one proof per
 * criterion iteration, benches/blake2f.rs:138-144): `count` proofs of the context's circuit, concatenated in `proofs`
 * with lengths `proof_lens`.  The transcript phase of each proof runs on the host; their final multi-scalar
 * multiplications are combined with weights drawn from XorShiftRng(seed) into ONE size-n fixed-base MSM and one
 * variable-base MSM.  ZK_OK = all accepted; ZK_E_VERIFY = at least one is rejected (malformed proofs are reported as
 * they are met; a failing combined MSM does not say which proof is wrong — fall back to zk_verify_proof). */
int32_t zk_verify_proofs_batch(zk_ctx* ctx, const uint8_t* proofs, const uint64_t* proof_lens, uint64_t count,
                               const uint8_t seed[16]);

/* `MockProver::run(k, &circuit, vec![]).verify()` (blake2f/table16/spread_table.rs:759-763) for the
 * BLAKE2f circuit of the context's keys: every gate on every usable row, every lookup input against
 * the spread table, every copy constraint.  The witness comes from `inputs` (n_compressions x 213 B,
 * host) or, when advice_override is not NULL, from that host buffer (ZK_NUM_ADVICE x 2^k x 32 B,
 * column-major Montgomery cells, the layout of zk_blake2f_witness_batch).  ZK_E_VERIFY on the first
 * failure, described in failure[3] = {kind (1 gate, 2 lookup, 3 copy), row, gate / copy index; a failed
 * chaining copy of zk_blake2f_keygen_chained reports index 0xffff}. */
int32_t zk_mock_verify(zk_ctx* ctx, const uint8_t* inputs, uint64_t n_compressions,
                       const void* advice_override, uint64_t failure[3]);

/* ---- one MSM split across the GPUs of a box (BASELINE configs[3]) --------------
 * The reference is single-threaded and has no counterpart; these calls are what a multi-process
 * Rust harness (one process per GPU) adds around `Params::new` / `create_proof`
 * (blake2f-circuit/benches/blake2f.rs:85,125).  After zk_dist_init every MSM of the context
 * (`Params::commit`, `commit_lagrange`, the IPA rounds) covers only this rank's contiguous range
 * of the base points; the per-rank partial points are all-gathered over NCCL and summed, so all
 * ranks obtain the same commitment and — fed the same records and seed — the same proof bytes as
 * a single GPU.  A group also shards the transforms of the witness columns by column
 * (zk_dist_column_block), the quotient by row (zk_dist_quotient_rows) and the polynomial evaluations by
 * coefficient range; the remaining steps run replicated.
 *   rank 0: zk_dist_unique_id(id); share id with the other ranks by any means
 *   all:    zk_ctx_create; zk_dist_init(ctx, id, rank, world); params; keygen; create_proof */
#define ZK_DIST_ID_BYTES 128
int32_t zk_dist_unique_id(uint8_t out[ZK_DIST_ID_BYTES]);
/* Collective over the group; must precede zk_params_* on the context.  world == 1 is a no-op. */
int32_t zk_dist_init(zk_ctx* ctx, const uint8_t id[ZK_DIST_ID_BYTES], int32_t rank, int32_t world);
int32_t zk_dist_info(const zk_ctx* ctx, int32_t* rank, int32_t* world);
/* The contiguous range [lo, hi) of the n base points that `rank` of `world` tabulates and sums
 * (host-only helper; the partition the MSM split uses). */
int32_t zk_dist_range(uint64_t n_points, int32_t rank, int32_t world, uint64_t* lo, uint64_t* hi);
/* How a group shards the transforms of the witness columns (SURVEY.md 8e "NTT only across columns"):
 * the ZK_NUM_WITNESS_COLUMNS columns (12 advice, permuted input, permuted table, 4 permutation products,
 * lookup product) are cut into blocks of *per_rank = ceil(columns / world) slots; `rank` transforms
 * slots [lo, hi) (empty for the last ranks when world does not divide the count) and the slot arrays,
 * padded to per_rank * world slots, are all-gathered in place.  Host-only helper. */
#define ZK_NUM_WITNESS_COLUMNS 19
int32_t zk_dist_column_block(int32_t rank, int32_t world, uint32_t* lo, uint32_t* hi, uint32_t* per_rank);
/* How a group shards the quotient: `rank` evaluates rows [*row_lo, *row_hi) of the coset-major 3n-row
 * domain (3n / world rows; ZK_E_INVALID when world does not divide 3n) and, because a row reads its
 * rotations -1, +1, -6 inside its own coset, needs the rows listed in `segments` (pairs start, length;
 * sorted, disjoint) of every column — the rows the other ranks send it.  *n_segments: in = capacity in
 * pairs, out = pairs needed (ZK_E_BUFFER if short).  Host-only helper. */
int32_t zk_dist_quotient_rows(uint64_t n, int32_t rank, int32_t world, uint64_t* row_lo, uint64_t* row_hi,
                              uint64_t* segments, uint32_t* n_segments);

#ifdef __cplusplus
}
#endif
#endif /* ZKODST_H */

"""The prover's own batched pipelines at the published size (n = 2^19, BASELINE configs[2]) against the CPU
oracle, kernel by kernel and then as a whole proof:

  * zk_commit_batch  (msm_fixed.cu: digits, counting sort, chunked accumulation, heavy buckets, bucket
    reduction) = Params::commit / commit_lagrange, for the scalar shapes a proof produces: full-width,
    all-zero, carries 0/1/2, 16-bit dense cells, 32-bit spread cells, p - 1 everywhere, one heavy bucket, and
    the index-masked L / R vectors of the inner-product rounds;
  * zk_ntt_fp_batch and zk_coeff_to_cosets (ntt.cu batched passes, blockIdx.y / z, fused coset scaling);
  * configs[2] itself: 64 twelve-round compressions at k = 19, reference seed, proof bytes and verifying key
    equal to the oracle's (blake2f-circuit/benches/blake2f.rs:125 `create_proof`).
"""
import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
K = 19
N = 1 << K


def mont(oracle, v):
    """Montgomery limbs of the integer v (oracle field op 4 = from canonical)."""
    return oracle_lib.Oracle._limbs(oracle.field_op(0, 4, int(v))[1])


def mont_column(oracle, values):
    """Montgomery form of small integers, vectorised through a lookup of the distinct values."""
    values = np.asarray(values, dtype=np.uint64)
    uniq, inv = np.unique(values, return_inverse=True)
    table = np.stack([mont(oracle, int(u)) for u in uniq])
    return np.ascontiguousarray(table[inv])


@pytest.fixture(scope="module")
def setup19(oracle, zk):
    seed = zk.REFERENCE_SEED
    ctx = zk.Context(0)
    ctx.params_generate_substitute(K, seed)
    op = oracle_lib.OracleProver(oracle, k=K, seed=seed)
    assert ctx.params_write() == op.params_bytes(), "URS differs from the oracle's"
    bases = {0: op.points(0, N), 1: op.points(1, N)}
    w = op.points(2, 0)[0]
    yield ctx, op, bases, w, seed
    op.close()
    ctx.close()


def oracle_commit(oracle, scalars, bases, blind, w):
    s = np.concatenate([scalars, blind[None, :]])
    b = np.concatenate([bases, w[None, :]])
    return oracle_lib.msm(oracle, np.ascontiguousarray(s), np.ascontiguousarray(b))


def shaped_columns(oracle):
    """name -> (n, 4) Montgomery scalars, the shapes the commitments of one proof take."""
    rnd = np.random.RandomState(19)
    p = oracle.consts(0)["MOD"]
    cols = {}
    cols["full_width"] = oracle_lib.random_fields(oracle, bytes(range(3, 19)), N)
    cols["all_zero"] = np.zeros((N, 4), dtype=np.uint64)
    cols["carries_0_1_2"] = mont_column(oracle, rnd.randint(0, 3, size=N))
    dense = rnd.randint(0, 1 << 16, size=N).astype(np.uint64)
    cols["dense_16bit"] = mont_column(oracle, dense)
    # spread form of 16-bit values: bits interleaved with zeros (32-bit cells of the a_2 column)
    spread = np.zeros(N, dtype=np.uint64)
    for b in range(16):
        spread |= ((dense >> np.uint64(b)) & np.uint64(1)) << np.uint64(2 * b)
    cols["spread_32bit"] = mont_column(oracle, spread)
    cols["p_minus_1"] = np.tile(mont(oracle, p - 1), (N, 1))
    heavy = oracle_lib.random_fields(oracle, bytes(range(5, 21)), N)
    heavy[: N - 1000] = mont(oracle, 0x1234_5678_9ABC_DEF0_0FED_CBA9_8765_4321)
    cols["one_heavy_bucket"] = heavy
    return cols


@pytest.mark.parametrize("basis", [0, 1])
def test_commit_batch_shapes_at_2p19(setup19, oracle, basis):
    ctx, _, bases, w, _ = setup19
    cols = shaped_columns(oracle)
    names = sorted(cols)
    scalars = np.ascontiguousarray(np.stack([cols[n] for n in names]))
    blinds = oracle_lib.random_fields(oracle, bytes(range(7, 23)), len(names))
    out = np.zeros((len(names), 8), dtype=np.uint64)
    ctx.commit_batch(basis, scalars, len(names), blinds, out)
    for i, name in enumerate(names):
        want = oracle_commit(oracle, cols[name], bases[basis], blinds[i], w)
        assert np.array_equal(out[i], want), name


def test_commit_batch_more_columns_than_one_pipeline(setup19, oracle):
    """17 columns = one full pipeline of 16 jobs and one of 1; columns repeat so the oracle runs 3 MSMs."""
    ctx, _, bases, w, _ = setup19
    kinds = [oracle_lib.random_fields(oracle, bytes([s] * 16), N) for s in (1, 2, 3)]
    scalars = np.ascontiguousarray(np.stack([kinds[i % 3] for i in range(17)]))
    one = mont(oracle, 1)
    blinds = np.tile(one, (17, 1))
    out = np.zeros((17, 8), dtype=np.uint64)
    ctx.commit_batch(1, scalars, 17, blinds, out)
    for j in range(3):
        want = oracle_commit(oracle, kinds[j], bases[1], one, w)
        for i in range(j, 17, 3):
            assert np.array_equal(out[i], want), (i, j)


@pytest.mark.parametrize("mask,select", [(N >> 1, 0), (N >> 1, 1), (1 << 3, 1), (1, 0)])
def test_commit_batch_index_mask(setup19, oracle, mask, select):
    """The L / R vectors of an inner-product round: terms whose index has (or lacks) one bit."""
    ctx, _, bases, w, _ = setup19
    sc = oracle_lib.random_fields(oracle, bytes(range(9, 25)), N)
    blind = mont(oracle, 7)
    out = np.zeros((1, 8), dtype=np.uint64)
    ctx.commit_batch(0, sc, 1, blind, out, index_mask=mask, index_select=select)
    keep = ((np.arange(N) & mask) != 0) == bool(select)
    masked = sc.copy()
    masked[~keep] = 0
    assert np.array_equal(out[0], oracle_commit(oracle, masked, bases[0], blind, w))


def test_commit_batch_is_linear_at_2p19(setup19, oracle):
    """Size-independent property on full-width columns: commit(a) + commit(b) == commit(a + b) (blinds add too)."""
    ctx, _, _, _, _ = setup19
    a = oracle_lib.random_fields(oracle, bytes(range(11, 27)), N)
    b = oracle_lib.random_fields(oracle, bytes(range(13, 29)), N)
    # a + b mod p on 4 x u64 limbs (Montgomery form is linear): python integers on a sparse support
    idx = np.arange(0, N, 257)
    sa, sb, sab = (np.zeros((N, 4), dtype=np.uint64) for _ in range(3))
    sa[idx], sb[idx] = a[idx], b[idx]
    for i in idx:
        sab[i] = oracle_lib.Oracle._limbs(oracle.field_op(0, 1, oracle_lib.Oracle._int(a[i]),
                                                          oracle_lib.Oracle._int(b[i]))[1])
    blinds = np.stack([mont(oracle, 3), mont(oracle, 4), mont(oracle, 7)])
    out = np.zeros((3, 8), dtype=np.uint64)
    ctx.commit_batch(1, np.ascontiguousarray(np.stack([sa, sb, sab])), 3, blinds, out)
    both = np.ascontiguousarray(out[:2])
    total = np.zeros(8, dtype=np.uint64)
    ctx.msm(np.stack([mont(oracle, 1)] * 2), both, 2, total)
    assert np.array_equal(total, out[2])


@pytest.mark.parametrize("log_n,batch", [(19, 5), (17, 19), (12, 40), (10, 3)])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_batch_matches_oracle(ctx, oracle, log_n, batch, inverse):
    n = 1 << log_n
    data = oracle_lib.random_fields(oracle, bytes(range(2, 18)), n * batch).reshape(batch, n, 4)
    out = np.zeros_like(data)
    ctx.ntt_batch(data, out, log_n, batch, inverse=inverse)
    for b in range(batch):
        assert np.array_equal(out[b], oracle_lib.ntt(oracle, data[b], log_n, inverse)), b


def test_coeff_to_cosets_matches_oracle(setup19, oracle):
    """Three cosets of the n-th roots = rows 4i + j (j < 3) of halo2's extended domain of 4n points."""
    ctx, _, _, _, _ = setup19
    ctx.keygen(12, 64)
    ncols = 4
    coeffs = oracle_lib.random_fields(oracle, bytes(range(4, 20)), N * ncols).reshape(ncols, N, 4)
    out = np.zeros((ncols, 3, N, 4), dtype=np.uint64)
    ctx.coeff_to_cosets(coeffs, ncols, out)
    for c in range(ncols):
        ext = oracle_lib.coeff_to_extended(oracle, coeffs[c], K).reshape(N, 4, 4)  # [i][j] = extended[4 i + j]
        for j in range(3):
            assert np.array_equal(out[c, j], ext[:, j, :]), (c, j)


def test_config2_proof_bytes_match_oracle_k19(setup19, zk):
    """BASELINE configs[2] — the configuration bench.py publishes: 64 twelve-round compressions at k = 19."""
    ctx, op, _, _, seed = setup19
    n = 64
    assert zk.min_k(12, n) == K
    inputs = zk.synthetic_inputs(n)
    ctx.keygen(12, n)
    op.keygen(12, n)
    assert ctx.vk_bytes() == op.vk_bytes(), "verifying key differs from the oracle's"
    proof = ctx.create_proof(inputs, n, seed)
    ref = op.create_proof(inputs, n, seed)
    assert len(proof) == len(ref) == 4192
    diff = [i // 32 for i in range(0, len(ref), 32) if proof[i:i + 32] != ref[i:i + 32]]
    assert not diff, "proof chunks %s differ" % diff[:8]
    assert op.verify(proof)[0] == 0
    assert ctx.verify_proof(ref)
    assert ctx.mock_verify(inputs, n) is None


def test_k18_proof_bytes_match_oracle(oracle, zk):
    """The smallest batch that leaves k = 17 (27 compressions -> k = 18): another transform plan (9 + 9 stages) and
    MSM size; keys and proof bytes equal the oracle's."""
    n = 27
    k = zk.min_k(12, n)
    assert k == 18
    seed = zk.REFERENCE_SEED
    ctx = zk.Context(0)
    ctx.params_generate_substitute(k, seed)
    ctx.keygen(12, n)
    op = oracle_lib.OracleProver(oracle, k=k, seed=seed)
    op.keygen(12, n)
    assert ctx.vk_bytes() == op.vk_bytes()
    inputs = zk.synthetic_inputs(n)
    proof = ctx.create_proof(inputs, n, seed)
    assert proof == op.create_proof(inputs, n, seed)
    assert len(proof) == 4064 + 64          # two more IPA rounds than k = 17: one (L, R) pair
    op.close()
    ctx.close()


def test_k20_proof_accepted_by_oracle_verifier(oracle, zk):
    """k = 20 (128 compressions): the oracle's keygen_vk of the same circuit gives the same verifying key and its
    verify_proof accepts the GPU proof (the oracle's own prover would take minutes at this size)."""
    n = 128
    k = zk.min_k(12, n)
    assert k == 20
    seed = zk.REFERENCE_SEED
    ctx = zk.Context(0)
    ctx.params_generate_substitute(k, seed)
    ctx.keygen(12, n)
    inputs = zk.synthetic_inputs(n)
    proof = ctx.create_proof(inputs, n, seed)
    assert ctx.verify_proof(proof) and ctx.mock_verify(inputs, n) is None
    op = oracle_lib.OracleProver(oracle, k=k, seed=seed)
    op.keygen_vk(12, n)
    assert ctx.vk_bytes() == op.vk_bytes()
    rc, msg = op.verify(proof)
    assert rc == 0, msg
    bad = bytearray(proof)
    bad[700] ^= 4
    assert op.verify(bytes(bad))[0] != 0 and not ctx.verify_proof(bytes(bad))
    op.close()
    ctx.close()


def test_commit_batch_rejects_a_mask_of_two_bits(setup19, zk):
    ctx, _, _, _, _ = setup19
    sc = np.zeros((N, 4), dtype=np.uint64)
    out = np.zeros((1, 8), dtype=np.uint64)
    with pytest.raises(zk.ZkError):
        ctx.commit_batch(0, sc, 1, np.zeros((1, 4), dtype=np.uint64), out, index_mask=6, index_select=1)

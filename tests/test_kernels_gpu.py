"""K2-K5 parity: MSM and NTT through the C ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
SEED = bytes(range(1, 17))


@pytest.fixture(scope="module")
def urs(oracle):
    p = oracle_lib.OracleProver(oracle, k=14, seed=SEED)
    g = p.points(0, 1 << 14)
    p.close()
    return g


def small_scalars(oracle, values):
    out = np.zeros((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        out[i] = oracle_lib.Oracle._limbs(oracle.field_op(0, 4, int(v))[1])
    return out


@pytest.mark.parametrize("n", [1, 2, 31, 500, 5000, 1 << 14])
def test_msm_random_scalars(ctx, oracle, urs, n):
    scalars = oracle_lib.random_fields(oracle, SEED, n)
    bases = np.ascontiguousarray(urs[:n])
    out = np.zeros(8, dtype=np.uint64)
    ctx.msm(scalars, bases, n, out)
    assert np.array_equal(out, oracle_lib.msm(oracle, scalars, bases))


def test_msm_edge_cases(ctx, oracle, urs):
    n = 4096
    bases = np.ascontiguousarray(urs[:n])
    out = np.zeros(8, dtype=np.uint64)
    # all-zero scalars -> identity (encoded as all-zero affine)
    zeros = np.zeros((n, 4), dtype=np.uint64)
    ctx.msm(zeros, bases, n, out)
    assert not out.any()
    # skewed tiny scalars (tag / carry columns): exercises the heavy-bucket path
    rnd = np.random.RandomState(7)
    vals = rnd.randint(0, 3, size=n)
    sc = small_scalars(oracle, vals)
    ctx.msm(sc, bases, n, out)
    assert np.array_equal(out, oracle_lib.msm(oracle, sc, bases))
    # p - 1 everywhere (maximal scalar, signed-digit carries ripple to the top window)
    pm1 = oracle.consts(0)["MOD"] - 1
    sc = small_scalars(oracle, [pm1] * 64)
    b64 = np.ascontiguousarray(urs[:64])
    ctx.msm(sc, b64, 64, out)
    assert np.array_equal(out, oracle_lib.msm(oracle, sc, b64))
    # repeated base with cancelling scalars: s*P + (p - s)*P = identity
    sc = small_scalars(oracle, [5, oracle.consts(0)["MOD"] - 5])
    b2 = np.ascontiguousarray(np.stack([urs[3], urs[3]]))
    ctx.msm(sc, b2, 2, out)
    assert not out.any()
    # 16-bit dense column shape: many equal points in few buckets
    vals = rnd.randint(0, 1 << 16, size=n)
    sc = small_scalars(oracle, vals)
    ctx.msm(sc, bases, n, out)
    assert np.array_equal(out, oracle_lib.msm(oracle, sc, bases))


def test_msm_linearity(ctx, oracle, urs):
    """MSM(a) + MSM(b) == MSM(a + b): a size-independent property, checked via the oracle add."""
    n = 1 << 14
    a = oracle_lib.random_fields(oracle, SEED, n)
    b = oracle_lib.random_fields(oracle, bytes(range(2, 18)), n)
    ab = np.zeros_like(a)
    for i in range(0, n, 97):  # sparse sample keeps the python loop short; other rows are zero
        ab[i] = oracle_lib.Oracle._limbs(oracle.field_op(0, 1, oracle_lib.Oracle._int(a[i]),
                                                          oracle_lib.Oracle._int(b[i]))[1])
    mask = np.zeros(n, dtype=bool)
    mask[::97] = True
    a[~mask] = 0
    b[~mask] = 0
    out_a, out_b, out_ab = (np.zeros(8, dtype=np.uint64) for _ in range(3))
    ctx.msm(a, urs, n, out_a)
    ctx.msm(b, urs, n, out_b)
    ctx.msm(ab, urs, n, out_ab)
    # out_a + out_b via an MSM with unit scalars over the two results
    one = small_scalars(oracle, [1, 1])
    both = np.ascontiguousarray(np.stack([out_a, out_b]))
    out_sum = np.zeros(8, dtype=np.uint64)
    ctx.msm(one, both, 2, out_sum)
    assert np.array_equal(out_sum, out_ab)


# every round plan of the pass scheduler: single passes of 1..10 stages (rounds 3/2/1), two passes
# (11 = 6 + 5 ... 20 = 10 + 10), three passes (21 = 7 + 7 + 7, 23 = 8 + 8 + 7) and the column-grouped later passes
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 17, 19, 20, 21, 23])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(ctx, oracle, log_n, inverse):
    n = 1 << log_n
    data = oracle_lib.random_fields(oracle, SEED, n)
    ref = oracle_lib.ntt(oracle, data, log_n, inverse)
    got = data.copy()
    ctx.ntt(got, log_n, inverse=inverse)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("log_n", [21, 22, 24, 25])
def test_ntt_roundtrip_large(ctx, oracle, log_n):
    """Full-size property: iNTT(NTT(x)) == x at the extended-domain sizes of configs 3 and 4
    (2^21 .. 2^25: three passes of 7 to 9 stages over 1024-element tiles)."""
    import torch
    n = 1 << log_n
    data = oracle_lib.random_fields(oracle, SEED, 1 << 12)
    big = np.tile(data, (n >> 12, 1))
    big[:, 0] ^= np.arange(n, dtype=np.uint64) & np.uint64(0xFFFF)  # make rows distinct
    for i in range(0, n, 1 << 15):  # keep the top limb reduced
        pass
    big[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    d = torch.from_numpy(big.view(np.int64)).cuda()
    ctx.ntt(d, log_n, inverse=False, on_device=True)
    ctx.ntt(d, log_n, inverse=True, on_device=True)
    ctx.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), big)

"""verify_proof and the MockProver-equivalent checker through the C ABI (zk_verify_proof,
zk_mock_verify), cross-checked against the oracle: each side must accept the other's proofs and
both must reject the same tampered ones."""
import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(oracle, zk):
    seed = zk.REFERENCE_SEED
    ctx = zk.Context(0)
    ctx.params_generate_substitute(17, seed)
    ctx.keygen(12, 2)
    op = oracle_lib.OracleProver(oracle, k=17, seed=seed)
    op.keygen(12, 2)
    inputs = zk.synthetic_inputs(2)
    yield ctx, op, seed, inputs
    op.close()
    ctx.close()


def test_device_verifier_accepts_device_and_oracle_proofs(setup):
    ctx, op, seed, inputs = setup
    proof = ctx.create_proof(inputs, 2, seed)
    assert ctx.verify_proof(proof), ctx.last_error()
    assert op.verify(proof)[0] == 0
    ref = op.create_proof(inputs, 2, bytes(range(16)))  # a different prover seed
    assert ref != proof
    assert ctx.verify_proof(ref), ctx.last_error()


def test_tampering_is_rejected_like_the_oracle(setup):
    ctx, op, seed, inputs = setup
    proof = ctx.create_proof(inputs, 2, seed)
    n_points_head = 12 + 2 + 4 + 1 + 1 + 3  # commitments before the evaluations
    offsets = [5, 32 * 13 + 3, 32 * n_points_head + 7, 32 * (n_points_head + 30) + 1,
               len(proof) - 40, len(proof) - 1]
    for off in offsets:
        bad = bytearray(proof)
        bad[off] ^= 0x04
        mine = ctx.verify_proof(bytes(bad))
        theirs = op.verify(bytes(bad))[0] == 0
        assert not mine and not theirs, off
    assert not ctx.verify_proof(proof[:-32])            # truncated
    assert not ctx.verify_proof(proof + b"\x00" * 32)   # trailing bytes
    assert ctx.verify_proof(proof)                        # the context is still usable


def test_wrong_statement_is_rejected(setup, zk):
    ctx, op, seed, inputs = setup
    proof = ctx.create_proof(inputs, 2, seed)
    ctx.keygen(12, 1)   # a different circuit (vk) at the same params
    try:
        assert not ctx.verify_proof(proof)
    finally:
        ctx.keygen(12, 2)


def test_mock_verify_passes_and_locates_failures(setup, oracle, zk):
    ctx, op, seed, inputs = setup
    assert ctx.mock_verify(inputs, 2) is None
    adv, _, _ = oracle.witness(17, 12, inputs, 2)
    adv = np.ascontiguousarray(adv)
    assert ctx.mock_verify(None, 2, advice_override=adv) is None
    R = zk.rows_per_compression(12)
    one = oracle_lib.Oracle._limbs(oracle.field_op(0, 4, 1)[1])
    # break a lookup row: dense cell (halo2 column 8) of some row inside region 1
    bad = adv.copy()
    row = R + 100
    bad[8, row] = (bad[8, row] + np.array([1, 0, 0, 0], dtype=np.uint64))
    fail = ctx.mock_verify(None, 2, advice_override=bad)
    assert fail is not None and fail[0] in (1, 2) and abs(int(fail[1]) - row) <= 1
    rc, _ = oracle.mock_verify_mont(17, 12, 2, bad)
    assert rc != 0
    # tamper single cells (small value, so the cell stays a valid 64-bit word): the device checker
    # and the oracle's MockProver-equivalent must agree on every one
    val = np.array(oracle_lib.Oracle._limbs(oracle.field_op(0, 4, 12345)[1]), dtype=np.uint64)
    kinds = set()
    cells = [(8, 3), (9, 4), (1, 2), (0, 5), (4, 9), (1, 300), (4, 299), (4, 308), (4, 311), (4, 332), (0, 401),
             (1, 293), (8, 290), (8, 377), (4, R + 700), (2, 2000), (5, 4000), (10, 5),
             # pinned inputs: the IV_0 word cell (pin constant), the final-flag bit and word (final flag)
             (1, 33), (6, 137), (1, 137), (1, R + 33), (6, R + 137)]
    for col, r in cells:
        bad = adv.copy()
        bad[col, r] = val
        fail = ctx.mock_verify(None, 2, advice_override=bad)
        ref_fails = oracle.mock_verify_mont(17, 12, 2, bad)[0] != 0
        assert (fail is not None) == ref_fails, (col, r, fail)
        if fail is not None:
            kinds.add(int(fail[0]))
    assert {1, 2} <= kinds


def test_batch_verification(setup, zk):
    """zk_verify_proofs_batch (halo2's BatchVerifier): several proofs, one combined final MSM; a single bad proof
    anywhere in the batch is caught, as are malformed bytes."""
    ctx, op, seed, inputs = setup
    proofs = [ctx.create_proof(inputs, 2, bytes([i]) * 16) for i in range(1, 5)]
    assert len(set(proofs)) == 4
    wseed = bytes(range(16))
    assert ctx.verify_proofs_batch(proofs, wseed)
    assert ctx.verify_proofs_batch(proofs[:1], wseed)
    for victim in range(4):
        bad = list(proofs)
        b = bytearray(bad[victim])
        b[-20] ^= 1                      # the final scalar f: only the combined MSM can notice
        bad[victim] = bytes(b)
        assert not ctx.verify_proofs_batch(bad, wseed), victim
        assert op.verify(bad[victim])[0] != 0
    b = bytearray(proofs[2])
    b[5] ^= 0x40                         # a commitment: a different transcript, rejected by the batch as well
    assert not ctx.verify_proofs_batch([proofs[0], bytes(b)], wseed)
    assert not ctx.verify_proofs_batch([proofs[0], proofs[1][:-1]], wseed)

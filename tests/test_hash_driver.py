"""EIP-152 wire format and the multi-block hashing driver (host-only entry points of the C ABI;
the reference's `Blake2f::{update, finalize}`, blake2f-circuit/src/blake2f.rs:88-181)."""
import hashlib
import json
import os

import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
VECS = json.load(open(os.path.join(G, "eip152.json")))


@pytest.mark.parametrize("idx", range(len(VECS)))
def test_compress_matches_eip152_vectors(zk, idx):
    v = VECS[idx]
    assert zk.blake2f_compress(bytes.fromhex(v["input"])).hex() == v["output"]


def test_eip152_rejections(zk):
    rec = bytes.fromhex(VECS[1]["input"])
    assert zk.eip152_validate(rec) == int.from_bytes(rec[:4], "big")
    for bad in (rec[:-1], rec + b"\x00", b"", rec[:212] + b"\x02"):   # EIP-152 vectors 0-3
        with pytest.raises(zk.ZkError) as e:
            zk.eip152_validate(bad)
        assert e.value.code == -5


@pytest.mark.parametrize("length", [0, 1, 3, 127, 128, 129, 255, 256, 257, 1000, 4096])
def test_record_chain_hashes_like_hashlib(zk, oracle, length):
    msg = bytes((i * 7 + 3) & 0xFF for i in range(length))
    records, digest = zk.blake2b_records(msg)
    assert digest == hashlib.blake2b(msg).digest()
    n = len(records) // 213
    assert n == max(1, (length + 127) // 128)
    # replay the chain with the ORACLE's F: every record's h is the previous record's output
    h = None
    for i in range(n):
        rec = records[213 * i: 213 * (i + 1)]
        assert rec[:4] == (12).to_bytes(4, "big") and rec[212] == (1 if i == n - 1 else 0)
        if h is not None:
            assert rec[4:68] == h
        rc, h = oracle.blake2f(rec)
        assert rc == 0
    assert h == digest


@pytest.mark.gpu
def test_prove_a_multi_block_hash(zk, ctx, oracle):
    """A 3-block message proved as one circuit; the digest cells hold BLAKE2b-512(msg)."""
    import numpy as np
    import oracle_lib
    msg = b"zk-odst" * 50
    records, digest = zk.blake2b_records(msg)
    n = len(records) // 213
    assert n == 3
    adv = np.empty((12, 1 << 17, 4), dtype=np.uint64)
    dig = np.empty((n, 8), dtype=np.uint64)
    ctx.witness_batch(17, 12, records, n, adv, dig)
    assert dig[n - 1].tobytes() == digest == hashlib.blake2b(msg).digest()
    seed = zk.REFERENCE_SEED
    ctx.params_generate_substitute(17, seed)
    ctx.keygen(12, n)
    assert ctx.mock_verify(records, n) is None
    proof = ctx.create_proof(records, n, seed)
    assert ctx.verify_proof(proof)
    op = oracle_lib.OracleProver(oracle, params_bytes=ctx.params_write())
    op.keygen(12, n)
    assert op.verify(proof)[0] == 0
    # The chained circuit (zk_blake2f_keygen_chained; CompressionConfig::initialize_with_state,
    # compression.rs:1096-1111): h of block i + 1 is copy-constrained to the output of block i, so the proof binds
    # the whole hash.  Keys, proof bytes and verdicts equal the oracle's; unrelated records are rejected.
    chain = bytes([0, 1, 1])
    ctx.keygen(12, n, chain=chain)
    op.keygen_chained(12, n, chain)
    assert ctx.vk_bytes() == op.vk_bytes()
    assert ctx.mock_verify(records, n) is None
    chained = ctx.create_proof(records, n, seed)
    assert chained == op.create_proof(records, n, seed)
    assert chained != proof
    assert ctx.verify_proof(chained) and op.verify(chained)[0] == 0
    assert not ctx.verify_proof(proof) and op.verify(proof)[0] != 0      # a proof for the unchained keys
    unrelated = zk.synthetic_inputs(n)
    fail = ctx.mock_verify(unrelated, n)
    assert fail is not None and fail[0] == 3 and fail[2] == 0xffff, fail  # a chaining copy
    assert not ctx.verify_proof(ctx.create_proof(unrelated, n, seed))
    with pytest.raises(zk.ZkError):
        ctx.keygen(12, n, chain=bytes([1, 0, 0]))
    op.close()

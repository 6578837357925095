"""The completed BLAKE2f circuit in the oracle: EIP-152 outputs appear in the digest cells and
the MockProver-equivalent check (the reference's only live assertion,
spread_table.rs:759-763) passes; tampering is rejected."""
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
VECS = json.load(open(os.path.join(G, "eip152.json")))


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 5])
def test_vector_mock_prover(oracle, idx):
    v = VECS[idx]
    rec = bytes.fromhex(v["input"])
    rounds = int.from_bytes(rec[:4], "big")
    _, raw, dig = oracle.witness(17, rounds, rec, 1, mont=False, raw=True)
    assert dig.tobytes().hex() == v["output"]
    rc, msg = oracle.mock_verify_raw(17, rounds, 1, raw)
    assert rc == 0, msg


def test_rows_and_capacity(oracle, zk):
    assert oracle.rows_per_compression(12) == 4996 == zk.rows_per_compression(12)
    assert oracle.rows_per_compression(0) == 292
    assert zk.min_k(12, 1) == 17 and zk.min_k(12, 26) == 17 and zk.min_k(12, 27) == 18
    assert zk.min_k(12, 64) == 19 and zk.min_k(12, 256) == 21 and zk.min_k(12, 1024) == 23


def test_batch_mock_prover_and_tamper(oracle, zk):
    n = 3
    inputs = zk.synthetic_inputs(n)
    _, raw, dig = oracle.witness(17, 12, inputs, n, mont=False, raw=True)
    for i in range(n):
        rc, out = oracle.blake2f(inputs[213 * i:213 * (i + 1)])
        assert rc == 0 and out == dig[i].tobytes()
    assert oracle.mock_verify_raw(17, 12, n, raw)[0] == 0
    R = 4996
    # a lookup cell, a carry cell, a copied cell, a digest word: each must be caught
    for col, row in [(8, R + 1000), (6, 165), (1, 2 * R + 164), (0, 3 * R - 6)]:
        bad = raw.copy()
        bad[col, row] ^= 1
        rc, msg = oracle.mock_verify_raw(17, 12, n, bad)
        assert rc == 1, (col, row, msg)


def test_layout_matches_product_tables(oracle, zk):
    for rounds in (0, 1, 2, 12):
        assert oracle.layout_hash(rounds) == zk.layout_hash(rounds)


def test_circuit_description(oracle):
    d = oracle.describe(17, 12, 1)
    assert d["degree"] == "4" and d["blinding_factors"] == "5"
    assert d["num_advice"] == "12" and d["num_fixed"] == "10"
    assert d["permutation_columns"] == "8;9;1;2;0;3;4;5;"
    assert d["n_polys"] == "23"

"""K1 parity: the CUDA witness path through the C ABI vs the oracle, bit-exact."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
VECS = json.load(open(os.path.join(G, "eip152.json")))


def gpu_witness(ctx, k, rounds, inputs, n):
    adv = np.empty((12, 1 << k, 4), dtype=np.uint64)
    dig = np.empty((max(n, 1), 8), dtype=np.uint64)
    ctx.witness_batch(k, rounds, inputs, n, adv, dig)
    return adv, dig[:n]


@pytest.mark.parametrize("idx", range(len(VECS)))
def test_golden_vectors_bit_exact(ctx, oracle, idx):
    v = VECS[idx]
    rec = bytes.fromhex(v["input"])
    rounds = int.from_bytes(rec[:4], "big")
    adv, dig = gpu_witness(ctx, 17, rounds, rec, 1)
    assert dig.tobytes().hex() == v["output"]
    ref, _, _ = oracle.witness(17, rounds, rec, 1)
    assert np.array_equal(adv, ref)


def test_batch_bit_exact_and_mock_prover(ctx, oracle, zk):
    n = 26  # fills k = 17
    inputs = zk.synthetic_inputs(n)
    adv, dig = gpu_witness(ctx, 17, 12, inputs, n)
    ref, _, rdig = oracle.witness(17, 12, inputs, n)
    assert np.array_equal(dig, rdig)
    assert np.array_equal(adv, ref)
    rc, msg = oracle.mock_verify_mont(17, 12, n, adv)
    assert rc == 0, msg


def test_empty_and_ragged(ctx, oracle, zk):
    adv, _ = gpu_witness(ctx, 17, 12, b"", 0)
    assert not adv.any()
    # stale data in the scratch buffer must not leak into a smaller batch
    gpu_witness(ctx, 17, 12, zk.synthetic_inputs(5), 5)
    inputs = zk.synthetic_inputs(2, stream=3)
    adv, _ = gpu_witness(ctx, 17, 12, inputs, 2)
    ref, _, _ = oracle.witness(17, 12, inputs, 2)
    assert np.array_equal(adv, ref)


def test_rejections(ctx, zk):
    rec = bytearray(zk.synthetic_inputs(1))
    rec[212] = 2
    adv = np.empty((12, 1 << 17, 4), dtype=np.uint64)
    with pytest.raises(zk.ZkError) as e:
        ctx.witness_batch(17, 12, bytes(rec), 1, adv)
    assert e.value.code == -5
    with pytest.raises(zk.ZkError) as e:
        ctx.witness_batch(17, 12, zk.synthetic_inputs(27), 27, adv)
    assert e.value.code == -4
    with pytest.raises(zk.ZkError) as e:
        ctx.witness_batch(17, 11, zk.synthetic_inputs(1), 1, adv)
    assert e.value.code == -5


@pytest.mark.parametrize("n", [1, 3, 9, 26])
def test_slice_plans_bit_exact(ctx, oracle, zk, n):
    """The (compression, slice) work items are sized from the batch: every slice count the planner can
    pick at k = 17 (n = 1 -> 20 slices ... n = 26 -> many items per block) must emit the same cells."""
    inputs = zk.synthetic_inputs(n, stream=n)
    adv, dig = gpu_witness(ctx, 17, 12, inputs, n)
    ref, _, rdig = oracle.witness(17, 12, inputs, n)
    assert np.array_equal(dig, rdig)
    assert np.array_equal(adv, ref)


def test_more_items_than_resident_blocks(ctx, zk):
    """700 compressions at k = 22: more (compression, slice) items than one wave of resident blocks, so
    blocks walk several items.  The region of compression j must equal what the single-wave path (checked
    against the oracle above) writes when the batch is cut in two, and the digests must equal F."""
    import torch
    n, k, half = 700, 22, 350
    R = zk.rows_per_compression(12)
    inputs = zk.synthetic_inputs(n, stream=5)
    d_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).cuda()
    big = torch.empty((12, 1 << k, 4), dtype=torch.int64, device="cuda")
    dig = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.witness_batch_device(k, 12, d_in, n, big, dig)
    ctx.synchronize()
    part = torch.empty_like(big)
    for lo in (0, half):
        ctx.witness_batch_device(k, 12, d_in[lo * 213:], half, part, None)
        ctx.synchronize()
        assert torch.equal(big[:, lo * R:(lo + half) * R], part[:, :half * R]), lo
    assert not big[:, n * R:].any()
    got = dig.cpu().numpy().view(np.uint64)
    for j in (0, 1, 349, 350, 591, 592, 593, 699):
        want = np.frombuffer(zk.blake2f_compress(inputs[213 * j:213 * (j + 1)]), dtype=np.uint64)
        assert np.array_equal(got[j], want), j


def test_device_buffer_alignment(ctx, zk):
    """Cells are written with 256-bit stores: a device advice buffer that is not 32-byte aligned is
    rejected (ZK_E_INVALID), not written with misaligned stores."""
    import torch
    d_in = torch.frombuffer(bytearray(zk.synthetic_inputs(1)), dtype=torch.uint8).cuda()
    buf = torch.empty(12 * (1 << 17) * 32 + 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(zk.ZkError) as e:
        ctx.witness_batch_device(17, 12, d_in, 1, buf.data_ptr() + 16)
    assert e.value.code == -1
    ctx.witness_batch_device(17, 12, d_in, 1, buf.data_ptr() + 32)
    ctx.synchronize()


def test_full_size_properties(ctx, oracle, zk):
    """Config 3 (64 compressions, k = 19): digests equal the oracle's F on every record and the
    device-resident path equals the host path; the oracle's per-cell check runs on a sample."""
    import torch
    n, k = 64, zk.min_k(12, 64)
    assert k == 19
    inputs = zk.synthetic_inputs(n)
    d_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).cuda()
    d_adv = torch.empty((12, 1 << k, 4), dtype=torch.int64, device="cuda")
    d_dig = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    ctx.set_stream(stream.cuda_stream)
    ctx.witness_batch_device(k, 12, d_in, n, d_adv, d_dig)
    ctx.synchronize()
    ctx.set_stream(None)
    dig = d_dig.cpu().numpy().view(np.uint64)
    for i in range(n):
        rc, out = oracle.blake2f(inputs[213 * i:213 * (i + 1)])
        assert rc == 0 and out == dig[i].tobytes()
    adv = d_adv.cpu().numpy().view(np.uint64)
    ref, _, _ = oracle.witness(k, 12, inputs, n)
    assert np.array_equal(adv, ref)

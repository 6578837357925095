"""Error behaviour and threading of the C ABI on a device (every export returns a
status, nothing unwinds, contexts are independent)."""
import ctypes
import threading

import pytest

pytestmark = pytest.mark.gpu


def test_call_order_and_argument_errors(zk):
    ctx = zk.Context(0)
    seed = zk.REFERENCE_SEED
    inputs = zk.synthetic_inputs(2)
    for call in (lambda: ctx.keygen(12, 2), lambda: ctx.create_proof(inputs, 2, seed),
                 lambda: ctx.verify_proof(b"\x00" * 4064), lambda: ctx.mock_verify(inputs, 2)):
        with pytest.raises(zk.ZkError) as e:
            call()
        assert e.value.code == -6          # ZK_E_STATE: params / keys missing
    ctx.params_generate_substitute(17, seed)
    with pytest.raises(zk.ZkError) as e:
        ctx.create_proof(inputs, 2, seed)
    assert e.value.code == -6
    with pytest.raises(zk.ZkError) as e:
        ctx.keygen(12, 27)                  # 27 regions do not fit 2^17 rows
    assert e.value.code == -4               # ZK_E_ROWS
    ctx.keygen(12, 2)
    with pytest.raises(zk.ZkError) as e:
        ctx.create_proof(zk.synthetic_inputs(3), 3, seed)
    assert e.value.code == -1               # batch size differs from keygen
    with pytest.raises(zk.ZkError) as e:
        ctx.dist_init(b"\x00" * 128, 0, 2)  # joining a group after params exist
    assert e.value.code == -6
    # proof buffer too small: required size reported, then the call succeeds
    lib = ctx.lib
    ln = ctypes.c_uint64(16)
    small = ctypes.create_string_buffer(16)
    rc = lib.zk_create_proof(ctx.h, inputs, 2, bytes(seed), ctypes.cast(small, ctypes.c_void_p), ctypes.byref(ln))
    assert rc == -8 and ln.value == 4064
    proof = ctx.create_proof(inputs, 2, seed)
    assert len(proof) == 4064 and ctx.verify_proof(proof)
    ctx.close()


def test_bad_records_are_input_errors(zk):
    import torch
    ctx = zk.Context(0)
    seed = zk.REFERENCE_SEED
    ctx.params_generate_substitute(17, seed)
    ctx.keygen(12, 2)
    good = zk.synthetic_inputs(2)
    bad = bytearray(good)
    bad[212 + 213] = 2                      # final-block flag of the second record
    with pytest.raises(zk.ZkError) as e:
        ctx.create_proof(bytes(bad), 2, seed)
    assert e.value.code == -5               # ZK_E_INPUT, checked on the host
    d_bad = torch.frombuffer(bad, dtype=torch.uint8).cuda()
    with pytest.raises(zk.ZkError) as e:
        ctx.create_proof(d_bad, 2, seed, on_device=True)
    assert e.value.code == -5               # ZK_E_INPUT, raised by the witness kernel
    wrong_rounds = bytearray(good)
    wrong_rounds[3] = 11
    with pytest.raises(zk.ZkError) as e:
        ctx.create_proof(bytes(wrong_rounds), 2, seed)
    assert e.value.code == -5
    assert ctx.verify_proof(ctx.create_proof(good, 2, seed))   # the context recovers
    ctx.close()


def test_contexts_are_independent_across_threads(zk):
    """Four contexts on one device, four host threads (bench.py --streams): same bytes as alone."""
    seed = zk.REFERENCE_SEED
    inputs = zk.synthetic_inputs(3)
    ctxs = [zk.Context(0) for _ in range(4)]
    for c in ctxs:
        c.params_generate_substitute(17, seed)
        c.keygen(12, 3)
    alone = ctxs[0].create_proof(inputs, 3, seed)
    out = [[] for _ in ctxs]

    def work(i):
        if i % 2:
            ctxs[i].set_blocking_sync(True)
        for _ in range(3):
            out[i].append(ctxs[i].create_proof(inputs, 3, seed))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(ctxs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(p == alone for o in out for p in o)
    for c in ctxs:
        c.close()


def test_status_flag_is_per_call(zk):
    """A record rejected by a kernel in one call must not fail the next call (round-1 advisor finding): the
    asynchronous device-buffer witness entry latches a status bit when it meets f = 2; neither zk_mock_verify nor
    zk_create_proof with valid records may inherit it, and zk_mock_verify validates its own records."""
    import torch
    ctx = zk.Context(0)
    seed = zk.REFERENCE_SEED
    ctx.params_generate_substitute(17, seed)
    ctx.keygen(12, 2)
    good = zk.synthetic_inputs(2)
    bad = bytearray(good)
    bad[212] = 2
    # mock_verify: host-side EIP-152 validation, as create_proof
    with pytest.raises(zk.ZkError) as e:
        ctx.mock_verify(bytes(bad), 2)
    assert e.value.code == -5
    wrong_rounds = bytearray(good)
    wrong_rounds[3] = 11
    with pytest.raises(zk.ZkError) as e:
        ctx.mock_verify(bytes(wrong_rounds), 2)
    assert e.value.code == -5
    assert ctx.mock_verify(good, 2) is None
    # a stale status bit: the device-buffer witness call is asynchronous and is NOT followed by zk_ctx_synchronize
    d_bad = torch.frombuffer(bad, dtype=torch.uint8).cuda()
    d_adv = torch.empty((12, 1 << 17, 4), dtype=torch.int64, device="cuda")
    ctx.witness_batch_device(17, 12, d_bad, 2, d_adv)
    torch.cuda.synchronize()
    assert ctx.mock_verify(good, 2) is None                       # clears and owns the flag
    ctx.witness_batch_device(17, 12, d_bad, 2, d_adv)
    torch.cuda.synchronize()
    assert ctx.verify_proof(ctx.create_proof(good, 2, seed))      # starts from a clean flag
    # absurd batch sizes are refused before any size is multiplied out (no 64-bit wrap)
    with pytest.raises(zk.ZkError) as e:
        ctx.witness_batch_device(17, 12, d_bad, (1 << 64) // 4996 + 1, d_adv)
    assert e.value.code == -4
    ctx.close()

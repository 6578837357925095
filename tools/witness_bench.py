"""K1 alone: the witness kernel over device-resident EIP-152 records at several batch sizes, timed with
the context's CUDA events, next to a plain memset of the same number of bytes (store-only HBM rate)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import zk_odst_b200 as zk

only = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ctx = zk.Context(0)
R = zk.rows_per_compression(12)
out = {}
for n in ([only] if only else [64, 256, 1024]):
    k = zk.min_k(12, n)
    d_in = torch.frombuffer(bytearray(zk.synthetic_inputs(n)), dtype=torch.uint8).cuda()
    adv = torch.empty((12, 1 << k, 4), dtype=torch.int64, device="cuda")
    dig = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.enable_timing(True)
    best = 1e9
    for it in range(3 if only else 12):
        ctx.witness_batch_device(k, 12, d_in, n, adv, dig)
        ctx.synchronize()
        ms = ctx.timing_report()["witness"][0]
        if it >= 2:
            best = min(best, ms)
    ctx.enable_timing(False)
    nbytes = n * R * 12 * 32
    flat = adv.view(-1)[: nbytes // 8]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mbest = 1e9
    for it in range(6):
        e0.record()
        flat.zero_()
        e1.record()
        torch.cuda.synchronize()
        mbest = min(mbest, e0.elapsed_time(e1))
    out[n] = {"k": k, "bytes": nbytes, "witness_ms": best, "witness_gbs": nbytes / best / 1e6,
              "memset_ms": mbest, "memset_gbs": nbytes / mbest / 1e6}
print(json.dumps(out))

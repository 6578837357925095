"""K4 alone: size-2^log_n transforms over device-resident data, timed with the context's CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
import zk_odst_b200 as zk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = zk.Context(0)
n = 1 << log_n
d = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda")
ctx.ntt(d, log_n, inverse=False, on_device=True)
ctx.synchronize()
ctx.enable_timing(True)
ctx.timing_report()
for _ in range(reps):
    ctx.ntt(d, log_n, inverse=False, on_device=True)
ctx.synchronize()
ms = ctx.timing_report()["ntt"][0] / reps
mults = (n // 2) * log_n
print(json.dumps({"log_n": log_n, "ms": ms, "gmul_per_s": mults / ms / 1e6, "tmac_per_s": mults * 136 / ms / 1e9,
                  "gbs": 64 * n / ms / 1e6}))

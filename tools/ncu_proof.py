"""One proof of configs[2] (k = 19, 64 compressions) bracketed by cudaProfilerStart/Stop: the workload of
the `ncu --set full --profile-from-start off` captures summarised under profiles/."""
import sys

import torch

sys.path.insert(0, ".")
import zk_odst_b200 as zk

k, n = 19, 64
ctx = zk.Context(0)
ctx.params_generate_substitute(k, zk.REFERENCE_SEED)
ctx.keygen(12, n)
inputs = zk.synthetic_inputs(n)
ctx.create_proof(inputs, n, zk.REFERENCE_SEED)
torch.cuda.synchronize()
torch.cuda.profiler.start()
proof = ctx.create_proof(inputs, n, zk.REFERENCE_SEED)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(len(proof))

"""One small proof (k = 17, 2 compressions) plus a verify and a row check: the workload run under
compute-sanitizer (memcheck / racecheck) in tools/gpu_job*.sh."""
import sys

sys.path.insert(0, ".")
import zk_odst_b200 as zk

ctx = zk.Context(0)
k, n = 17, 2
ctx.params_generate_substitute(k, zk.REFERENCE_SEED)
ctx.keygen(12, n)
inputs = zk.synthetic_inputs(n)
proof = ctx.create_proof(inputs, n, zk.REFERENCE_SEED)
assert ctx.verify_proof(proof)
assert ctx.mock_verify(inputs, n) is None
print("proof bytes", len(proof))

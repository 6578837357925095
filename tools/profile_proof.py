"""Per-phase device time of one proof (CUDA events per kernel class) next to wall clock."""
import json
import sys
import time

sys.path.insert(0, ".")
import zk_odst_b200 as zk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 19
ncomp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = zk.Context(0)
t = time.time(); ctx.params_generate_substitute(k, zk.REFERENCE_SEED); t_params = time.time() - t
t = time.time(); ctx.keygen(12, ncomp); t_keygen = time.time() - t
inputs = zk.synthetic_inputs(ncomp)
ctx.create_proof(inputs, ncomp, zk.REFERENCE_SEED)
ctx.enable_timing(True)
ctx.timing_report()
l0 = ctx.launch_count()
t = time.time()
for _ in range(reps):
    proof = ctx.create_proof(inputs, ncomp, zk.REFERENCE_SEED)
wall = (time.time() - t) / reps
rep = ctx.timing_report()
print(json.dumps({"k": k, "n_compressions": ncomp, "params_s": t_params, "keygen_s": t_keygen,
                  "proof_wall_ms": wall * 1e3, "proof_bytes": len(proof),
                  "launches_per_proof": (ctx.launch_count() - l0) / reps,
                  "device_ms_per_proof": {n: (ms / reps, cnt / reps) for n, (ms, cnt) in rep.items() if cnt}}))

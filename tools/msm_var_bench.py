"""Times the stand-alone variable-base MSM (zk_msm_vesta / msm.cu) on device-resident operands at 2^12, 2^14 and 2^16
points (profiles/r02_msm_variable_base.json).  Uses the oracle only to obtain base points."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import zk_odst_b200 as zk
ctx = zk.Context(0)
ctx.params_generate_substitute(17, zk.REFERENCE_SEED)
sys.path.insert(0, "tests")
import oracle_lib
o = oracle_lib.load()
op = oracle_lib.OracleProver(o, k=16, seed=zk.REFERENCE_SEED)
g = op.points(0, 1 << 16)
rnd = np.random.RandomState(3)
sc = rnd.randint(0, 1 << 62, size=(1 << 16, 4), dtype=np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda(); d_g = torch.from_numpy(g.view(np.int64)).cuda()
out = np.zeros(8, dtype=np.uint64)
res = {}
for n in (1 << 12, 1 << 14, 1 << 16):
    ctx.msm(d_sc, d_g, n, out, on_device=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): ctx.msm(d_sc, d_g, n, out, on_device=True)
    torch.cuda.synchronize(); res[n] = (time.perf_counter() - t0) / 10 * 1e3
print(json.dumps(res))

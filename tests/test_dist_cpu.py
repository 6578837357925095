"""Host logic of the MSM split on CPU, world_size 2 over gloo: every rank sums
the terms of its zk_dist_range slice, the partial points are all-gathered and added — the result
must equal the unsplit MSM bit for bit.  The arithmetic here is the oracle's (this is a test of
the partition / exchange / combine structure that dist.cu implements over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SEED = bytes(range(1, 17))
K = 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_path):
    import torch
    import torch.distributed as dist
    import oracle_lib
    import zk_odst_b200 as zk
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    oracle = oracle_lib.load()
    p = oracle_lib.OracleProver(oracle, k=K, seed=SEED)
    bases = p.points(0, 1 << K)[:n]
    p.close()
    scalars = oracle_lib.random_fields(oracle, SEED, n)
    lo, hi = zk.dist_range(n, rank, world)
    if hi > lo:
        part = oracle_lib.msm(oracle, np.ascontiguousarray(scalars[lo:hi]), np.ascontiguousarray(bases[lo:hi]))
    else:
        part = np.zeros(8, dtype=np.uint64)
    mine = torch.from_numpy(part.view(np.int64).copy())
    gathered = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = np.stack([g.numpy().view(np.uint64) for g in gathered])
    one = oracle_lib.Oracle._limbs(oracle.field_op(0, 4, 1)[1])
    ones = np.tile(np.array(one, dtype=np.uint64), (world, 1))
    total = oracle_lib.msm(oracle, np.ascontiguousarray(ones), np.ascontiguousarray(parts.reshape(world, 8)))
    whole = oracle_lib.msm(oracle, scalars, np.ascontiguousarray(bases))
    ok = np.array_equal(total, whole)
    with open(out_path + ".%d" % rank, "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 2, 777, 1 << K])
def test_range_split_msm_world2(tmp_path, n):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    for r in range(2):
        assert open(out + ".%d" % r).read() == "ok"


def test_ranges_partition_the_points():
    import zk_odst_b200 as zk
    for n in (0, 1, 5, 1 << 19, (1 << 23) + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = zk.dist_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n


def test_column_blocks_cover_the_witness_columns():
    """zk_dist_column_block: every one of the 19 witness columns is transformed by exactly one rank and
    rank r's block starts at slot r * per_rank of the padded array (the in-place all-gather layout)."""
    import zk_odst_b200 as zk
    for world in (1, 2, 3, 4, 5, 8, 19, 24):
        owner = [None] * 19
        for r in range(world):
            lo, hi, per = zk.dist_column_block(r, world)
            assert per == -(-19 // world) and hi - lo <= per
            assert lo == min(r * per, 19)
            for s in range(lo, hi):
                assert owner[s] is None
                owner[s] = r
        assert all(o is not None for o in owner), world


def _column_worker(rank, world, port, out_path):
    """The exchange prover.cu performs with ncclAllGather, over gloo: every rank fills only its own block
    of the padded slot array, the blocks are all-gathered in place, every rank ends with all columns."""
    import torch
    import torch.distributed as dist
    import zk_odst_b200 as zk
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 64
    lo, hi, per = zk.dist_column_block(rank, world)
    slots = torch.full((per * world, n), -1, dtype=torch.int64)
    for s in range(lo, hi):  # "transform" of column s: anything that depends on (s, row) only
        slots[s] = torch.arange(n, dtype=torch.int64) * 1000 + s
    mine = slots[rank * per:(rank + 1) * per].clone()
    dist.all_gather(list(slots.view(world, per, n).unbind(0)), mine)
    want = torch.arange(n, dtype=torch.int64)[None, :] * 1000 + torch.arange(19, dtype=torch.int64)[:, None]
    ok = torch.equal(slots[:19], want)
    # quotient rows: rank r evaluates rows [r * rows, (r + 1) * rows) of the 3n-row domain
    en = 3 * n
    rows = en // world
    h = torch.full((en,), -1, dtype=torch.int64)
    h[rank * rows:(rank + 1) * rows] = torch.arange(rank * rows, (rank + 1) * rows) * 7
    dist.all_gather(list(h.view(world, rows).unbind(0)), h[rank * rows:(rank + 1) * rows].clone())
    ok = ok and torch.equal(h, torch.arange(en) * 7)
    with open(out_path + ".%d" % rank, "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_column_sharded_transforms_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "cols")
    mp.spawn(_column_worker, args=(2, port, out), nprocs=2, join=True)
    for r in range(2):
        assert open(out + ".%d" % r).read() == "ok"


def test_quotient_row_segments_cover_the_rotations():
    """zk_dist_quotient_rows: the segments a rank receives are exactly the rows its share of the quotient
    reads — every row of its range and the rotations -1, +1, -6 inside the row's own coset — and the
    ranks' ranges tile the 3n-row domain."""
    import zk_odst_b200 as zk
    for n in (16, 64, 1 << 10):
        for world in (1, 2, 3, 4, 6, 8):
            if (3 * n) % world:
                continue
            prev = 0
            for r in range(world):
                (lo, hi), segs = zk.dist_quotient_rows(n, r, world)
                assert lo == prev and hi - lo == 3 * n // world
                prev = hi
                need = set()
                for i in range(lo, hi):
                    c, row = divmod(i, n)
                    for rot in (0, -1, 1, -6):
                        need.add(c * n + (row + rot) % n)
                got = set()
                last_end = -1
                for start, length in segs:
                    assert length > 0 and start > last_end  # sorted, disjoint, not even adjacent
                    last_end = start + length
                    got.update(range(start, start + length))
                assert got == need, (n, world, r)
            assert prev == 3 * n


# ---- the range-sharded steps of a group: grand product and division by (X - p) ---------------------
P_MOD = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001


def _range_worker(rank, world, port, n, out_path):
    """What prover.cu does per rank in a group, on Python integers: scan the own range from a neutral start, exchange
    one field element per rank (all_gather over gloo), correct the range.  Compared with the serial recurrences."""
    import random
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rnd = random.Random(5)
    ratio = [rnd.randrange(1, P_MOD) for _ in range(n)]
    coeffs = [rnd.randrange(P_MOD) for _ in range(n)]
    pt, init = rnd.randrange(P_MOD), rnd.randrange(1, P_MOD)
    cnt = n // world
    lo, hi = rank * cnt, (rank + 1) * cnt

    def gather(v):  # one field element per rank
        t = torch.tensor([(v >> (62 * i)) & ((1 << 62) - 1) for i in range(5)], dtype=torch.int64)
        out = [torch.zeros(5, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(out, t)
        return [sum(int(o[i]) << (62 * i) for i in range(5)) for o in out]

    # grand product (prover.cu grand_product): z[0] = init, z[i + 1] = z[i] * ratio[i]
    z_local, acc = [], 1
    for i in range(lo, hi):
        z_local.append(acc)
        acc = acc * ratio[i] % P_MOD
    totals = gather(acc)
    prefix = init
    for q in range(rank):
        prefix = prefix * totals[q] % P_MOD
    z_mine = [v * prefix % P_MOD for v in z_local]
    z_ref, acc = [], init
    for i in range(n):
        z_ref.append(acc)
        acc = acc * ratio[i] % P_MOD
    ok = z_mine == z_ref[lo:hi]

    # division by (X - pt) (multiopen): y_0 = 0, y_{j+1} = y_j pt + c_{n-1-j}; quotient coefficient i = y_{n-1-i}
    scanned, y = [], 0
    for t in range(cnt):           # t = distance from the top of the range
        scanned.append(y)
        y = (y * pt + coeffs[hi - 1 - t]) % P_MOD
    ends = gather(y)
    p_cnt = pow(pt, cnt, P_MOD)
    carry = 0
    for q in range(world - 1, rank, -1):
        carry = (carry * p_cnt + ends[q]) % P_MOD
    quot_mine = {hi - 1 - t: (scanned[t] + carry * pow(pt, t, P_MOD)) % P_MOD for t in range(cnt)}
    quot_ref, y = [0] * n, 0
    for j in range(n):
        quot_ref[n - 1 - j] = y
        y = (y * pt + coeffs[n - 1 - j]) % P_MOD
    ok = ok and all(quot_mine[i] == quot_ref[i] for i in range(lo, hi))
    # (the quotient really is the division: q(X) (X - pt) + rem == c(X), checked at a point)
    if rank == 0:
        x = rnd.randrange(P_MOD)
        ev = lambda cs: sum(c * pow(x, i, P_MOD) for i, c in enumerate(cs)) % P_MOD
        ok = ok and (ev(quot_ref) * (x - pt) + y) % P_MOD == ev(coeffs)
    with open(out_path + ".%d" % rank, "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_range_sharded_scans(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "r")
    mp.spawn(_range_worker, args=(world, port, 64, out), nprocs=world, join=True)
    for r in range(world):
        assert open(out + ".%d" % r).read() == "ok", r

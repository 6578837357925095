#!/usr/bin/env python
"""bench.py — headline benchmark of the BLAKE2f proving path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle restatement, same config)

A "step" is one complete proof (witness -> commitments -> quotient -> multiopen -> IPA) of one
batch of synthetic EIP-152 records: BASELINE.json configs[2], 64 twelve-round compressions in
one circuit (k = 19).  With N GPUs every rank proves its own independent batch (configs[4]:
independent proof streams, weak scaling, no data-path collective); the same run then also proves ONE
configs[3]-shape circuit sharded over all N ranks (`strong_split`) and checks that a sharded proof has
the single-GPU bytes (`split_parity`).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definitions.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROUNDS = 12
N_COMPRESSIONS = 64
METRIC = "blake2f_12round_compressions_proved_per_sec"
UNIT = "compressions/s"
# SURVEY.md §8d accounting: a full-width MSM term = 16 bucket additions x 11 Fq mults x 136 MAC,
# a <= 32-bit advice term = 2 windows.
MAC_FULL, MAC_SMALL = 16 * 11 * 136, 2 * 11 * 136
EXECUTED_PER_ALGORITHMIC = (10 * 97) / (11 * 136.0)   # IMAD.WIDE the accumulation issues per algorithmic MAC
MAC_PER_FP_MUL = 136
ROWS_PER_COMPRESSION = 292 + 392 * ROUNDS
# dram__bytes_read.sum + dram__bytes_write.sum of one full-width accumulation launch (k = 19), from the
# `ncu --set full` capture named here (not measured in the run: DRAM counters need the profiler)
TRAFFIC_CAPTURE = {"bytes": None, "source": None}
_tc = os.path.join(ROOT, "profiles", "accumulate_traffic.json")
if os.path.exists(_tc):
    TRAFFIC_CAPTURE = json.load(open(_tc))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---- CPU arm: the oracle (halo2-0.3.0-equivalent restatement) ------------------------------------------
# The reference itself is Rust and does not compile (no cargo in the image), so this arm is the
# repo's C++ port of its algorithm: kind "port".  Phase labels follow benchmarking/src/constants.rs:1-3.
CPU_SAMPLE_K, CPU_SAMPLE_N = 17, 26   # bounded sample: the largest batch that fits the default k


def oracle_modules():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import zk_odst_b200 as zk
    return oracle_lib, zk


def cpu_timed_setup(oracle_lib, oracle, k, n_compressions, seed):
    """[Setup generation]: Params + keygen_vk + keygen_pk of the oracle; returns (prover, seconds)."""
    t0 = time.perf_counter()
    op = oracle_lib.OracleProver(oracle, k=k, seed=seed)
    t1 = time.perf_counter()
    op.keygen(ROUNDS, n_compressions)
    t2 = time.perf_counter()
    return op, {"params_s": t1 - t0, "keygen_s": t2 - t1}


def cpu_prove(op, inputs, n, seed):
    t0 = time.perf_counter()
    proof = op.create_proof(inputs, n, seed)
    dt = time.perf_counter() - t0
    synth_ms, rest_ms = op.last_proof_ms()
    return proof, dt, synth_ms * 1e-3


def workload_config(k, n, split=False, world=1):
    what = ("configs[3] shape: ONE proof of %d compressions, every MSM split by point range over %d "
            "GPUs (NCCL all-gather of partial points), the witness transforms sharded by column, the "
            "quotient by row, evaluations by coefficient range; the rest replicated" % (n, world)) if split else (
            "configs[2]: %d twelve-round BLAKE2f compressions in one circuit, "
            "full create_proof (Pasta/IPA)" % n)
    return {"workload": what,
            "k": k, "rounds": ROUNDS, "compressions_per_proof": n,
            "rows_per_compression": ROWS_PER_COMPRESSION,
            "rows_used_frac": round(n * ROWS_PER_COMPRESSION / float(1 << k), 3),
            "params": "substitute URS (zk_params_generate_substitute, reference seed)",
            "l2": "per proof the prover streams > 2 GB of column/coset data (advice cosets alone "
                  "12 x 3 x 2^k x 32 B = %d MB), far above the 126 MB L2; no explicit flush" %
                  (12 * 3 * (1 << k) * 32 >> 20)}


def run_reference(args, rank):
    """The reference arm: the CPU restatement proves configs[2] itself (k = 19, 64 compressions) with every
    host thread; then, at --gpus 1, the single-thread figure (the pinned reference has rayon off:
    blake2f-circuit/Cargo.toml:15) on the bounded k = 17 sample."""
    if rank != 0:
        return
    oracle_lib, zk = oracle_modules()
    oracle = oracle_lib.load()
    cores = oracle_lib.set_threads(oracle, 0)
    t_run = time.perf_counter()
    n = args.compressions
    k = 17
    while n * ROWS_PER_COMPRESSION > (1 << k) - 6:
        k += 1
    seed = zk.REFERENCE_SEED
    inputs = zk.synthetic_inputs(n)
    op, setup = cpu_timed_setup(oracle_lib, oracle, k, n, seed)
    prove_budget_s = 150.0
    steps, t_used, synth_s, proof = 0, 0.0, 0.0, None
    t_start = time.perf_counter()
    while steps < max(1, args.steps):
        proof, dt, sy = cpu_prove(op, inputs, n, seed)
        steps += 1
        t_used += dt
        synth_s += sy
        if time.perf_counter() - t_start + dt > prove_budget_s:
            break
    dt = t_used / steps
    t0 = time.perf_counter()
    rc, msg = op.verify(proof)
    verify_s = time.perf_counter() - t0
    assert rc == 0, "oracle rejected its own proof: " + msg
    op.close()
    val = n / dt
    phases = {"[Setup generation]": setup["params_s"] + setup["keygen_s"], "[Proof generation]": dt,
              "[Proof verification]": verify_s, "witness synthesis (inside proof generation)": synth_s / steps,
              "params_s": setup["params_s"], "keygen_s": setup["keygen_s"], "unit": "s"}
    one_thread = None
    elapsed = time.perf_counter() - t_run
    if args.gpus == 1 and not args.no_one_thread and elapsed < 170.0:
        oracle_lib.set_threads(oracle, 0)
        op1, _ = cpu_timed_setup(oracle_lib, oracle, CPU_SAMPLE_K, CPU_SAMPLE_N, seed)
        oracle_lib.set_threads(oracle, 1)
        _, dt1, sy1 = cpu_prove(op1, zk.synthetic_inputs(CPU_SAMPLE_N), CPU_SAMPLE_N, seed)
        oracle_lib.set_threads(oracle, 0)
        op1.close()
        one_thread = {"value": CPU_SAMPLE_N / dt1, "unit": UNIT, "cores": 1, "proof_generation_s": dt1,
                      "witness_synthesis_s": sy1,
                      "sample": "one proof of %d compressions at k=%d on ONE thread (the pinned reference is "
                                "single-threaded: rayon off); a k=19 proof on one thread would take ~%d s" %
                                (CPU_SAMPLE_N, CPU_SAMPLE_K, int(dt * cores * 0.8))}
    cfg = workload_config(k, n)
    sample = ("%d proof(s) of %d compressions at k=%d (create_proof incl. witness synthesis; setup timed "
              "separately), %.1f s each, C++ oracle restating halo2_proofs 0.3.0, std::thread x %d" %
              (steps, n, k, dt, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic", "config": cfg, "proofs_per_sec": 1.0 / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "phases": phases, "one_thread": one_thread},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "proof_sha256": hashlib.sha256(proof).hexdigest(),
    }
    print(json.dumps(line))


# ---- sharded single proof: parity and strong scaling (world > 1) ---------------------------------------
def split_section(args, zk, torch, dist, rank, world, local_rank, barrier, max_over_ranks):
    """(split_parity, strong_split): one k = 17 proof sharded over all ranks must have the single-GPU bytes;
    then ONE configs[3]-shape circuit (k = 21 on 2 GPUs, k = 23 on 4 / 8) is proved by the whole group and by
    rank 0 alone."""
    seed = zk.REFERENCE_SEED
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(zk.dist_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    g = zk.Context(local_rank)
    g.set_blocking_sync(world * 2 > (os.cpu_count() or 1))
    g.dist_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def digest_all_equal(blob):
        h = torch.frombuffer(bytearray(hashlib.sha256(blob).digest()), dtype=torch.uint8).cuda()
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        return all(bool(torch.equal(x, hs[0])) for x in hs)

    # -- parity: k = 17, 3 compressions, every rank's bytes == rank 0's == a single GPU's
    kp, np_ = 17, 3
    rec = zk.synthetic_inputs(np_)
    g.params_generate_substitute(kp, seed)
    g.keygen(ROUNDS, np_)
    mine = g.vk_bytes() + g.create_proof(rec, np_, seed)
    same = digest_all_equal(mine)
    single_equal = None
    if rank == 0:
        s = zk.Context(local_rank)
        s.params_generate_substitute(kp, seed)
        s.keygen(ROUNDS, np_)
        single_equal = (s.vk_bytes() + s.create_proof(rec, np_, seed)) == mine
        s.close()
    parity = {"k": kp, "compressions": np_, "ranks_agree": same, "equals_single_gpu": single_equal,
              "ok": bool(same and single_equal)} if rank == 0 else None
    barrier()

    # -- strong scaling of one proof
    n = args.split_compressions or (256 if world == 2 else 1024)
    k = zk.min_k(ROUNDS, n)
    rec = zk.synthetic_inputs(n)
    t0 = time.perf_counter()
    g.params_generate_substitute(k, seed)
    g.keygen(ROUNDS, n)
    setup_s = time.perf_counter() - t0
    d_in = torch.frombuffer(bytearray(rec), dtype=torch.uint8).cuda()
    proof = g.create_proof(d_in, n, seed, on_device=True)   # warm-up (allocates the workspace)
    g.enable_timing(True)
    g.timing_report()
    steps = 3
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        proof = g.create_proof(d_in, n, seed, on_device=True)
    torch.cuda.synchronize()
    ms_split = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    rep = g.timing_report()
    g.enable_timing(False)
    agree = digest_all_equal(proof)
    classes = {name: ms / steps for name, (ms, cnt) in rep.items() if cnt}
    g.close()
    barrier()
    ms_single, single_same = None, None
    if rank == 0 and not args.no_split_single:
        s = zk.Context(local_rank)
        s.params_generate_substitute(k, seed)
        s.keygen(ROUNDS, n)
        p1 = s.create_proof(d_in, n, seed, on_device=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            p1 = s.create_proof(d_in, n, seed, on_device=True)
        ms_single = (time.perf_counter() - t0) / steps * 1e3
        single_same = p1 == proof
        assert s.verify_proof(proof), s.last_error()
        s.close()
    barrier()
    if rank != 0:
        return None, None
    nrows = 1 << k
    per_rank_slots = -(-19 // world)
    sharded = classes.get("msm", 0.0) + classes.get("ntt", 0.0) + classes.get("quotient", 0.0) + \
        classes.get("collapse", 0.0) + classes.get("witness", 0.0)
    strong = {
        "workload": workload_config(k, n, True, world)["workload"], "k": k, "compressions_per_proof": n,
        "ms_per_proof": ms_split, "ms_single_gpu": ms_single,
        "speedup": ms_single / ms_split if ms_single else None,
        "efficiency": ms_single / ms_split / world if ms_single else None,
        "compressions_per_sec": n / (ms_split * 1e-3),
        "ranks_agree": agree, "equals_single_gpu": single_same, "setup_s": setup_s, "steps": steps,
        "rank0_class_ms": classes,
        # what rank 0's timed kernel classes do not cover: the lookup permutation and the random polynomials
        # (replicated), rank 0's share of the grand products, evaluations and multiopen (sharded by range), the
        # replicated vector updates of the folding rounds, host transcript round trips and the NCCL exchanges
        "replicated_ms": ms_split - sharded,
        # bytes a rank receives per proof over NCCL (rank 0's figures; 17 of the 19 column slots are transformed)
        "allgather_bytes": {
            # its own coefficient range of the columns other ranks transformed (grouped send/recv)
            "coefficient_ranges_received_per_rank": (17 - min(per_rank_slots, 10)) * (nrows // world) * 32,
            # its quotient rows (+ rotation halo) of the coset columns other ranks transformed (grouped send/recv)
            "coset_row_segments_received_per_rank": int(17 * (3 * nrows // world + 21) * 32 * (world - 1) / world),
            # rows of the quotient cosets it inverse-transforms, then its coefficient range of the others
            "h_rows_and_ranges_received_per_rank": int(-(-3 // world) * nrows * 32 * (world - 1) / world) +
                                                   (3 - -(-3 // world)) * (nrows // world) * 32,
            # all-gathers: grand-product rows (5 columns), p' before the folding rounds, folded generators
            "allgathered_rows_per_rank": 6 * nrows * 32 + (nrows // 32) * 128 * world,
            "partial_points_per_msm_batch": 128 * 16 * world},
    }
    return parity, strong


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU oracle legs (cpu_baseline and the parity gates that need it)")
    ap.add_argument("--no-one-thread", action="store_true", help="reference arm: skip the single-thread figure")
    ap.add_argument("--compressions", type=int, default=N_COMPRESSIONS)
    ap.add_argument("--streams", type=int, default=8,
                    help="concurrent proof streams per GPU (one context and host thread each)")
    ap.add_argument("--blocking-sync", type=int, default=-1,
                    help="1: host threads sleep while waiting for the device, 0: spin, -1: sleep only "
                         "when ranks x streams would oversubscribe the cores")
    ap.add_argument("--msm-split", action="store_true",
                    help="configs[3] as the headline: ONE proof stream, every MSM split by point range across the "
                         "ranks (NCCL all-gather of partial points), transforms by column, quotient by row; "
                         "strong scaling")
    ap.add_argument("--no-split-section", action="store_true",
                    help="N > 1: skip the split_parity / strong_split sub-records")
    ap.add_argument("--no-split-single", action="store_true",
                    help="strong_split: skip the single-GPU proof of the same circuit (no efficiency)")
    ap.add_argument("--split-compressions", type=int, default=0,
                    help="strong_split: compressions per proof (default 256 on 2 GPUs, 1024 on 4 / 8)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import zk_odst_b200 as zk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.compressions
    k = zk.min_k(ROUNDS, n)
    nrows = 1 << k
    seed = zk.REFERENCE_SEED
    split = args.msm_split and world > 1
    S = 1 if split else max(1, args.streams)   # concurrent proof streams per GPU (one context each)
    # a context holds the window tables of g / g_lagrange, the keys' coset forms and a proof's workspace:
    # about 9.5 KB per row (k = 23: ~80 GB); larger circuits get fewer concurrent contexts instead of an OOM
    S = max(1, min(S, int(0.8 * torch.cuda.get_device_properties(local_rank).total_memory // (nrows * 9728))))
    inputs = zk.synthetic_inputs(n, stream=0 if split else rank)  # independent batch per rank
    ev_stream = torch.cuda.Stream()  # carries only the timing events
    torch.cuda.set_stream(ev_stream)
    ctxs = [zk.Context(local_rank) for _ in range(S)]
    ctx = ctxs[0]
    # host threads of this box: world x S; they wait on blocking events instead of spinning when the
    # cores would be oversubscribed
    blocking = args.blocking_sync == 1 or (args.blocking_sync < 0 and world * S * 2 > (os.cpu_count() or 1))
    for c in ctxs:
        c.set_blocking_sync(blocking)
    if split:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(zk.dist_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.dist_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    for c in ctxs:
        c.params_generate_substitute(k, seed)
        c.keygen(ROUNDS, n)
    d_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).cuda()
    h_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_streams(total, src, on_device):
        """`total` proofs shared by the S contexts (the first total % S take one more), each context on its
        own CUDA stream from its own host thread (create_proof returns after the proof bytes reached the
        host).  Returns the last proof of every context that proved one."""
        out = [None] * S
        counts = [total // S + (1 if i < total % S else 0) for i in range(S)]

        def work(i):
            for _ in range(counts[i]):
                out[i] = ctxs[i].create_proof(src, n, seed, on_device=on_device)
        if S == 1:
            work(0)
        else:
            ts = [threading.Thread(target=work, args=(i,)) for i in range(S)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return [p for p in out if p is not None]

    def timed(total, src, on_device):
        """Device time (CUDA events around the region; every stream is idle at both ends) per proof."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ev_stream)
        proofs = run_streams(total, src, on_device)
        e1.record(ev_stream)
        torch.cuda.synchronize()
        dt_host = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        assert abs(dt - dt_host) < 0.05 * dt_host + 0.01, (dt, dt_host)
        return dt / total, proofs

    # ---- warm-up, then one context alone with per-kernel-class CUDA-event timing ------------------
    run_streams(args.warmup * S, d_in, True)   # W warm-up proofs on every context
    barrier()
    single_steps = max(3, min(args.steps, 5))
    ctx.enable_timing(True)
    ctx.timing_report()
    t0 = time.perf_counter()
    for _ in range(single_steps):
        proof = ctx.create_proof(d_in, n, seed, on_device=True)
    single_ms = (time.perf_counter() - t0) / single_steps * 1e3
    report = ctx.timing_report()
    ctx.enable_timing(False)

    # ---- value: inputs resident in HBM, S concurrent proof streams ------------------------------------
    # exactly --steps proofs are timed, shared by the S contexts
    steps_done = max(1, args.steps)
    launches0 = sum(c.launch_count() for c in ctxs)
    with ClockSampler(local_rank) as clocks:
        sec_per_proof, proofs = timed(steps_done, d_in, True)
    launches = sum(c.launch_count() for c in ctxs) - launches0
    assert all(p == proof for p in proofs), "streams disagree on the proof bytes"
    ms_per_step = max_over_ranks(sec_per_proof * 1e3)
    jobs = 1 if split else world  # split: all ranks work on the same proof
    value = jobs * n / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) records in, proof bytes out, through the C-ABI call ---------------------
    run_streams(S, h_in, False)
    e2e_steps = max(2 * S, min(steps_done, 5 * S))
    e2e_sec, proofs_e2e = timed(e2e_steps, h_in, False)
    e2e_ms = max_over_ranks(e2e_sec * 1e3)
    e2e_val = jobs * n / (e2e_ms * 1e-3)
    assert all(p == proof for p in proofs_e2e), "host-input and device-input proofs differ"

    # ---- the other two calls of the reference's sequence, for the record (one context, rank 0) -----
    verify_ms = mock_ms = None
    if rank == 0 and not split:   # (a group context's verifier MSM is a collective: not from rank 0 alone)
        assert ctx.verify_proof(proof), ctx.last_error()
        t0 = time.perf_counter()
        for _ in range(5):
            ok = ctx.verify_proof(proof)
        verify_ms = (time.perf_counter() - t0) / 5 * 1e3
        assert ok
        assert ctx.mock_verify(inputs, n) is None
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.mock_verify(inputs, n)
        mock_ms = (time.perf_counter() - t0) / 5 * 1e3
    barrier()

    # ---- the sharded single proof, visible to the driver's scaling run --------------------------------
    split_parity = strong_split = None
    if world > 1 and not split and not args.no_split_section:
        for c in ctxs[1:]:
            c.close()   # the configs[3] circuit needs the memory of the extra proof streams
        ctxs = ctxs[:1]
        split_parity, strong_split = split_section(args, zk, torch, dist, rank, world, local_rank, barrier,
                                                   max_over_ranks)

    if rank == 0:
        hbm_peak, hbm_kind = load_peaks()
        int_peak = ctx.bench_int_pipe(1, 20000)  # mad.wide.u32 instructions/s = 32x32->64 MAC/s
        # carry-chained form (mad.lo.cc + madc.hi.cc, two instructions per MAC): what multi-limb arithmetic can issue
        chain_peak = ctx.bench_int_pipe(2, 20000) / 2.0
        steps = single_steps
        per = {name: (ms / steps, cnt / steps) for name, (ms, cnt) in report.items() if cnt}
        R = ROWS_PER_COMPRESSION
        # dominant kernel: MSM bucket accumulation
        # terms the accumulate launches process per proof: 13 full-width commitments (2 lookup, 5 grand
        # products, random, 3 h pieces, q', s), 12 advice columns (<= 32-bit cells), and the IPA: `fold` rounds
        # of n terms on the original generators, then k - fold rounds of n / 2^fold terms on the folded ones
        # (halo2's schedule would be 2n MSM terms plus n generator-folding scalar multiplications)
        # (a group of 2^e ranks keeps e more rounds on the original generators: prover_state.h ipa_fold_rounds)
        fold = 5
        if split:
            w = world
            while w > 1 and fold < 8:
                fold, w = fold + 1, w >> 1
        if not (k > fold + 8 and (not split or (1 << fold) % world == 0)):
            fold = k
        ipa_terms = fold + (k - fold) / float(1 << fold)
        full_terms = 13 + ipa_terms
        # (MSM-split group: the timed launches are rank 0's, which process 1 / world of every MSM's terms)
        share = world if split else 1
        mac_per_proof = nrows * (full_terms * MAC_FULL + 12 * MAC_SMALL) / share
        acc_ms, acc_launches = per.get("msm_accumulate", (0.0, 0.0))
        msm_ms, _ = per.get("msm", (0.0, 0.0))
        achieved = mac_per_proof / (acc_ms * 1e-3) / 1e12 if acc_ms else None
        roofline = {
            "kernel": "fixed_accumulate_kernel (+ heavy-bucket kernels)", "bound": "int_pipe",
            "achieved": achieved, "peak": int_peak / 1e12, "unit": "TMAC/s",
            "frac": achieved / (int_peak / 1e12) if achieved else None,
            "traffic": TRAFFIC_CAPTURE["bytes"] if k == 19 else None,
            "traffic_source": TRAFFIC_CAPTURE["source"],
            "peak_source": "zk_bench_int_pipe mode 1 (mad.wide.u32) measured in this run; "
                           "MEASURED_PEAKS.json has no integer peak",
            # for information: the MACs the kernel EXECUTES (10 products of 97 carry-chained IMAD.WIDE per mixed
            # addition, tools/gen_mont_ptx.py, against the 11 x 136 of the algorithmic count above) against the
            # carry-chained MAC rate (mode 2) — the rate a multi-limb product can actually issue at; `frac` above
            # stays algorithmic MACs against the plain mad.wide peak
            "peak_carry_chained_tmacs": chain_peak / 1e12,
            "executed_tmacs": achieved * EXECUTED_PER_ALGORITHMIC if achieved else None,
            "frac_of_carry_chained_peak": (achieved * EXECUTED_PER_ALGORITHMIC / (chain_peak / 1e12)
                                           if achieved else None),
            "launches_per_proof": acc_launches, "avg_launch_ms": acc_ms / acc_launches if acc_launches else None,
            "algorithmic_mac_per_proof": mac_per_proof,
            "accounting": "SURVEY.md 8d per-term figures (23936 MAC full-width, 2992 MAC advice) x the "
                          "terms the launches process: %.2f n full-width + 12 n advice%s"
                          % (full_terms, " (rank 0's 1/%d share of each MSM)" % world if split else ""),
        }
        ntt_ms, ntt_launches = per.get("ntt", (0.0, 0.0))
        # 17 inverse transforms to coefficients (the 19 witness columns less the two never-queried advice
        # columns), 17 columns x 3 coset transforms, 3 inverse transforms of h's coset values: all of size n
        # (the quotient lives on three cosets of the n-th roots)
        # (MSM-split group: rank 0 transforms ceil(19 / world) of the columns and evaluates 1 / world of the
        # quotient rows; the 3 transforms of h stay replicated)
        ext = 3 * nrows // share
        blk_hi = min(-(-19 // share), 19)   # rank 0's block of the 19 column slots is [0, blk_hi)
        own_cols = min(blk_hi, 10) + max(0, blk_hi - 12)   # slots 10, 11 (never-queried advice columns) are skipped
        n_transforms = own_cols + own_cols * 3 + 3
        ntt_bytes = n_transforms * 64 * nrows
        ntt_mac = n_transforms * (nrows // 2) * k * MAC_PER_FP_MUL
        wit_ms, _ = per.get("witness", (0.0, 0.0))
        q_ms, _ = per.get("quotient", (0.0, 0.0))
        others = {
            "witness": {"bound": "hbm", "ms_per_proof": wit_ms,
                        "achieved_gbs": n * (R * 12 * 32 + 213) / (wit_ms * 1e-3) / 1e9 if wit_ms else None,
                        "peak_gbs": hbm_peak, "peak_source": hbm_kind},
            "ntt": {"bound": "int_pipe", "ms_per_proof": ntt_ms, "transforms_per_proof": n_transforms, "batched_calls_per_proof": ntt_launches,
                    "achieved_gbs": ntt_bytes / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None,
                    "achieved_tmacs": ntt_mac / (ntt_ms * 1e-3) / 1e12 if ntt_ms else None},
            "quotient": {"bound": "int_pipe", "ms_per_proof": q_ms,
                         "achieved_gbs": (61 * 32 * ext) / (q_ms * 1e-3) / 1e9 if q_ms else None},
            "msm_total_ms_per_proof": msm_ms,
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps_done,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if split else "weak", "vs_baseline": None,
            "dtype": "u64 (4x64-bit Montgomery limbs)",
            "data": "synthetic", "config": workload_config(k, n, split, world),
            "proofs_per_sec": jobs / (ms_per_step * 1e-3), "proof_bytes": len(proof),
            "streams_per_gpu": S, "single_stream_ms_per_proof": single_ms, "blocking_sync": bool(blocking),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": len(inputs),
                    "d2h_bytes_per_step": len(proof),
                    "note": "h2d = the 213-byte EIP-152 records from pinned host memory, d2h = the "
                            "proof bytes; the prover's random polynomials are generated on the "
                            "device (XorShift jump-ahead), challenges/commitments move as <1 KB "
                            "transcript round trips inside the call"},
            "gpu_launches": int(launches), "roofline": roofline, "kernels": others,
            "verify_proof_ms": verify_ms, "mock_verify_ms": mock_ms,
            "proof_sha256": hashlib.sha256(proof).hexdigest(),
        }
        if split_parity is not None:
            line["split_parity"] = split_parity
            line["strong_split"] = strong_split
        if not args.no_cpu_baseline:
            # The CPU oracle, after the timed regions: (1) the bounded k = 17 sample is the cpu_baseline AND a
            # parity gate — the GPU proves the same records with the same seed and must produce the same bytes
            # ("GPU proof bytes == oracle proof bytes"); (2) the bench's own k = 19 proof is
            # checked by the ORACLE's verify_proof under the oracle's own keygen_vk of the same circuit.
            oracle_lib, _ = oracle_modules()
            oracle = oracle_lib.load()
            cores = oracle_lib.set_threads(oracle, 0)
            cin = zk.synthetic_inputs(CPU_SAMPLE_N)
            op, csetup = cpu_timed_setup(oracle_lib, oracle, CPU_SAMPLE_K, CPU_SAMPLE_N, seed)
            cproof, cdt, csynth = cpu_prove(op, cin, CPU_SAMPLE_N, seed)
            t0 = time.perf_counter()
            assert op.verify(cproof)[0] == 0
            cverify = time.perf_counter() - t0
            small = zk.Context(local_rank)
            small.params_generate_substitute(CPU_SAMPLE_K, seed)
            small.keygen(ROUNDS, CPU_SAMPLE_N)
            gproof = small.create_proof(cin, CPU_SAMPLE_N, seed)
            k17_equal = gproof == cproof and small.vk_bytes() == op.vk_bytes()
            small.close()
            op.close()
            gate = {"k17_sample_proof_and_vk_bytes_equal_oracle": bool(k17_equal)}
            if not split:
                t0 = time.perf_counter()
                big = oracle_lib.OracleProver(oracle, k=k, seed=seed)
                big.keygen_vk(ROUNDS, n)
                vk_equal = big.vk_bytes() == ctx.vk_bytes()
                rc, msg = big.verify(proof)
                big.close()
                gate.update({"bench_proof_accepted_by_oracle_verifier": rc == 0, "bench_vk_bytes_equal_oracle": bool(vk_equal),
                             "oracle_check_s": time.perf_counter() - t0})
            gate["ok"] = all(v for kk, v in gate.items() if kk != "oracle_check_s")
            line["parity_gate"] = gate
            line["cpu_baseline"] = {
                "value": CPU_SAMPLE_N / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "one proof of %d compressions at k=%d (create_proof incl. witness synthesis; "
                          "params/keygen excluded), %.1f s, C++ oracle restating halo2_proofs 0.3.0, std::thread x %d; "
                          "the same-config (k=19) figure and the one-thread figure are the --impl reference arm's"
                          % (CPU_SAMPLE_N, CPU_SAMPLE_K, cdt, cores),
                "phases": {"[Setup generation]": csetup["params_s"] + csetup["keygen_s"], "[Proof generation]": cdt,
                           "[Proof verification]": cverify, "witness synthesis (inside proof generation)": csynth,
                           "unit": "s"}}
            assert gate["ok"], "parity gate failed: %s" % gate
        print(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

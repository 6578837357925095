#!/usr/bin/env python
"""bench.py — headline benchmark of the BLAKE2f proving path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle restatement)

A "step" is one complete proof (witness -> commitments -> quotient -> multiopen -> IPA) of one
batch of synthetic EIP-152 records: BASELINE.json configs[2], 64 twelve-round compressions in
one circuit (k = 19).  With N GPUs every rank proves its own independent batch (configs[4]:
independent proof streams, weak scaling, no data-path collective).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definitions.
"""
import argparse
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
MAC_PER_FP_MUL = 136


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


# ---- CPU arm: the oracle (halo2-0.3.0-equivalent restatement), all host threads ------------------
CPU_SAMPLE_K, CPU_SAMPLE_N = 17, 26   # bounded sample: the largest batch that fits the default k


def cpu_oracle_setup():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import zk_odst_b200 as zk
    oracle = oracle_lib.load()
    op = oracle_lib.OracleProver(oracle, k=CPU_SAMPLE_K, seed=zk.REFERENCE_SEED)
    op.keygen(ROUNDS, CPU_SAMPLE_N)
    return op, zk.synthetic_inputs(CPU_SAMPLE_N), zk.REFERENCE_SEED


def cpu_prove_once(op, inputs, seed):
    t0 = time.perf_counter()
    proof = op.create_proof(inputs, CPU_SAMPLE_N, seed)
    dt = time.perf_counter() - t0
    return CPU_SAMPLE_N / dt, dt, proof


def cpu_sample_text(dt):
    return ("one proof of %d compressions at k=%d (create_proof only; params/keygen excluded), "
            "%.1f s, C++ oracle restating halo2_proofs 0.3.0, std::thread x %d" %
            (CPU_SAMPLE_N, CPU_SAMPLE_K, dt, os.cpu_count() or 1))


def workload_config(k, n, split=False, world=1):
    what = ("configs[3] shape: ONE proof of %d compressions, every MSM split by point range over %d "
            "GPUs (NCCL all-gather of partial points), the witness transforms sharded by column, the "
            "quotient by row, evaluations by coefficient range; the rest replicated" % (n, world)) if split else (
            "configs[2]: %d twelve-round BLAKE2f compressions in one circuit, "
            "full create_proof (Pasta/IPA)" % n)
    return {"workload": what,
            "k": k, "rounds": ROUNDS, "compressions_per_proof": n,
            "rows_per_compression": 292 + 392 * ROUNDS,
            "params": "substitute URS (zk_params_generate_substitute, reference seed)",
            "l2": "per proof the prover streams > 2 GB of column/coset data (advice cosets alone "
                  "12 x 3 x 2^k x 32 B = %d MB), far above the 126 MB L2; no explicit flush" %
                  (12 * 3 * (1 << k) * 32 >> 20)}


def run_reference(args, rank):
    if rank != 0:
        return
    op, inputs, seed = cpu_oracle_setup()
    budget_s = 150.0
    steps, t_used, last_dt = 0, 0.0, 0.0
    t_start = time.perf_counter()
    while steps < max(1, args.steps):
        _, dt, _ = cpu_prove_once(op, inputs, seed)
        steps += 1
        t_used += dt
        last_dt = dt
        if time.perf_counter() - t_start + dt > budget_s:
            break
    dt = t_used / steps
    val = CPU_SAMPLE_N / dt
    cfg = workload_config(CPU_SAMPLE_K, CPU_SAMPLE_N)
    cfg["workload"] += " — CPU arm runs the bounded sample below (k=17 is the smallest circuit)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic", "config": cfg, "proofs_per_sec": 1.0 / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": cpu_sample_text(last_dt)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compressions", type=int, default=N_COMPRESSIONS)
    ap.add_argument("--streams", type=int, default=4,
                    help="concurrent proof streams per GPU (one context and host thread each)")
    ap.add_argument("--blocking-sync", type=int, default=-1,
                    help="1: host threads sleep while waiting for the device, 0: spin, -1: sleep only "
                         "when ranks x streams would oversubscribe the cores")
    ap.add_argument("--msm-split", action="store_true",
                    help="configs[3]: ONE proof stream, every MSM split by point range across the "
                         "ranks (NCCL all-gather of partial points), transforms by column, quotient by row; "
                         "strong scaling")
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
    if rank == 0 and not split:
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

    if rank == 0:
        hbm_peak, hbm_kind = load_peaks()
        int_peak = ctx.bench_int_pipe(1, 20000)  # mad.wide.u32 instructions/s = 32x32->64 MAC/s
        steps = single_steps
        per = {name: (ms / steps, cnt / steps) for name, (ms, cnt) in report.items() if cnt}
        R = 292 + 392 * ROUNDS
        # dominant kernel: MSM bucket accumulation
        # terms the accumulate launches process per proof: 13 full-width commitments (2 lookup, 5 grand
        # products, random, 3 h pieces, q', s), 12 advice columns (<= 32-bit cells), and the IPA: 5 rounds
        # of n terms on the original generators, then k - 5 rounds of n / 32 terms on the folded ones
        # (halo2's schedule would be 2n MSM terms plus n generator-folding scalar multiplications)
        fold = 5 if (k > 13 and (not split or 32 % world == 0)) else k
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
            # dram__bytes_read.sum + dram__bytes_write.sum of one full-width launch (8.39 M sorted
            # entries, k = 19) from the ncu --set full capture profiles/r01_fixed_accumulate_ncu_full_v2.txt;
            # algorithmic bytes of that launch: 8.39 M x (64 B point + 4 B index) = 0.57 GB
            "traffic": 1.073e9 if k == 19 else None,
            "peak_source": "zk_bench_int_pipe mode 1 (mad.wide.u32) measured in this run; "
                           "MEASURED_PEAKS.json has no integer peak",
            "launches_per_proof": acc_launches, "avg_launch_ms": acc_ms / acc_launches if acc_launches else None,
            "algorithmic_mac_per_proof": mac_per_proof,
            "accounting": "SURVEY.md 8d per-term figures (23936 MAC full-width, 2992 MAC advice) x the "
                          "terms the launches process: %.2f n full-width + 12 n advice%s"
                          % (full_terms, " (rank 0's 1/%d share of each MSM)" % world if split else ""),
        }
        ntt_ms, ntt_launches = per.get("ntt", (0.0, 0.0))
        # 19 inverse transforms to coefficients, 19 columns x 3 coset transforms, 3 inverse transforms of
        # h's coset values: all of size n (the quotient lives on three cosets of the n-th roots)
        # (MSM-split group: rank 0 transforms ceil(19 / world) of the columns and evaluates 1 / world of the
        # quotient rows; the 3 transforms of h stay replicated)
        ext = 3 * nrows // share
        own_cols = -(-19 // share)
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
        }
        if not args.no_cpu_baseline:
            op, cin, cseed = cpu_oracle_setup()
            v, cdt, _ = cpu_prove_once(op, cin, cseed)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": cpu_sample_text(cdt)}
        print(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

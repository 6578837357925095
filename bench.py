#!/usr/bin/env python
"""bench.py — headline benchmark of the BLAKE2f proving path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle restatement)

A "step" is one pass of the hot path over one batch of synthetic EIP-152 records
(BASELINE.json configs[2]: 64 twelve-round compressions in one circuit, k = 19).
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


def cpu_witness_baseline(inputs, n, k):
    """Oracle (CPU restatement) timed on the host cores: the `port` baseline."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle = oracle_lib.load()
    t0 = time.perf_counter()
    oracle.witness(k, ROUNDS, inputs, n)
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import zk_odst_b200 as zk
    n = N_COMPRESSIONS
    k = zk.min_k(ROUNDS, n)
    inputs = zk.synthetic_inputs(n)
    for _ in range(args.warmup):
        cpu_witness_baseline(inputs, n, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_witness_baseline(inputs, n, k)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": "blake2f_witness_compressions_per_sec", "value": val,
        "unit": "compressions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(k, n),
        "cpu_baseline": {"value": val, "unit": "compressions/s", "cores": 1, "kind": "port",
                         "sample": "%d compressions, witness generation (oracle, 1 thread)" % n},
        "e2e": {"value": val, "unit": "compressions/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(k, n):
    return {"workload": "configs[2]: %d twelve-round BLAKE2f compressions in one circuit" % n,
            "k": k, "rounds": ROUNDS, "compressions_per_batch": n,
            "rows_per_compression": 292 + 392 * ROUNDS,
            "phase": "witness generation (K1); prover phases land in later commits",
            "l2": "advice output 12*2^k*32 B = %d MB per step exceeds the 126 MB L2; no flush" %
                  (12 * (1 << k) * 32 >> 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_odst_b200 as zk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = N_COMPRESSIONS
    k = zk.min_k(ROUNDS, n)
    nrows = 1 << k
    inputs = zk.synthetic_inputs(n, stream=rank)  # independent batch per rank (weak scaling)
    ctx = zk.Context(local_rank)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream: events and kernels share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.enable_timing(True)

    d_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).cuda()
    d_adv = torch.empty((12, nrows, 4), dtype=torch.int64, device="cuda")
    d_dig = torch.empty((n, 8), dtype=torch.int64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ctx.witness_batch_device(k, ROUNDS, d_in, n, d_adv, d_dig)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        # per-launch duration of the dominant kernel, measured live with CUDA events on the
        # launching stream (a few extra steps, outside the step timing)
        for _ in range(min(args.steps, 10)):
            step()
            kernel_ms.append(ctx.last_kernel_ms(0))
    launches = ctx.launch_count() - launches0 - 2 * min(args.steps, 10)
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # end-to-end through the C-ABI call with host buffers (pinned), copies inside the timing
    h_in = torch.frombuffer(bytearray(inputs), dtype=torch.uint8).pin_memory()
    h_adv = torch.empty((12, nrows, 4), dtype=torch.int64).pin_memory()
    h_dig = torch.empty((n, 8), dtype=torch.int64).pin_memory()
    ctx.set_stream(None)
    e2e_steps = max(3, min(args.steps, 5))
    ctx.witness_batch(k, ROUNDS, h_in, n, h_adv, h_dig)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.witness_batch(k, ROUNDS, h_in, n, h_adv, h_dig)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * n / float(te.item())

    if rank == 0:
        peak, peak_kind = load_peaks()
        R = 292 + 392 * ROUNDS
        alg_bytes = n * (R * 12 * 32 + 213)
        k_ms = sorted(kernel_ms)[len(kernel_ms) // 2]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "blake2f_witness_compressions_per_sec", "value": value,
            "unit": "compressions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(k, n), "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": "compressions/s",
                    "h2d_bytes_per_step": len(inputs),
                    "d2h_bytes_per_step": 12 * nrows * 32 + n * 64},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "blake2f_witness_kernel", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes},
        }
        if not args.no_cpu_baseline:
            v, dt = cpu_witness_baseline(inputs, n, k)
            line["cpu_baseline"] = {
                "value": v, "unit": "compressions/s", "cores": 1, "kind": "port",
                "sample": "%d compressions, witness generation, %.2f s" % (n, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

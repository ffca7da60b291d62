#!/usr/bin/env python3
"""bench.py — headline benchmark of the commitment hot path (BASELINE.json: BN254 G1 MSM Mpoints/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 24] [--curve bn254]

A "step" is ONE CommitmentKey::commit (src/commitment.rs:78-87) over one synthetic vector:
N = 1 : 2^24 uniform random BN254 scalars against a 2^24-point key (BASELINE.json configs[1], the size the
        metric's target is quoted on);
N > 1 : the key is sharded by point range, 2^24 points per rank ("weak": at N = 4 this is configs[4]'s 2^26
        MSM); every rank produces a 128-byte XYZZ partial, one NCCL all_gather moves N x 128 B, rank 0 folds
        and normalises.
`value` = points committed by all ranks / max-over-ranks CUDA-event time, scalars resident in HBM.
`e2e`   = the same through the host-buffer C-ABI call (mira_msm_commit / mira_msm_partial): pinned host
          scalars -> H2D -> MSM -> 64-byte result on the host, copies inside the timed region.
`--impl reference` times the CPU restatement of the reference's rayon multiexp (oracle/, "port": the Rust
reference cannot be built here) on the box's host cores on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED_BASES = 0x4D495241          # "MIRA"
SEED_SCALARS = 0x4D495242
IMAD_WIDE_PER_CLK_PER_SM = 32    # measured, profiles/r01_intpipe_microbench.jsonl
MACS_PER_MODMUL = 136            # 8-limb CIOS: 2*8^2 + 8
MODMUL_PER_MADD = 10             # XYZZ mixed add 8M + 2S (canonical count: 1,360 wide MACs per pair)
MACS_PER_MADD_EXECUTED = 8 * 136 + 200   # 8 products + one dual product a*b + c*d under a single reduction (192 + 8)
ACC_DRAM_BYTES_PER_PAIR = 29.08e9 / 201326592   # measured, profiles/r01_accumulate_v12.txt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def curve_id(name: str) -> int:
    return {"bn254": 0, "grumpkin": 1}[name]


# ------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    from mira_b200 import CommitmentKey, combine_partials
    import gpu_util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — mira_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG >= VERSION (WARN included); the contract is ONE
        # JSON line, so run it silent unless the caller asks for NCCL logs explicitly
        if "MIRA_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["MIRA_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    curve = curve_id(args.curve)
    n = 1 << args.log_n                      # points per rank
    first = rank * n

    bases = gpu_util.gen_bases_dev(curve, SEED_BASES, n, first=first, device=local)
    scalars = gpu_util.gen_scalars_dev(curve, SEED_SCALARS, n, 0, first=first, device=local)
    ck = CommitmentKey(curve, bases, device=local, on_device=True)
    ck.prepare(n)
    del bases
    host_scalars = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    host_scalars.copy_(scalars)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    gather_buf = torch.empty(world * 128, dtype=torch.uint8, device=dev) if world > 1 else None

    def step(device_resident: bool) -> bytes:
        if world == 1:
            if device_resident:
                return ck.commit_device(scalars.data_ptr(), n, stream.cuda_stream)
            return ck.commit(host_scalars)
        if device_resident:
            part = ck.partial(scalars.data_ptr(), n, on_device=True, stream=stream.cuda_stream)
        else:
            part = ck.partial(host_scalars)
        mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
        dist.all_gather_into_tensor(gather_buf, mine)     # N x 128 B over NVLink: the path's one exchange step
        if rank == 0:
            return combine_partials(curve, gather_buf.cpu().numpy().tobytes(), local)
        return b""

    def timed(device_resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(device_resident)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        for _ in range(steps):
            res = step(device_resident)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, res_dev = timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    st = ck.stats()                       # launches / window of the device-resident commit (the timed `value` region)
    ms_e2e, res_e2e = timed(False, args.steps, max(1, min(args.warmup, 2)))

    # per-kernel time of the dominant kernel (bucket accumulation) from CUDA events inside the library
    ck.set_profiling(True)
    acc_ms = []
    for _ in range(3):
        if world == 1:
            ck.commit_device(scalars.data_ptr(), n, stream.cuda_stream)
        else:
            ck.partial(scalars.data_ptr(), n, on_device=True, stream=stream.cuda_stream)
        acc_ms.append(ck.stats())
    ck.set_profiling(False)
    prof = {k: statistics.mean(s[k] for s in acc_ms) for k in ("ms_digits", "ms_sort", "ms_accumulate", "ms_reduce", "ms_total")}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert res_dev == res_e2e and len(res_dev) == 64, "device-resident and host-buffer commits disagree"

    total_points = n * world
    ms_step = ms_dev / args.steps
    value = total_points / (ms_step * 1e-3) / 1e6
    e2e_value = total_points / (ms_e2e / args.steps * 1e-3) / 1e6
    hbm_peak, peak_src, sm_max = measured_peaks()
    entries = st["entries"]
    # roofline of k_accumulate.  Algorithmic HBM bytes per launch: one 64 B affine point + 4 B key + 4 B ref per
    # (point, window) entry.  Its binding resource is the integer pipe (IMAD.WIDE.U32), reported beside it.
    acc_s = prof["ms_accumulate"] * 1e-3
    alg_bytes = entries * 72.0
    macs = entries * MACS_PER_MADD_EXECUTED
    macs_canonical = entries * MODMUL_PER_MADD * MACS_PER_MODMUL
    imad_peak = 148 * IMAD_WIDE_PER_CLK_PER_SM * sm_max * 1e6
    out = {
        "metric": "BN254 G1 MSM Mpoints/s" if curve == 0 else "Grumpkin G1 MSM Mpoints/s",
        "value": round(value, 2), "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery, integer)", "data": "synthetic",
        "config": {"workload": f"{args.curve} G1 MSM, 2^{args.log_n} points per GPU, uniform random scalars, "
                               f"CommitmentKey::commit", "points_total": total_points, "points_per_gpu": n,
                   "window_bits": st["window_bits"], "windows": st["windows"], "buckets": st["buckets"],
                   "parallelism": f"point-range shards x{world}, 128 B NCCL all_gather of XYZZ partials" if world > 1 else "single GPU",
                   "l2": "inputs (scalars + fixed-base table) are >> 126 MB L2; no flush needed",
                   "seed": SEED_BASES},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "Mpoints/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 64 + 4 * world + (128 * world if world > 1 else 0)},
        "gpu_launches": int(st["kernel_launches"]) * args.steps * world,
        "phases_ms": {k: round(v, 3) for k, v in prof.items()},
        "roofline": {"kernel": "k_accumulate (+k_combine)", "bound": "hbm", "achieved": round(alg_bytes / acc_s / 1e9, 1),
                     "peak": hbm_peak, "unit": "GB/s", "frac": round(alg_bytes / acc_s / 1e9 / hbm_peak, 4),
                     "traffic": round(entries * ACC_DRAM_BYTES_PER_PAIR / 1e9, 2), "traffic_unit": "GB per launch",
                     "traffic_source": "ncu --set full, profiles/r01_accumulate_v12.txt: 29.08 GB dram read+write for 201.3 M pairs "
                                       "(each 64 B gathered point costs a 128 B DRAM burst), scaled to this launch's pairs",
                     "peak_source": peak_src,
                     "note": "kernel is integer-pipe bound, see roofline_imad"},
        "roofline_imad": {"kernel": "k_accumulate (+k_combine)", "bound": "imad.wide.u32", "achieved": round(macs / acc_s / 1e12, 3),
                          "peak": round(imad_peak / 1e12, 3), "unit": "T wide-MAC/s", "frac": round(macs / acc_s / imad_peak, 4),
                          "achieved_canonical": round(macs_canonical / acc_s / 1e12, 3),
                          "note": "achieved = wide MACs actually executed (1,288 per pair); achieved_canonical uses SURVEY.md "
                                  "8d's 10 x 136 = 1,360 per pair",
                          "peak_source": "measured 32 IMAD.WIDE.U32 lanes/clk/SM x 148 SMs x max SM clock"},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(curve, args.cpu_sample_log_n, first_seed=(SEED_BASES, SEED_SCALARS))
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(curve: int, log_n: int, first_seed, reps: int = 1) -> dict:
    import oracle_lib as O
    n = 1 << log_n
    bases = O.gen_bases(curve, first_seed[0], n)
    scalars = O.gen_scalars(curve, first_seed[1], n)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        O.commit(curve, bases, scalars, 0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": round(n / best / 1e6, 4), "unit": "Mpoints/s", "cores": O.num_cores(), "kind": "port",
            "sample": f"first 2^{log_n} points of the same workload, {best:.2f} s; C restatement of halo2 best_multiexp "
                      f"(oracle/mira_oracle.c), one chunk per core"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle_lib as O
    curve = curve_id(args.curve)
    log_n = args.cpu_sample_log_n
    n = 1 << log_n
    bases = O.gen_bases(curve, SEED_BASES, n)
    scalars = O.gen_scalars(curve, SEED_SCALARS, n)
    for _ in range(args.warmup):
        O.commit(curve, bases, scalars, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.commit(curve, bases, scalars, 0)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt / 1e6
    out = {
        "impl": "reference", "metric": "BN254 G1 MSM Mpoints/s" if curve == 0 else "Grumpkin G1 MSM Mpoints/s",
        "value": round(value, 4), "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64x4 (254-bit Montgomery, integer)", "data": "synthetic",
        "config": {"workload": f"{args.curve} G1 MSM, 2^{args.log_n} points per GPU, uniform random scalars, CommitmentKey::commit",
                   "sample": f"each step commits the first 2^{log_n} points of that workload on the host CPU"},
        "cpu_baseline": {"value": round(value, 4), "unit": "Mpoints/s", "cores": O.num_cores(), "kind": "port",
                         "sample": f"2^{log_n} points per step; C restatement of halo2 best_multiexp (the Rust reference "
                                   f"cannot be built in this image: no cargo/rustc, un-vendored git deps)"},
        "e2e": {"value": round(value, 4), "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ fold-step replay
def run_fold_step(args):
    """`--workload fold-step`: BASELINE.json's second metric (IVC fold-step ms), as a replay of the hot-path work of one
    SnarkStar fold step (tools/fold_step.py; the Rust driver itself cannot be built here).  Single GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fold_step as F
    hbm_peak, peak_src, sm_max = measured_peaks()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cpu = F.CpuFoldStep(args.log_rows)
        for _ in range(min(args.warmup, 1)):
            cpu.step()
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 2))
        for _ in range(steps):
            cpu.step()
        ms = (time.perf_counter() - t0) / steps * 1e3
        import oracle_lib as O
        out = {"impl": "reference", "metric": "IVC fold-step hot-path ms", "value": round(ms, 1), "unit": "ms", "n_gpus": 1,
               "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(ms, 1), "higher_is_better": False,
               "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (254-bit Montgomery, integer)", "data": "synthetic",
               "config": {"workload": f"SnarkStar fold-step replay, k={args.log_rows}: witness commits + cross-term evaluation + "
                                      f"cross-term commits + fold, both circuits", "points_per_step": F.points_per_step(cpu.sh)},
               "cpu_baseline": {"value": round(ms, 1), "unit": "ms", "cores": O.num_cores(), "kind": "port",
                                "sample": "one full step on the host CPU through oracle/ (C restatement; the Rust reference cannot be built here)"},
               "e2e": {"value": round(ms, 1), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out), flush=True)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — mira_b200 has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_fold_step_sharded(args, F)
    g = F.GpuFoldStep(args.log_rows)

    def timed(from_host, steps, warmup):
        for _ in range(warmup):
            res = g.step(from_host)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(g.stream)
        for _ in range(steps):
            res = g.step(from_host)
        e1.record(g.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, res

    sampler = ClockSampler(0)
    sampler.start()
    ms_dev, res_dev = timed(False, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, res_e2e = timed(True, args.steps, 1)
    assert res_dev == res_e2e
    # phase split with CUDA events (one extra step): evaluation kernels alone
    sh, st = g.sh, g.state
    ev_ms, muls = 0.0, 0
    from mira_b200 import witness as W
    for s, t in zip(sh, st):
        dom = W.PlonkEvalDomain(s["meta"]["num_advice"], 0, t["ch"], [], t["fixed"], [t["W1"]], [t["W2"]])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(g.stream)
        W.evaluate_rows_multi(t["progs"], dom, outs=t["T"], stream=g.stream.cuda_stream)
        e1.record(g.stream)
        torch.cuda.synchronize()
        ev_ms += e0.elapsed_time(e1)
        muls += t["progs"][0].stats()["muls"] * s["rows"]        # products actually executed by the merged program
    imad_peak = 148 * IMAD_WIDE_PER_CLK_PER_SM * sm_max * 1e6
    pts = F.points_per_step(sh)
    h2d = sum(s["n_w"] * 32 for s in sh)
    ncommit = sum(1 + len(s["progs"]) for s in sh)
    launches = sum(int(t["ck"].stats()["kernel_launches"]) for t in st)   # last commit of each key; lower bound per commit
    out = {"metric": "IVC fold-step hot-path ms", "value": round(ms_dev, 3), "unit": "ms", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(ms_dev, 3), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
           "dtype": "u32x8 (254-bit Montgomery, integer)", "data": "synthetic",
           "config": {"workload": f"SnarkStar fold-step replay, k={args.log_rows}: witness commits + cross-term evaluation + "
                                  f"cross-term commits + fold, both circuits", "points_per_step": pts, "commits_per_step": ncommit,
                      "l2": "working set (keys' fixed-base tables, W, fixed columns) is >> 126 MB L2; no flush needed",
                      "seed": F.SEED},
           "clocks": clocks,
           "e2e": {"value": round(ms_e2e, 3), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64 * ncommit},
           "gpu_launches": (ncommit * launches // max(len(st), 1) + g.launches_per_step()) * args.steps,
           "phases_ms": {"cross_term_evaluation": round(ev_ms, 3), "commits_and_fold": round(ms_dev - ev_ms, 3)},
           "roofline": {"kernel": "k_eval_rows (all 11 cross-term programs, merged per circuit)", "bound": "hbm", "achieved": None, "peak": hbm_peak,
                        "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                        "note": "integer-pipe bound kernel, see roofline_imad; the MSM kernels' roofline is in the default workload's line"},
           "roofline_imad": {"kernel": "k_eval_rows", "bound": "imad.wide.u32", "achieved": round(muls * MACS_PER_MODMUL / (ev_ms * 1e-3) / 1e12, 3),
                             "peak": round(imad_peak / 1e12, 3), "unit": "T wide-MAC/s",
                             "frac": round(muls * MACS_PER_MODMUL / (ev_ms * 1e-3) / imad_peak, 4)},
           "commitments_sha256": __import__("hashlib").sha256(b"".join(res_dev)).hexdigest()}
    if not args.no_cpu_baseline:
        import oracle_lib as O
        cpu = F.CpuFoldStep(args.log_rows, inputs=g.host_inputs())
        t0 = time.perf_counter()
        want = cpu.step()
        dt = (time.perf_counter() - t0) * 1e3
        assert want == res_dev, "GPU fold-step commitments differ from the CPU oracle's"
        out["cpu_baseline"] = {"value": round(dt, 1), "unit": "ms", "cores": O.num_cores(), "kind": "port",
                               "sample": "one full step of the same workload (same bytes) through oracle/; commitments compared bit for bit"}
    print(json.dumps(out), flush=True)


def run_fold_step_sharded(args, F):
    """`--workload fold-step` under torchrun: the circuit's rows are cut into one range per rank (STRONG scaling: the
    step is the same whatever N is).  Per step a rank queues its partial commitments, evaluation and folds on its own
    stream; the one collective is an all_gather of 13 x 128 B per rank; every rank then folds the partials on its GPU."""
    import torch
    import torch.distributed as dist
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if "MIRA_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["MIRA_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)
    dist.init_process_group("nccl", device_id=dev)
    g = F.ShardedGpuFoldStep(args.log_rows, rank, world, device=local)
    gathered = torch.empty(world * g.n_commits * 128, dtype=torch.uint8, device=dev)

    def step():
        part = g.step()
        with torch.cuda.stream(g.stream):
            dist.all_gather_into_tensor(gathered, part)
        return g.combine(gathered, world)          # 13 commitments on every rank; returns when they are in host memory

    for _ in range(args.warmup):
        res = step()
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(g.stream)
    for _ in range(args.steps):
        res = step()
    e1.record(g.stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    # every rank must hold the same commitments
    digest = torch.tensor(list(b"".join(res)), dtype=torch.uint8, device=dev)
    all_d = torch.empty(world * digest.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_d, digest)
    same = bool((all_d.view(world, -1) == digest).all().item())
    launches = sum(int(t["ck"].stats()["kernel_launches"]) for t in g.state)
    if rank == 0:
        assert same, "ranks disagree on the combined commitments"
        ncommit = g.n_commits
        out = {"metric": "IVC fold-step hot-path ms", "value": round(ms.item(), 3), "unit": "ms", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms.item(), 3), "higher_is_better": False, "scaling": "strong",
               "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery, integer)", "data": "synthetic",
               "config": {"workload": f"SnarkStar fold-step replay, k={args.log_rows}: witness commits + cross-term evaluation + "
                                      f"cross-term commits + fold, both circuits", "points_per_step": F.points_per_step(g.sh),
                          "commits_per_step": ncommit,
                          "parallelism": f"row-range shards x{world}: local evaluation and folds, one all_gather of {ncommit} x 128 B XYZZ "
                                         f"partials per rank, combine on every rank",
                          "l2": "working set (keys' fixed-base tables, W, fixed columns) is >> 126 MB L2; no flush needed",
                          "seed": F.SEED},
               "clocks": clocks,
               "e2e": None,
               "gpu_launches": (2 * launches + 2 * len(g.state) * 2 + 2) * args.steps * world,
               "commitments_sha256": __import__("hashlib").sha256(b"".join(res)).hexdigest()}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24, help="log2(points per GPU)")
    ap.add_argument("--curve", default="bn254", choices=["bn254", "grumpkin"])
    ap.add_argument("--cpu-sample-log-n", type=int, default=21, help="log2(points) of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="msm", choices=["msm", "fold-step"],
                    help="msm: the headline BN254 MSM (default); fold-step: replay of one SnarkStar IVC fold step")
    ap.add_argument("--log-rows", type=int, default=19, help="fold-step: log2(rows) of the circuit tables (k)")
    args = ap.parse_args()
    if args.workload == "fold-step":
        return run_fold_step(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU (the driver uses torchrun directly)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

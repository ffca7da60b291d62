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
MACS_PER_MODMUL = 136            # 8-limb CIOS: 2*8^2 + 8
MODMUL_PER_MADD = 10             # XYZZ mixed add 8M + 2S (canonical count: 1,360 wide MACs per pair)
MACS_PER_MADD_EXECUTED = 8 * 136 + 200   # 8 products + one dual product a*b + c*d under a single reduction (192 + 8)
ACC_DRAM_BYTES_PER_PAIR = 29.08e9 / 201326592   # measured, profiles/r01_accumulate_v12.txt


def imad_peak_lanes():
    """IMAD.WIDE.U32 issue rate (lanes / clk / SM) from the tracked measurement profiles/imad_peak.json."""
    p = os.path.join(ROOT, "profiles", "imad_peak.json")
    with open(p) as f:
        d = json.load(f)
    return float(d["imad_wide_lanes_per_clk_per_sm"]), d["source"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback", 1965.0


def setup_nccl_logging():
    """NCCL's INFO log goes to stderr (NCCL_DEBUG_FILE) so that stdout stays ONE JSON line while a caller's NCCL_DEBUG
    setting (the driver reads the rank count from it) is kept, not deleted."""
    if "MIRA_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["MIRA_NCCL_DEBUG"]
    if os.environ.get("NCCL_DEBUG") and "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"


def workload_config(curve_name: str, total_log: int, world: int) -> dict:
    """The workload both arms (ours / --impl reference) name: same keys, same values."""
    n_total = 1 << total_log
    return {"workload": f"{curve_name} G1 MSM, 2^{total_log} points, uniform random scalars, CommitmentKey::commit"
                        + (f", key sharded by point range over {world} GPUs" if world > 1 else ""),
            "points_total": n_total, "points_per_gpu": n_total // world,
            "parallelism": (f"point-range shards x{world} (2^{total_log} / {world} points per rank), one all_gather of a 128 B "
                            f"XYZZ partial per rank over NCCL, combine on the device") if world > 1 else "single GPU",
            "l2": "inputs (scalars + fixed-base table) are >> 126 MB L2; no flush needed",
            "seed": SEED_BASES}


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

    from mira_b200 import CommitmentKey, combine_partials_device
    from mira_b200.sharding import shard_range
    import gpu_util
    import oracle_lib as O          # the CHECKER (parity key) and the cpu_baseline leg; never on the timed GPU path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — mira_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        setup_nccl_logging()
        dist.init_process_group("nccl", device_id=dev)
    curve = curve_id(args.curve)
    # N = 1: configs[1], 2^24 points on one GPU.  N > 1: configs[4], ONE fixed 2^26-point MSM split by point range
    # (strong scaling: total work does not grow with N).
    total_log = args.log_n if world == 1 else args.log_total
    stream = torch.cuda.current_stream()

    class Shard:
        """This rank's point range of a 2^total_log-point key, its scalars (device + pinned host) and context."""
        def __init__(self, total_log):
            self.n_total = 1 << total_log
            self.lo, self.hi = shard_range(self.n_total, world, rank)
            self.n = self.hi - self.lo
            self.bases = gpu_util.gen_bases_dev(curve, SEED_BASES, self.n, first=self.lo, device=local)
            self.scalars = gpu_util.gen_scalars_dev(curve, SEED_SCALARS, self.n, 0, first=self.lo, device=local)
            self.ck = CommitmentKey(curve, self.bases, device=local, on_device=True)
            self.ck.prepare(self.n)
            self.host_scalars = torch.empty(self.n * 32, dtype=torch.uint8, pin_memory=True)
            self.host_scalars.copy_(self.scalars)
            self.part = torch.zeros(128, dtype=torch.uint8, device=dev)
            self.gathered = torch.zeros(world * 128, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()

        def step(self, device_resident: bool) -> bytes:
            ck, n = self.ck, self.n
            if world == 1:
                if device_resident:
                    return ck.commit_device(self.scalars.data_ptr(), n, stream.cuda_stream)
                return ck.commit(self.host_scalars)
            if device_resident:       # nothing leaves the device until the 64-byte result: partial -> all_gather -> combine
                ck.partial_batch_device([self.scalars.data_ptr()], n, self.part.data_ptr(), stream.cuda_stream)
            else:                     # host scalars: the C-ABI call copies them in slices behind the accumulation
                self.part.copy_(torch.frombuffer(bytearray(ck.partial(self.host_scalars)), dtype=torch.uint8), non_blocking=False)
            dist.all_gather_into_tensor(self.gathered, self.part)       # N x 128 B over NVLink: the path's one exchange step
            return combine_partials_device(curve, self.gathered.data_ptr(), world, 1, 128, local, stream.cuda_stream)[0]

        def timed(self, device_resident: bool, steps: int, warmup: int):
            for _ in range(warmup):
                self.step(device_resident)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            res = None
            for _ in range(steps):
                res = self.step(device_resident)
            e1.record(stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), res

    sh = Shard(total_log)
    n, ck = sh.n, sh.ck
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, res_dev = sh.timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    st = ck.stats()                       # launches / window of the device-resident commit (the timed `value` region)
    ms_e2e, res_e2e = sh.timed(False, args.steps, args.warmup)

    # per-kernel time of the dominant kernel (bucket accumulation) from CUDA events inside the library
    ck.set_profiling(True)
    acc_ms = []
    for _ in range(3):
        if world == 1:
            ck.commit_device(sh.scalars.data_ptr(), n, stream.cuda_stream)
        else:
            ck.partial(sh.scalars.data_ptr(), n, on_device=True, stream=stream.cuda_stream)
        acc_ms.append(ck.stats())
    ck.set_profiling(False)
    prof = {k: statistics.mean(s[k] for s in acc_ms) for k in ("ms_digits", "ms_sort", "ms_accumulate", "ms_reduce", "ms_total")}

    # ---- parity: the CUDA result against the CPU oracle on the same bytes (the oracle is the checker, outside every timed region)
    parity = {}
    if world == 1:
        pl = min(args.cpu_sample_log_n, total_log)
        host_bases = sh.bases.cpu().numpy()
        host_sc = sh.host_scalars.numpy()
        t0 = time.perf_counter()
        want_prefix = O.commit(curve, host_bases[: 64 << pl].tobytes(), host_sc[: 32 << pl].tobytes(), 0)
        prefix_s = time.perf_counter() - t0
        got_prefix = ck.commit_device(sh.scalars.data_ptr(), 1 << pl, stream.cuda_stream)
        parity = {"prefix_n": 1 << pl, "prefix_equal": got_prefix == want_prefix}
        if not args.no_full_parity:
            t0 = time.perf_counter()
            want_full = O.commit(curve, host_bases.tobytes(), host_sc.tobytes(), 0)
            parity.update({"n": n, "equal": res_dev == want_full and res_e2e == want_full,
                           "oracle_seconds": round(time.perf_counter() - t0, 2),
                           "what": "the timed 2^%d-point commitment (device-resident and host-buffer call) == oracle/ commit of "
                                   "the same bases and scalars, 64 bytes compared" % total_log})
        else:
            parity.update({"n": 1 << pl, "equal": parity["prefix_equal"]})
        del host_bases
    else:
        # (a) the combined commitment == point_add (oracle) over the ranks' own normalised commitments;
        # (b) every rank checks a prefix of ITS shard against the oracle on the same bytes
        mine = ck.commit_device(sh.scalars.data_ptr(), n, stream.cuda_stream)
        mine_t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
        all_t = torch.empty(world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(all_t, mine_t)
        pl = min(19, n.bit_length() - 1)
        got_prefix = ck.commit_device(sh.scalars.data_ptr(), 1 << pl, stream.cuda_stream)
        want_prefix = O.commit(curve, sh.bases[: 64 << pl].cpu().numpy().tobytes(), sh.host_scalars[: 32 << pl].numpy().tobytes(), 0)
        ok = torch.tensor([1 if got_prefix == want_prefix else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            pts = all_t.cpu().numpy().tobytes()
            acc = bytes(64)
            for g in range(world):
                acc = O.point_add(curve, acc, pts[64 * g:64 * g + 64])
            parity = {"n": sh.n_total, "equal": acc == res_dev and res_dev == res_e2e,
                      "what": "combined commitment == oracle point_add over the ranks' own affine commitments; host-buffer "
                              "path gives the same bytes",
                      "rank_prefix_n": 1 << pl, "rank_prefix_equal_all_ranks": bool(ok.item()),
                      "rank_prefix_what": "every rank: GPU commit of the first 2^%d points of its shard == oracle/ commit" % pl}

    # ---- extras: strong-scaling companions (not the headline)
    extras = {}
    if world > 1:
        # where a multi-rank step's time goes: every rank's own MSM (no collective), and the exchange + combine alone
        def local_ms(reps=3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                ck.partial_batch_device([sh.scalars.data_ptr()], n, sh.part.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        mine_ms = torch.tensor([local_ms()], device=dev, dtype=torch.float64)
        all_ms = torch.empty(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(all_ms, mine_ms)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            dist.all_gather_into_tensor(sh.gathered, sh.part)
            combine_partials_device(curve, sh.gathered.data_ptr(), world, 1, 128, local, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        g0.record(stream)
        for _ in range(20):
            dist.all_gather_into_tensor(sh.gathered, sh.part)
        g1.record(stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            combine_partials_device(curve, sh.gathered.data_ptr(), world, 1, 128, local, stream.cuda_stream)
        combine_ms = (time.perf_counter() - t0) / 20 * 1e3
        extras["step_breakdown"] = {"rank_msm_ms": [round(float(x), 3) for x in all_ms.cpu().tolist()],
                                    "exchange_and_combine_ms": round(e0.elapsed_time(e1) / 10, 3),
                                    "all_gather_alone_ms": round(g0.elapsed_time(g1) / 20, 3),
                                    "combine_alone_ms": round(combine_ms, 3),
                                    "what": "per-rank device time of the rank's own partial MSM (no collective); all_gather of N x 128 B + "
                                            "combine + 64 B D2H alone (rank 0's clock)"}
    sh.ck.close()
    del sh.bases
    if not args.no_extras:
        if world == 1:
            # the 2^26-point MSM of configs[4] on ONE GPU: the N = 1 point of the strong-scaling curve
            big = Shard(args.log_total)
            ms_big, res_big = big.timed(True, max(2, min(args.steps, 3)), 1)
            extras["strong_2p%d" % args.log_total] = {
                "points_total": big.n_total, "n_gpus": 1, "ms_per_step": round(ms_big / max(2, min(args.steps, 3)), 3),
                "value": round(big.n_total / (ms_big / max(2, min(args.steps, 3)) * 1e-3) / 1e6, 2), "unit": "Mpoints/s",
                "window_bits": big.ck.stats()["window_bits"]}
            big.ck.close()
            del big
        else:
            # the 2^24-point MSM of configs[1] split over the same ranks
            small = Shard(args.log_n)
            k = max(3, args.steps)
            ms_small, _ = small.timed(True, k, 2)
            if rank == 0:
                extras["strong_2p%d" % args.log_n] = {
                    "points_total": small.n_total, "n_gpus": world, "ms_per_step": round(ms_small / k, 3),
                    "value": round(small.n_total / (ms_small / k * 1e-3) / 1e6, 2), "unit": "Mpoints/s",
                    "window_bits": small.ck.stats()["window_bits"]}
            small.ck.close()
            del small

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert res_dev == res_e2e and len(res_dev) == 64, "device-resident and host-buffer commits disagree"
    assert parity.get("equal", False) and parity.get("prefix_equal", True) and parity.get("rank_prefix_equal_all_ranks", True), \
        f"PARITY FAILURE against the CPU oracle: {parity}"

    total_points = sh.n_total
    ms_step = ms_dev / args.steps
    value = total_points / (ms_step * 1e-3) / 1e6
    e2e_value = total_points / (ms_e2e / args.steps * 1e-3) / 1e6
    hbm_peak, peak_src, sm_max = measured_peaks()
    lanes, lanes_src = imad_peak_lanes()
    entries = st["entries"]
    # roofline of k_accumulate (the dominant kernel).  Its binding resource is the integer pipe: executed wide MACs
    # (IMAD.WIDE.U32) per launch over the kernel's CUDA-event time, against the measured issue rate of that
    # instruction.  Algorithmic HBM bytes (one 64 B affine point + 4 B key + 4 B ref per entry) are the secondary bound.
    acc_s = prof["ms_accumulate"] * 1e-3
    alg_bytes = entries * 72.0
    macs = entries * MACS_PER_MADD_EXECUTED
    macs_canonical = entries * MODMUL_PER_MADD * MACS_PER_MODMUL
    imad_peak = 148 * lanes * sm_max * 1e6
    cfg = workload_config(args.curve, total_log, world)
    out = {
        "metric": "BN254 G1 MSM Mpoints/s" if curve == 0 else "Grumpkin G1 MSM Mpoints/s",
        "value": round(value, 2), "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery, integer)", "data": "synthetic",
        "config": cfg,
        "msm": {"window_bits": st["window_bits"], "windows": st["windows"], "buckets": st["buckets"], "pairs_per_rank": entries},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "Mpoints/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": total_points * 32, "d2h_bytes_per_step": 64 * world + (128 * world if world > 1 else 0)},
        "gpu_launches": int(st["kernel_launches"] + (2 if world > 1 else 0)) * args.steps * world,
        "parity": parity,
        "phases_ms": {k: round(v, 3) for k, v in prof.items()},
        "roofline": {"kernel": "k_accumulate (+k_combine)", "bound": "imad.wide.u32", "achieved": round(macs / acc_s / 1e12, 3),
                     "peak": round(imad_peak / 1e12, 3), "unit": "T wide-MAC/s", "frac": round(macs / acc_s / imad_peak, 4),
                     "traffic": None,
                     "achieved_canonical": round(macs_canonical / acc_s / 1e12, 3),
                     "note": "integer-pipe bound kernel (no tensor cores: carry-chained 254-bit modular arithmetic). achieved = "
                             "wide MACs actually executed (1,288 per pair: 8 products + 1 dual product); achieved_canonical uses "
                             "SURVEY.md 8d's 10 x 136 = 1,360 per pair. HBM side: roofline_hbm",
                     "peak_source": f"{lanes:g} IMAD.WIDE.U32 lanes/clk/SM ({lanes_src}) x 148 SMs x max SM clock {sm_max:g} MHz"},
        "roofline_hbm": {"kernel": "k_accumulate (+k_combine)", "bound": "hbm", "achieved": round(alg_bytes / acc_s / 1e9, 1),
                         "peak": hbm_peak, "unit": "GB/s", "frac": round(alg_bytes / acc_s / 1e9 / hbm_peak, 4),
                         "traffic": round(entries * ACC_DRAM_BYTES_PER_PAIR / 1e9, 2), "traffic_unit": "GB per launch",
                         "traffic_source": "ncu --set full, profiles/r01_accumulate_v12.txt: 29.08 GB dram read+write for 201.3 M pairs "
                                           "(each 64 B gathered point costs a 128 B DRAM burst), scaled to this launch's pairs",
                         "peak_source": peak_src},
    }
    out.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = {"value": round((1 << pl) / prefix_s / 1e6, 4), "unit": "Mpoints/s", "cores": O.num_cores(), "kind": "port",
                               "sample": f"first 2^{pl} points of the same workload, {prefix_s:.2f} s; C restatement of halo2 best_multiexp "
                                         f"(oracle/mira_oracle.c), one chunk per core; its 64-byte result is the parity.prefix check"}
    if world == 1 and not args.no_extras and not args.no_fold_step:
        out["fold_step"] = fold_step_summary(args)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port of halo2's rayon multiexp; the Rust binary cannot be
    built here) on the box's host cores, all threads, on the SAME workload `config` as our arm.  A step commits a bounded
    prefix of that workload, sized from a probe so that the whole --steps/--warmup run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle_lib as O
    curve = curve_id(args.curve)
    total_log = args.log_n if world == 1 else args.log_total
    probe_log = min(19, total_log)
    bases = O.gen_bases(curve, SEED_BASES, 1 << probe_log)
    scalars = O.gen_scalars(curve, SEED_SCALARS, 1 << probe_log)
    t0 = time.perf_counter()
    O.commit(curve, bases, scalars, 0)
    probe = time.perf_counter() - t0
    # largest sample <= the workload whose (steps + warmup) commits plus key generation fit in ~150 s
    budget = 150.0 / max(1, args.steps + args.warmup)
    log_n = probe_log
    while log_n < min(total_log, args.cpu_sample_max_log_n) and probe * (1 << (log_n + 1 - probe_log)) * 1.15 <= budget:
        log_n += 1
    n = 1 << log_n
    if log_n > probe_log:
        bases = O.gen_bases(curve, SEED_BASES, n)     # all host cores; ~10 s at 2^23
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
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "u64x4 (254-bit Montgomery, integer)", "data": "synthetic",
        "config": workload_config(args.curve, total_log, world),
        "cpu_baseline": {"value": round(value, 4), "unit": "Mpoints/s", "cores": O.num_cores(), "kind": "port",
                         "sample": f"each step commits the first 2^{log_n} points of the workload on the host CPU (all cores); C "
                                   f"restatement of halo2 best_multiexp (the Rust reference cannot be built in this image: no "
                                   f"cargo/rustc, un-vendored git deps)"},
        "e2e": {"value": round(value, 4), "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ fold-step replay
def fold_step_summary(args) -> dict:
    """BASELINE.json's second metric inside the default N = 1 line: one SnarkStar fold step (k = 19) replayed on the GPU
    (device-resident and from pinned host memory), the same step through oracle/ on the host cores, commitments compared
    bit for bit.  See run_fold_step for the standalone line."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import torch
    import fold_step as F
    import oracle_lib as O
    k = args.log_rows
    g = F.GpuFoldStep(k)

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
    ms_dev, res_dev = timed(False, max(3, args.steps), 2)
    ms_e2e, res_e2e = timed(True, max(3, args.steps), 1)
    cpu = F.CpuFoldStep(k, inputs=g.host_inputs())
    t0 = time.perf_counter()
    want = cpu.step()
    cpu_ms = (time.perf_counter() - t0) * 1e3
    import gpu_util
    folds_equal = all(gpu_util.to_bytes(sg["W_out"]) == sc["W_out"].tobytes() and gpu_util.to_bytes(sg["E_out"]) == sc["E_out"].tobytes()
                      for sg, sc in zip(g.state, cpu.state))
    return {"k": k, "ms": round(ms_dev, 3), "e2e_ms": round(ms_e2e, 3), "cpu_baseline_ms": round(cpu_ms, 1), "cpu_cores": O.num_cores(),
            "equal": res_dev == want and res_e2e == want and folds_equal, "commitments": len(want),
            "points_per_step": F.points_per_step(g.sh), "h2d_bytes_per_step": sum(s["n_w"] * 32 for s in g.sh),
            "what": "SnarkStar fold-step replay (tools/fold_step.py): 2 witness commits, 11 cross-term evaluations, 11 cross-term "
                    "commits, 2 folds; all 13 commitments and both folded witnesses == oracle/ on the same bytes"}


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
    imad_peak = 148 * imad_peak_lanes()[0] * sm_max * 1e6
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
    setup_nccl_logging()
    dist.init_process_group("nccl", device_id=dev)
    g = F.ShardedGpuFoldStep(args.log_rows, rank, world, device=local)
    gathered = torch.empty(world * g.n_commits * 128, dtype=torch.uint8, device=dev)

    part_buf = g.partials

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
    # where the step's time goes: every rank's own share (no collective), and the exchange + combine alone
    torch.cuda.synchronize()
    dist.barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(g.stream)
    for _ in range(5):
        g.step()
    l1.record(g.stream)
    torch.cuda.synchronize()
    mine_ms = torch.tensor([l0.elapsed_time(l1) / 5], device=dev, dtype=torch.float64)
    all_ms = torch.empty(world, device=dev, dtype=torch.float64)
    dist.all_gather_into_tensor(all_ms, mine_ms)
    dist.barrier()
    torch.cuda.synchronize()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record(g.stream)
    for _ in range(10):
        with torch.cuda.stream(g.stream):
            dist.all_gather_into_tensor(gathered, part_buf)
        g.combine(gathered, world)
    x1.record(g.stream)
    torch.cuda.synchronize()
    breakdown = {"rank_share_ms": [round(float(x), 3) for x in all_ms.cpu().tolist()],
                 "exchange_and_combine_ms": round(x0.elapsed_time(x1) / 10, 3)}
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
               "step_breakdown": breakdown,
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
    ap.add_argument("--cpu-sample-max-log-n", type=int, default=24, help="--impl reference: largest sample the probe may pick")
    ap.add_argument("--log-total", type=int, default=26, help="log2(points) of the fixed MSM that N > 1 ranks split (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-parity", action="store_true", help="N = 1: skip the oracle run at the full 2^log-n (keeps the prefix check)")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling companion and the fold-step key")
    ap.add_argument("--no-fold-step", action="store_true")
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

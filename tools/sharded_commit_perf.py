"""Single-process multi-GPU commit (mira_msm_ctx_create_sharded): timing and bit-exactness against a single-device key.

    python tools/sharded_commit_perf.py [--log-n 24] [--devices 0,1,2,3,4,5,6,7] [--reps 5]

One JSON line per device count (1, 2, 4, ... up to the list given).  The call is host-synchronous (host scalars in,
64 bytes out), so it is timed with the wall clock around the C-ABI call; scalars are page-locked."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--devices", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--curve", type=int, default=0)
    args = ap.parse_args()
    import torch
    import gpu_util
    from mira_b200 import CommitmentKey
    ndev = torch.cuda.device_count()
    devs = [int(d) for d in args.devices.split(",")] if args.devices else list(range(ndev))
    n = 1 << args.log_n
    bases = gpu_util.gen_bases_dev(args.curve, 0x4D495241, n).cpu().numpy()
    sc_dev = gpu_util.gen_scalars_dev(args.curve, 0x4D495242, n, 0)
    scalars = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    scalars.copy_(sc_dev)
    del sc_dev
    torch.cuda.synchronize()
    want = None
    counts = [c for c in (1, 2, 4, 8, 16) if c <= len(devs)]
    for cnt in counts:
        use = devs[:cnt]
        t0 = time.perf_counter()
        ck = CommitmentKey.sharded(args.curve, bases, use)
        ck.prepare(n)
        setup_s = time.perf_counter() - t0
        got = ck.commit(scalars)                       # warm-up (workspace allocation)
        ck.commit(scalars)
        times = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            got = ck.commit(scalars)
            times.append((time.perf_counter() - t0) * 1e3)
        if want is None:
            want = got
        st = ck.stats()
        print(json.dumps({"what": "single-process sharded commit, host scalars (pinned) -> 64 B", "log_n": args.log_n, "devices": use,
                          "ms_best": round(min(times), 3), "ms_median": round(sorted(times)[len(times) // 2], 3),
                          "mpoints_per_s": round(n / (min(times) * 1e-3) / 1e6, 1), "equal_to_one_device": got == want,
                          "window_bits_shard0": st["window_bits"], "pairs": st["entries"], "launches": st["kernel_launches"],
                          "setup_s": round(setup_s, 2)}), flush=True)
        ck.close()


if __name__ == "__main__":
    main()

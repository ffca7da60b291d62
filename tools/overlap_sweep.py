"""Does the preparation (digits + radix sort) of slice k+1 hide behind the accumulation of slice k?  Sweep of the
development knobs: MIRA_ACC_PAD_KB (caps accumulation blocks per SM by unused shared memory), MIRA_RS_THREADS_RT (sort
block size), MIRA_PREP_PRIO (stream priority of the preparation) x the number of slices of a device-resident commit.
One child process per environment (the knobs are read once per process); one JSON line per configuration."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import gpu_util
    from mira_b200 import CommitmentKey
    log_n = int(os.environ.get("SWEEP_LOG_N", "24"))
    n = 1 << log_n
    bases = gpu_util.gen_bases_dev(0, 0x4D495241, n)
    sc = gpu_util.gen_scalars_dev(0, 0x4D495242, n, 0)
    ck = CommitmentKey(0, bases, on_device=True)
    ck.prepare(n)
    del bases
    stream = torch.cuda.current_stream()
    want = None
    for slices in [int(x) for x in os.environ.get("SWEEP_SLICES", "1,2,4,8").split(",")]:
        ck.set_pipeline(slices if slices > 1 else 0, 1 << 18)
        for _ in range(2):
            got = ck.commit_device(sc.data_ptr(), n, stream.cuda_stream)
        want = want or got
        assert got == want
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        e0.record(stream)
        for _ in range(reps):
            ck.commit_device(sc.data_ptr(), n, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        print(json.dumps({"log_n": log_n, "pad_kb": int(os.environ.get("MIRA_ACC_PAD_KB", "0")),
                          "rs_threads": int(os.environ.get("MIRA_RS_THREADS_RT", "512")), "prep_prio": int(os.environ.get("MIRA_PREP_PRIO", "0")),
                          "slices": slices, "ms": round(e0.elapsed_time(e1) / reps, 3)}), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    grid = [(0, 512, 0), (0, 256, 0), (0, 256, 1), (0, 512, 1), (57, 256, 1), (57, 256, 0), (57, 512, 1), (44, 256, 1)]
    for pad, rs, prio in grid:
        env = dict(os.environ, MIRA_ACC_PAD_KB=str(pad), MIRA_RS_THREADS_RT=str(rs), MIRA_PREP_PRIO=str(prio))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)


if __name__ == "__main__":
    main()

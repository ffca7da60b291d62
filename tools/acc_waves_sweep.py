"""Chunk length of k_accumulate (MIRA_ACC_WAVES = waves of 148 x 512 threads the list is cut into) against commit time,
single commits 2^16..2^22 and the batch of 6 x 2^19 / 6 x 2^16 a fold step issues.  One child per setting."""
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
    nmax = 1 << 22
    bases = gpu_util.gen_bases_dev(0, 1, nmax)
    ck = CommitmentKey(0, bases, on_device=True)
    stream = torch.cuda.current_stream()
    out = {"waves": os.environ.get("MIRA_ACC_WAVES") or "default", "lmin": os.environ.get("MIRA_ACC_LMIN", "32"),
           "red_min_log": os.environ.get("MIRA_RED_MIN_LOG", "default")}

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / reps, 3)
    for lg in (16, 18, 19, 20, 22):
        n = 1 << lg
        sc = gpu_util.gen_scalars_dev(0, 2, n, 0)
        out[f"single_2p{lg}"] = timed(lambda: ck.commit_device(sc.data_ptr(), n, stream.cuda_stream))
    for lg in (16, 19):
        n = 1 << lg
        T = [gpu_util.gen_scalars_dev(0, 10 + k, n, 0) for k in range(6)]
        ptrs = [t.data_ptr() for t in T]
        out[f"batch6_2p{lg}"] = timed(lambda: ck.commit_batch_device(ptrs, n, stream.cuda_stream))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for waves, lmin in (("0", "32"), ("1", "32"), ("2", "32"), ("3", "32"), ("4", "32"), ("1", "64"), ("2", "64")):
            env = dict(os.environ, MIRA_ACC_LMIN=lmin)
            if waves != "0":
                env["MIRA_ACC_WAVES"] = waves
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)

"""Maximum-size check on ONE GPU (SURVEY.md 8c "maximum sizes"): a 2^26-point BN254 MSM and the TensorStar witness
shape (58,720,256 scalars) committed whole, and again as two point-range shards folded with mira_msm_combine — the two
routes share no bucket, sort or table state, so agreement at this size is a size-independent parity property."""
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch

import gpu_util
from mira_b200 import CommitmentKey, combine_partials

for n, dist, label in ((1 << 26, 0, "2^26 uniform"), (58_720_256, 1, "zkml W (14 * 2^22) witness-like")):
    bases = gpu_util.gen_bases_dev(0, 7, n)
    sc = gpu_util.gen_scalars_dev(0, 8, n, dist)
    ck = CommitmentKey(0, bases, on_device=True)
    t0 = time.time(); ck.prepare(n); torch.cuda.synchronize(); prep = time.time() - t0
    ck.commit_device(sc.data_ptr(), n)
    ts = []
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time(); whole = ck.commit_device(sc.data_ptr(), n); ts.append((time.time() - t0) * 1e3)
    st = ck.stats()
    ck.close(); del ck
    torch.cuda.empty_cache()
    half = n // 2
    parts = b""
    for lo, hi in ((0, half), (half, n)):
        sub = CommitmentKey(0, bases[lo * 64:hi * 64], on_device=True)
        parts += sub.partial(sc[lo * 32:hi * 32].data_ptr(), hi - lo, on_device=True)
        sub.close(); del sub
        torch.cuda.empty_cache()
    sharded = combine_partials(0, parts)
    print(json.dumps({"case": label, "n": n, "c": st["window_bits"], "W": st["windows"], "prep_s": round(prep, 2),
                      "ms": [round(t, 1) for t in ts], "Mpts_s": round(n / min(ts) / 1e3, 1), "whole_equals_sharded": whole == sharded,
                      "commit": whole.hex()[:32]}), flush=True)
    assert whole == sharded
    del bases, sc
    torch.cuda.empty_cache()

"""Ad-hoc phase timing of the MSM on one GPU (development aid; bench.py is the contract)."""
import sys, time, json
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
from mira_b200 import CommitmentKey
import gpu_util

curve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
logs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 20, 22, 24]
wins = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
dist = int(sys.argv[4]) if len(sys.argv) > 4 else 0
nmax = 1 << max(logs)
t0 = time.time()
bases = gpu_util.gen_bases_dev(curve, 1, nmax)
torch.cuda.synchronize()
print(f"gen_bases 2^{max(logs)}: {time.time()-t0:.2f}s", flush=True)
ck = CommitmentKey(curve, bases, on_device=True)
for lg in logs:
    n = 1 << lg
    sc = gpu_util.gen_scalars_dev(curve, 2, n, dist)
    for c in wins:
        ck.set_window(c)
        t0 = time.time(); ck.prepare(n); torch.cuda.synchronize(); tprep = time.time() - t0
        ck.set_profiling(False)
        r0 = ck.commit_device(sc.data_ptr(), n)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            r = ck.commit_device(sc.data_ptr(), n)
            ts.append((time.time() - t0) * 1e3)
            assert r == r0
        ck.set_profiling(True)
        ck.commit_device(sc.data_ptr(), n)
        st = ck.stats()
        print(json.dumps({"curve": curve, "log_n": lg, "dist": dist, "c": st["window_bits"], "W": st["windows"], "prep_s": round(tprep, 2),
                          "wall_ms": [round(t, 2) for t in ts], "Mpts_s": round(n / min(ts) / 1e3, 1),
                          "digits": round(st["ms_digits"], 3), "sort": round(st["ms_sort"], 3), "acc": round(st["ms_accumulate"], 3),
                          "reduce": round(st["ms_reduce"], 3), "total": round(st["ms_total"], 3)}), flush=True)

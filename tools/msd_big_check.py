"""2^25 / 2^26-point commits through the MSD partition with 2^18 groups: prints the commitment so that two runs
(default, and MIRA_SORT_MSD=0 = LSD passes) can be compared, and the phase times.  python tools/msd_big_check.py LOG_N"""
import sys, json, hashlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from mira_b200 import CommitmentKey
import gpu_util
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = 1 << lg
bases = gpu_util.gen_bases_dev(0, 11, n)
ck = CommitmentKey(0, bases, on_device=True)
sc = gpu_util.gen_scalars_dev(0, 12, n, 0)
ck.prepare(n)
r = ck.commit_device(sc.data_ptr(), n)
ck.set_profiling(True)
ck.commit_device(sc.data_ptr(), n)
st = ck.stats()
print(json.dumps({"log_n": lg, "commitment_sha256": hashlib.sha256(r).hexdigest(), "c": st["window_bits"], "digits": round(st["ms_digits"], 3),
                  "sort": round(st["ms_sort"], 3), "acc": round(st["ms_accumulate"], 3), "reduce": round(st["ms_reduce"], 3), "total": round(st["ms_total"], 3)}))

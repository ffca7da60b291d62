"""The batched cross-term commit of a fold step (6 x 2^k uniform vectors against one key) in isolation, for ncu launch
lists:  ncu --metrics gpu__time_duration.sum ... python tools/batch_profile.py [log_rows] [count]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, gpu_util
from mira_b200 import CommitmentKey
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 19
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rows = 1 << lg
bases = gpu_util.gen_bases_dev(0, 1, rows)
ck = CommitmentKey(0, bases, on_device=True)
T = [gpu_util.gen_scalars_dev(0, 10 + k, rows, 0) for k in range(cnt)]
ptrs = [t.data_ptr() for t in T]
for _ in range(3):
    r = ck.commit_batch_device(ptrs, rows)
torch.cuda.synchronize()
st = ck.stats()
print("c", st["window_bits"], "W", st["windows"], "launches", st["kernel_launches"])

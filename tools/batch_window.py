import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, gpu_util
from mira_b200 import CommitmentKey
rows = 1 << 19
bases = gpu_util.gen_bases_dev(0, 1, rows)
ck = CommitmentKey(0, bases, on_device=True)
T = [gpu_util.gen_scalars_dev(0, 10 + k, rows, 0) for k in range(6)]
ptrs = [t.data_ptr() for t in T]
for c in (0, 16, 17, 18, 19, 20):
    ck.set_window(c)
    ref = ck.commit_batch_device(ptrs, rows)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.time(); r = ck.commit_batch_device(ptrs, rows); ts.append((time.time() - t0) * 1e3)
    st = ck.stats()
    print("c", st["window_bits"], "W", st["windows"], "batch-of-6 ms", round(min(ts), 3), "launches", st["kernel_launches"])

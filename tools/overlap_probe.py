"""Probe: do the sort phase of one commit and the accumulation of another overlap on the GPU?  Two contexts (2^23 points
each) driven from two host threads against the same two commits issued back to back from one thread."""
import sys, time, threading
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import gpu_util
from mira_b200 import CommitmentKey

n = 1 << 23
cks, scs = [], []
for i in range(2):
    b = gpu_util.gen_bases_dev(0, 31 + i, n)
    ck = CommitmentKey(0, b, on_device=True); ck.prepare(n); del b
    cks.append(ck); scs.append(gpu_util.gen_scalars_dev(0, 41 + i, n, 0))
want = [cks[i].commit_device(scs[i].data_ptr(), n) for i in range(2)]
torch.cuda.synchronize()
REPS = 6
t0 = time.time()
for _ in range(REPS):
    for i in range(2):
        assert cks[i].commit_device(scs[i].data_ptr(), n) == want[i]
seq = (time.time() - t0) / REPS * 1e3

def worker(i, delay):
    time.sleep(delay)
    for _ in range(REPS):
        assert cks[i].commit_device(scs[i].data_ptr(), n) == want[i]
for delay in (0.0, 0.010):
    th = [threading.Thread(target=worker, args=(i, delay * i)) for i in range(2)]
    t0 = time.time()
    for t in th: t.start()
    for t in th: t.join()
    par = (time.time() - t0 - delay) / REPS * 1e3
    print(f"two commits of 2^23: back to back {seq:.2f} ms, two threads (offset {delay*1e3:.0f} ms) {par:.2f} ms per pair")

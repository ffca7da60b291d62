"""Development aid: end-to-end commit time (pinned host scalars) against the H2D slicing threshold."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, gpu_util
from mira_b200 import CommitmentKey
n = 1 << 24
bases = gpu_util.gen_bases_dev(0, 1, n)
sc = gpu_util.gen_scalars_dev(0, 2, n, 0)
ck = CommitmentKey(0, bases, on_device=True); ck.prepare(n); del bases
host = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True); host.copy_(sc); torch.cuda.synchronize()
ref = ck.commit_device(sc.data_ptr(), n)
for smin in (0, 1 << 22, 1 << 20, 1 << 19, 1 << 17):
    ck.set_slice_min(smin)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.time(); r = ck.commit(host); ts.append((time.time() - t0) * 1e3)
        assert r == ref
    print("slice_min", smin, "e2e ms", [round(t, 2) for t in ts])

# pageable vs registered host memory (what a Rust Vec is until the shim registers it)
import ctypes, numpy as np
from mira_b200 import _native as N
ck.set_slice_min(1 << 19)
pageable = np.frombuffer(bytearray(host.numpy().tobytes()), dtype=np.uint8)
for label in ("pageable", "registered"):
    if label == "registered":
        t0 = time.time(); assert N.lib().mira_host_register(pageable.ctypes.data, pageable.nbytes) == 0
        print("mira_host_register of", pageable.nbytes >> 20, "MiB:", round((time.time() - t0) * 1e3, 1), "ms")
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.time(); r = ck.commit(pageable); ts.append((time.time() - t0) * 1e3)
        assert r == ref
    print(label, "host scalars: e2e ms", [round(t, 2) for t in ts])
N.lib().mira_host_unregister(pageable.ctypes.data)

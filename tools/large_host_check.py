"""Maximum-size check of the host-buffer path: a 2^26-point BN254 commit from PAGE-LOCKED host scalars (H2D slices, the
preparation of slice k+1 overlapped with the accumulation of slice k, two pair-list buffer sets) against the same commit
from device-resident scalars."""
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import gpu_util
from mira_b200 import CommitmentKey

n = 1 << 26
bases = gpu_util.gen_bases_dev(0, 7, n)
sc = gpu_util.gen_scalars_dev(0, 8, n, 0)
ck = CommitmentKey(0, bases, on_device=True); del bases
ck.prepare(n)
dev = ck.commit_device(sc.data_ptr(), n)
host = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
host.copy_(sc); torch.cuda.synchronize()
ts = []
for _ in range(2):
    t0 = time.time(); got = ck.commit(host); ts.append(round((time.time() - t0) * 1e3, 1))
    assert got == dev
free, total = torch.cuda.mem_get_info()
print(json.dumps({"case": "2^26 from page-locked host scalars", "ms": ts, "equals_device_commit": True, "commit": dev[:16].hex(),
                  "device_memory_in_use_GB": round((total - free) / 1e9, 1)}))

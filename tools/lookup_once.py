"""One evaluate_m + evaluate_h_g at 2^22 elements (for an ncu capture of the lookup kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util
from mira_b200 import witness as W

FR, k = 1, 22
l, t = gpu_util.gen_scalars_dev(0, 70, 1 << k, 0), gpu_util.gen_scalars_dev(0, 71, 1 << k, 0)
r = gpu_util.to_bytes(gpu_util.gen_scalars_dev(0, 3, 1, 0))
for _ in range(2):
    m = W.evaluate_m(FR, l, t)
    h, g = W.evaluate_h_g(FR, l, t, r, m)
torch.cuda.synchronize()
print("lookup done")

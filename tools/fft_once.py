"""One FFT per size on the GPU (for an ncu launch list of the FFT kernels): python tools/fft_once.py [log_n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mira_b200 import witness as W
FR = 1  # MIRA_FR

for k in [int(v) for v in sys.argv[1:]] or [24]:
    g = torch.Generator(device="cuda").manual_seed(k)
    x = torch.randint(0, 1 << 28, (1 << k, 8), dtype=torch.int32, device="cuda", generator=g).view(torch.uint8).reshape(-1)
    for _ in range(2):
        W.fft(FR, x, k)
    torch.cuda.synchronize()
    print("fft", k, "done")

"""Small driver for ncu: a few launches of the fused kernel on a uniform 256 x 18 s batch (B200FE_WS=1 profiles the experimental kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
B2, N2 = 256, 16000 * 18
wav2 = (torch.randn((B2, N2), device=dev) * 0.1).clamp_(-1, 1)
n2 = np.full(B2, N2, dtype=np.int64)
T2 = 1 + (N2 - 400) // 160
out = torch.empty((B2, T2, 80), device=dev)
fe = lasr_b200.GpuFbankFrontend()
for _ in range(3):
    fe(wav2, n2, out=out)
torch.cuda.synchronize()

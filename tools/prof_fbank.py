"""Small driver for ncu: a few launches of the fused kernel on a C2-shaped batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
mode = sys.argv[1] if len(sys.argv) > 1 else "none"
B = 256
rng = np.random.default_rng(1)
dur = rng.uniform(1, 35, B)
n = (dur * 16000).astype(np.int64)
nmax = int((n.max() + 3) // 4 * 4)
wav = (torch.randn((B, nmax), device="cuda") * 0.1).clamp_(-1, 1)
fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    fe(wav, n)
torch.cuda.synchronize()
print("done")

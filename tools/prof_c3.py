"""Small driver for ncu launch lists of the BASELINE config 3 variants: python tools/prof_c3.py [mean|warp] [iterations]"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200

mode = sys.argv[1] if len(sys.argv) > 1 else "warp"
dev = "cuda:0"
B, N = 512, 160000
g = torch.Generator(device=dev); g.manual_seed(2)
wav = (torch.randn((B, N), device=dev, generator=g) * 0.1).clamp_(-1, 1)
n = np.full(B, N, dtype=np.int64)
stats = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
fe = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=stats, specaug=True, time_warp=(mode == "warp"))
random.seed(2); np.random.seed(2)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    fe(wav, n)
torch.cuda.synchronize()
print("done")

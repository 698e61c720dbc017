import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from tools import bench_configs as bc
dev = torch.device("cuda:0")
wavs = bc.c1_inputs()
fe = lasr_b200.GpuFbankFrontend()
wav = torch.from_numpy(np.stack(wavs).astype(np.float32)).to(dev)
nd = torch.full((16,), 160000, dtype=torch.int64, device=dev)
nh = np.full(16, 160000, dtype=np.int64)
out = torch.empty((16, 998, 80), dtype=torch.float32, device=dev)
ol = torch.empty((16,), dtype=torch.int64, device=dev)
for name, n in (("device lengths (list builder)", nd), ("host lengths (full grid)", nh)):
    fe.launch_count = 0
    ms = bc._timed(lambda: fe(wav, n, max_frames=998, out=out, out_len=ol), 50, 5, dev)
    print(name, round(ms * 1e3, 1), "us per step")

"""cProfile of B200Collate calls (C2 float64 lists): where the interpreter's time goes inside the plug-in call."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_list
lists = [make_list(s)[0] for s in (1, 101, 201, 301)]
col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar")
for i in range(6):
    col(lists[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    col(lists[i % 4])
print("ms per call: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    col(lists[i % 4])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())

"""Timing sweep of the front end on the C2 (ragged) and a uniform batch: fused-kernel time vs step time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_batch, algorithmic_bytes

dev = "cuda:0"
def run(name, wav, n, K=20, **kw):
    fe = lasr_b200.GpuFbankFrontend(**kw)
    T, _ = fe.frame_counts(n)
    out = torch.empty((len(n), int(T.max()), 80), device=dev); ol = torch.empty((len(n),), dtype=torch.int64, device=dev)
    for _ in range(3): fe(wav, n, out=out, out_len=ol)
    torch.cuda.synchronize()
    fe.profile_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fe(wav, n, out=out, out_len=ol)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    fused = sum(a.elapsed_time(b) for a, b in fe.profile_events) / K
    nl = len(fe.profile_events) // K
    ab = algorithmic_bytes(n, T, len(n))
    print("%-44s step %.3f ms  fused %.3f ms (%d launches)  alg GB/s on fused %.0f  ns/frame %.3f" % (name, ms, fused, nl, ab / fused / 1e6, fused * 1e6 / T.sum()))

wav_np, n = make_batch(1)
wav = torch.from_numpy(wav_np).to(dev)
run("C2 ragged, no cmvn, 1 launch", wav, n)
run("C2 ragged, utt cmvn, 1 group", wav, n, cmvn="utt_meanvar", l2_chunk_bytes=1 << 40)
run("C2 ragged, utt cmvn, 48MB groups", wav, n, cmvn="utt_meanvar")
run("C2 ragged, utt cmvn, 96MB groups", wav, n, cmvn="utt_meanvar", l2_chunk_bytes=96 << 20)
# sorted by length (what the reference's batch_sort would feed)
order = np.argsort(n)
wav_s = wav[torch.from_numpy(order).to(dev)].contiguous(); n_s = n[order]
run("C2 sorted, utt cmvn, 48MB groups", wav_s, n_s, cmvn="utt_meanvar")
# uniform 256 x 18 s
N2 = 16000 * 18
wav2 = (torch.randn((256, N2), device=dev) * 0.1).clamp_(-1, 1); n2 = np.full(256, N2, dtype=np.int64)
run("uniform 256x18s, no cmvn", wav2, n2)
run("uniform 256x18s, utt cmvn 48MB", wav2, n2, cmvn="utt_meanvar")
# C3: 512 x 10 s global cmvn + specaug
N3 = 160000
wav3 = (torch.randn((512, N3), device=dev) * 0.1).clamp_(-1, 1); n3 = np.full(512, N3, dtype=np.int64)
st = lasr_b200.GpuFbankFrontend().accumulate_stats(wav3, n3).cpu().numpy()
run("C3 512x10s global cmvn + specaug(mean fill)", wav3, n3, cmvn="global", cmvn_stats=st, specaug=True)
run("C3 512x10s global cmvn + specaug(zero)", wav3, n3, cmvn="global", cmvn_stats=st, specaug=True, replace_with_zero=True)
run("C3 512x10s global cmvn only", wav3, n3, cmvn="global", cmvn_stats=st)

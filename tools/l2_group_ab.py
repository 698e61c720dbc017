"""C2 step (fbank + utterance CMVN, inputs in HBM) with the utterances of a call processed in L2-sized groups
(GpuFbankFrontend.l2_chunk_bytes): fused launch + post pass per group, so that the post pass finds the group's features in L2."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
lasr = importlib.import_module("lighting-asr_b200")
dev = torch.device("cuda:0")
wav_np, n_np = bench.make_batch()
wav = torch.from_numpy(wav_np).to(dev)
n = torch.from_numpy(n_np).to(dev)
Tmax = int(((n_np - 400) // 160 + 1).max())
out = {}
for chunk in (None, 100 << 20, 75 << 20, 50 << 20, 38 << 20):
    fe = lasr.GpuFbankFrontend(cmvn="utt_meanvar", l2_chunk_bytes=chunk)
    feats = torch.empty((len(n_np), Tmax, 80), dtype=torch.float32, device=dev)
    flen = torch.empty((len(n_np),), dtype=torch.int64, device=dev)
    for _ in range(3):
        fe(wav, n_np, max_frames=Tmax, out=feats, out_len=flen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fe(wav, n_np, max_frames=Tmax, out=feats, out_len=flen)
    e1.record()
    torch.cuda.synchronize()
    out[str(chunk)] = e0.elapsed_time(e1) / 20
print(json.dumps(out))

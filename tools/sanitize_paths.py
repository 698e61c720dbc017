"""One tiny call through every kernel path (a driver for compute-sanitizer where it is available; it is closed on this pool) --
plain, statistics + post pass, zero / mean-fill masks, peak norm, int16, packed in / out, padding tiles, device-built work
list, multi-stream streaming tiles (16 kHz and 8 kHz), time warp, encoder masks, ragged copy kernel."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
rng = np.random.default_rng(0)
n = np.array([401, 2000, 5500, 7777, 12000], dtype=np.int64)
wavs = [rng.uniform(-0.5, 0.5, k).astype(np.float32) for k in n]
nmax = int((n.max() + 3) // 4 * 4)
buf = np.zeros((len(n), nmax), dtype=np.float32)
for i, w in enumerate(wavs): buf[i, :len(w)] = w
wav = torch.from_numpy(buf).to(dev)
stats = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
random.seed(0); np.random.seed(0)
for kw in ({}, {"cmvn": "utt_meanvar"}, {"cmvn": "global", "cmvn_stats": stats, "specaug": True, "replace_with_zero": True},
           {"specaug": True, "cmvn": "utt_mean"}, {"peak_norm": True}, {"specaug": True, "time_warp": True}, {"num_mel_bins": 40},
           {"sample_frequency": 8000.0}):
    fe = lasr_b200.GpuFbankFrontend(**kw)
    fe(wav, n)
    if not kw.get("time_warp"):
        fe(wav, n, packed_out=True)
        fe(wav, torch.from_numpy(n).to(dev), max_frames=80)
    fe(wav[:, 1:], n - 1)
fe = lasr_b200.GpuFbankFrontend()
fe(torch.round(wav * 32767).to(torch.int16), n)
pk, lens, offs = fe.pack_host(wavs)
fe.extract_host(pk, lens, wav_offsets=offs, group_bytes=30000)
fe.extract_host(torch.from_numpy(buf).pin_memory(), n, group_bytes=30000)
fe.extract_host(pk, lens, wav_offsets=offs, packed_out=True)
for sf, ch in ((16000.0, 640), (8000.0, 320)):
    st = lasr_b200.StreamingFbank(11, device=dev, sample_frequency=sf)
    a = torch.rand((11, ch), device=dev) - 0.5
    for _ in range(4): st.push(a)
flen = torch.tensor([70, 5, 31], device=dev)
lasr_b200.mask.src_mask(flen, 70); lasr_b200.mask.subsampled_mask(flen, 70)
torch.cuda.synchronize()
print("sanitize paths done")

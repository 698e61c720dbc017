"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
with empty stub modules for the absent librosa / soundfile) together with the image's
torchaudio 2.11.0.  TEST INFRASTRUCTURE; run in the build container only:

    python -m oracle.gen_golden

The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    for name in ("librosa", "soundfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import lasr.data.datatrans as dt           # noqa: E402
    import lasr.utils.specaugment as sa        # noqa: E402
    from lasr.data.dataset import batch_list   # noqa: E402
    return dt, sa, batch_list


def pattern(T, D=80):
    t = np.arange(T, dtype=np.float64)[:, None]
    d = np.arange(D, dtype=np.float64)[None, :]
    return (10.0 + 3.0 * np.sin(0.37 * t + 0.11 * d) + 0.05 * d - 0.002 * t).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    dt, sa, batch_list = import_reference()
    reg = dt.register_trans
    import torch
    import torchaudio

    # ---- fbank:80 / norm on short seeded waveforms (what soundfile.read would return: float64) ----
    rng = np.random.default_rng(1234)
    fb = {}
    for i, n in enumerate((400, 559, 560, 4000, 8000, 12345)):
        kind = i % 3
        if kind == 0:
            w = rng.uniform(-0.5, 0.5, n)
        elif kind == 1:
            w = np.clip(rng.normal(0, 0.1, n), -1, 1)
        else:  # speech-like: decaying spectrum + noise, quantised like int16 PCM
            t = np.arange(n) / 16000.0
            w = 0.3 * np.sin(2 * np.pi * 180 * t) + 0.1 * np.sin(2 * np.pi * 1230 * t + 1.0) + 0.02 * rng.standard_normal(n)
            w = np.round(w * 32768.0) / 32768.0
        fb["wav_%d" % i] = w
        fb["fbank_%d" % i] = reg["fbank:80"](w)
        fb["norm_%d" % i] = reg["norm"](w)
        fb["norm_fbank_%d" % i] = reg["fbank:80"](reg["norm"](w))
        # subtract_mean is the only CMVN-like option on the reference's call path (datatrans.py:62)
        fb["fbank_cms_%d" % i] = dt.WavToKaldiFbank(w, subtract_mean=True)
    fb["batch_list"] = batch_list([fb["fbank_%d" % i] for i in range(6)], pad_value=0)
    fb["versions"] = np.array([torch.__version__, torchaudio.__version__, np.__version__])
    np.savez_compressed(os.path.join(OUT, "fbank_reference.npz"), **fb)

    # ---- SpecAugment masks (freq_mask then time_mask exactly as datatrans.py:137-150 calls them) ----
    sg = {}
    cases = []
    for seed in range(6):
        for T in (5, 12, 39, 98, 300):
            x = pattern(T)
            random.seed(seed)
            np.random.seed(seed)
            y = x.copy()
            y = sa.freq_mask(y, 27, 2, inplace=True, replace_with_zero=False)
            y = sa.time_mask(y, 40, 2, inplace=True, replace_with_zero=False)
            after = np.array([random.random(), np.random.rand()])
            key = "s%d_T%d" % (seed, T)
            sg[key + "_out"] = y
            sg[key + "_rng_after"] = after
            # replace_with_zero variant consumes the same draws
            random.seed(seed)
            np.random.seed(seed)
            z = x.copy()
            z = sa.freq_mask(z, 27, 2, inplace=True, replace_with_zero=True)
            z = sa.time_mask(z, 40, 2, inplace=True, replace_with_zero=True)
            sg[key + "_zero_changed"] = np.packbits(z != x)
            # full registry transform (time warp first): only the RNG position afterwards is recorded
            random.seed(seed)
            np.random.seed(seed)
            reg["specaug"](x.copy())
            sg[key + "_rng_after_full"] = np.array([random.random(), np.random.rand()])
            cases.append(key)
    # full registry transform outputs (time warp + masks) for the time-warp row F1 (smaller set: the arrays are stored)
    full_cases = []
    for seed in range(4):
        for T in (11, 12, 39, 98, 300):
            x = pattern(T)
            random.seed(seed)
            np.random.seed(seed)
            sg["full_s%d_T%d" % (seed, T)] = reg["specaug"](x.copy())
            full_cases.append("s%d_T%d" % (seed, T))
    sg["full_cases"] = np.array(full_cases)
    # the resize primitive itself: PIL BICUBIC on a float32 image, rows only
    from PIL import Image
    rs = np.random.RandomState(7)
    pil_cases = []
    for h, o in ((5, 1), (5, 10), (7, 12), (30, 26), (30, 35), (200, 203), (6, 1), (1, 4)):
        img = rs.normal(10, 5, (h, 80)).astype(np.float32)
        sg["pil_in_%d_%d" % (h, o)] = img
        sg["pil_out_%d_%d" % (h, o)] = np.asarray(Image.fromarray(img).resize((80, o), Image.BICUBIC))
        pil_cases.append("%d_%d" % (h, o))
    sg["pil_cases"] = np.array(pil_cases)
    sg["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "specaug_reference.npz"), **sg)
    # ---- encoder masks: make_pad_mask / Conv2dSubsampling's mask slicing / subfunction (row F2) ----
    from lasr.utils.mask import make_pad_mask
    mk = {}
    mcases = []
    rs = np.random.RandomState(3)
    for ci, (lens, T) in enumerate((([5, 3, 2], 5), ([1, 7, 8, 6], 8), ([998] * 3 + [1], 998), (list(rs.randint(1, 3499, 16)), 3498),
                                    ([11, 12, 13, 14, 15, 16], 16), ([6, 2], 6), ([3498, 98], 3600))):
        xlen = torch.tensor(lens)
        src = (~make_pad_mask(xlen.tolist(), max_length=T)).unsqueeze(-2)
        sub = src[:, :, :-2:2][:, :, :-2:2]
        mk["c%d_len" % ci] = np.asarray(lens, dtype=np.int64)
        mk["c%d_T" % ci] = np.asarray(T)
        mk["c%d_pad" % ci] = make_pad_mask(xlen.tolist(), max_length=T).numpy()
        mk["c%d_src" % ci] = src.numpy()
        mk["c%d_sub" % ci] = sub.numpy()
        mk["c%d_hslen" % ci] = torch.sum(sub.byte(), dim=-1).squeeze(-1).numpy()
        mcases.append(ci)
    mk["cases"] = np.asarray(mcases)
    np.savez_compressed(os.path.join(OUT, "mask_reference.npz"), **mk)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__" and not any(f in sys.argv for f in ("--batching", "--cmvn", "--options")):
    main()


def gen_batching():
    """Batch formation goldens: BatchAudioDataSet.check_dataset's shuffle / stable sort / filter and
    make_batch_size / make_batch_duration (R/lasr/data/dataset.py:260-305) run UNMODIFIED on synthetic (wav_len, token_len)
    items; only the parent's file-reading check_dataset (dataset.py:106-133: audio durations, tokenizer) is skipped."""
    import_reference()
    import lasr.data.dataset as ds
    rng = np.random.default_rng(77)
    out = {}
    cases = [dict(batch_type="duration", batch_duration=120, batch_sort=True), dict(batch_type="size", batch_size=7, batch_sort=True),
             dict(batch_type="duration", batch_duration=40, batch_sort=False), dict(batch_type="size", batch_size=5, batch_sort=True, min_token=2, text_freq=0.2)]
    parent = ds.AudioDataSet.check_dataset
    ds.AudioDataSet.check_dataset = lambda self: None
    try:
        for ci, kw in enumerate(cases):
            n = 60 + 11 * ci
            wav_len = np.round(rng.uniform(0.1, 33.0, n), 2)
            wav_len[9::9] = wav_len[1:-8:9][: len(wav_len[9::9])]      # ties: the stable sort keeps the shuffled order
            token_len = rng.integers(0, 60, n)
            obj = object.__new__(ds.BatchAudioDataSet)
            obj.train_set = [{"id": i, "wav_len": float(wav_len[i]), "token_len": int(token_len[i])} for i in range(n)]
            full = dict(batch_sort=True, batch_size=32, batch_duration=320, batch_type="size", max_duration=30, min_duration=0.3, text_freq=0.08,
                        min_token=0, max_token=5000)
            full.update(kw)
            for k, v in full.items():
                setattr(obj, k, v)
            random.seed(100 + ci)
            obj.check_dataset()
            flat = np.array([it["id"] for b in obj.train_set for it in b], dtype=np.int64)
            sizes = np.array([len(b) for b in obj.train_set], dtype=np.int64)
            out["c%d_wav_len" % ci] = wav_len
            out["c%d_token_len" % ci] = token_len
            out["c%d_flat" % ci] = flat
            out["c%d_sizes" % ci] = sizes
            out["c%d_kw" % ci] = np.array(repr(full))
            out["c%d_after" % ci] = np.array(random.random())           # the global generator's position afterwards
    finally:
        ds.AudioDataSet.check_dataset = parent
    out["cases"] = np.arange(len(cases))
    np.savez_compressed(os.path.join(OUT, "batch_reference.npz"), **out)
    print("wrote batch_reference.npz")


if __name__ == "__main__" and "--batching" in sys.argv:
    gen_batching()


def gen_cmvn():
    """CMVN goldens.  The reference has no CMVN code of its own, but the library it delegates all feature
    arithmetic to does: torchaudio.functional.sliding_window_cmn is torchaudio's port of Kaldi's apply-cmvn-sliding
    (TA functional/functional.py, `sliding_window_cmn`).  With center=True and a window longer than twice the matrix it
    normalises every frame with the statistics of ALL frames, i.e. utterance CMVN (norm_vars=False: mean only,
    norm_vars=True: mean and population variance); applied to the CONCATENATION of a corpus' feature matrices it is
    global CMVN from the corpus' accumulated statistics.  Inputs are the reference's own `fbank:80` outputs of
    fbank_reference.npz; evaluated in float64."""
    import torch
    import torchaudio
    import torchaudio.functional as F
    fb = np.load(os.path.join(OUT, "fbank_reference.npz"))
    feats = [fb["fbank_%d" % i] for i in range(6)]
    out = {}

    def cmn(x, norm_vars):
        t = torch.from_numpy(np.asarray(x, dtype=np.float64))[None]
        return F.sliding_window_cmn(t, cmn_window=2 * t.shape[1] + 1000, min_cmn_window=1, center=True, norm_vars=norm_vars)[0].numpy()

    for i, x in enumerate(feats):
        out["utt_mean_%d" % i] = cmn(x, False)
        if x.shape[0] > 1:          # one frame: zero variance, sliding_window_cmn yields 0 * inf
            out["utt_meanvar_%d" % i] = cmn(x, True)
    cat = np.concatenate(feats, axis=0)
    bounds = np.cumsum([0] + [x.shape[0] for x in feats])
    for nv, key in ((False, "global_mean"), (True, "global_meanvar")):
        y = cmn(cat, nv)
        for i in range(6):
            out["%s_%d" % (key, i)] = y[bounds[i]:bounds[i + 1]]
    out["versions"] = np.array([torch.__version__, torchaudio.__version__, np.__version__])
    np.savez_compressed(os.path.join(OUT, "cmvn_reference.npz"), **out)
    print("wrote cmvn_reference.npz")


if __name__ == "__main__" and "--cmvn" in sys.argv:
    gen_cmvn()


OPTION_SETS = {
    "sf8000": dict(sample_frequency=8000.0),                                   # BASELINE config 5: the SwitchBoard shape (200 / 80 samples, 256-point FFT)
    "mel40": dict(num_mel_bins=40),
    "win20": dict(frame_length=20.0),
    "magnitude": dict(use_power=False),
    "linear": dict(use_log_fbank=False),
    "hamming_nodc": dict(window_type="hamming", remove_dc_offset=False),
    "pre0_hf": dict(preemphasis_coefficient=0.0, high_freq=-200.0, low_freq=60.0),
    "bit24": dict(audio_bit=24),
    "sf8000_mel40_shift5": dict(sample_frequency=8000.0, num_mel_bins=40, frame_shift=5.0),
}


def gen_options():
    """fbank goldens for the other option sets of the reference's WavToKaldiFbank (datatrans.py:43-71), the 8 kHz family of BASELINE
    config 5 among them: the UNMODIFIED reference function called with keyword arguments, as a config.yaml `kwargs` block would."""
    import torch
    import torchaudio
    dt, _, _ = import_reference()
    rng = np.random.default_rng(4321)
    out = {}
    for i, n in enumerate((4000, 9001)):
        out["wav_%d" % i] = np.clip(rng.normal(0, 0.1, n), -1, 1) if i == 0 else np.round(rng.uniform(-0.4, 0.4, n) * 32768.0) / 32768.0
    for name, kw in OPTION_SETS.items():
        for i in range(2):
            out["%s_%d" % (name, i)] = dt.WavToKaldiFbank(out["wav_%d" % i], **kw)
    out["names"] = np.array(sorted(OPTION_SETS))
    out["versions"] = np.array([torch.__version__, torchaudio.__version__, np.__version__])
    np.savez_compressed(os.path.join(OUT, "fbank_options_reference.npz"), **out)
    print("wrote fbank_options_reference.npz")


if __name__ == "__main__" and "--options" in sys.argv:
    gen_options()

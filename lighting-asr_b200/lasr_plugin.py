"""Plug-in points into the unmodified reference.

1. Transform registry (``lasr.data.datatrans.register_trans``, a ``Register`` whose
   ``register(name)`` decorator overrides an existing key after a warning,
   lasr/utils/register.py:6-18).  Transforms are called as ``fn(array)`` with exactly one
   positional argument and no kwargs (lasr/data/dataset.py:196-197, lasr/process/asrprocess.py:54-55).
   ``install(register_trans)`` overrides ``fbank:80`` and adds fused keys.
2. Config-named classes (``{name: "module:Class", kwargs: {...}}`` -> ``BaseConfig`` ->
   ``Class(**kwargs)``; every YAML key must be a named parameter of ``__init__``,
   lasr/utils/generater.py:91-99).  ``B200Collate`` / ``make_dataset_class`` provide the batched
   collate that runs the whole batch through one fused launch in the main process.

CUDA must not be touched inside forked DataLoader workers (bin/train_lighting.py:228): use
``num_workers=0`` (the GPU front end replaces the 16 CPU workers) or the ASRProcess path.
"""
import numpy as np
import torch

from .frontend import GpuFbankFrontend


class GpuTransform:
    """Registry-compatible callable: 1-D waveform ndarray -> (T, 80) features.

    ``return_tensor=False`` returns a float32 ndarray like ``WavToKaldiFbank`` does
    (datatrans.py:104) so that ``batch_list`` can stack it; ``True`` returns the CUDA tensor,
    which ``ASRProcess.model_forward`` accepts through ``torch.as_tensor`` (asrprocess.py:63)."""

    def __init__(self, device="cuda:0", return_tensor=False, **frontend_kwargs):
        self.device = torch.device(device)
        self.return_tensor = return_tensor
        self.frontend = GpuFbankFrontend(**frontend_kwargs)
        self._pin = None                # grow-only pinned staging: the upload is then a true asynchronous DMA
        self._pin_free = None
        # Register.register prints `value.__name__` when it overrides a key (lasr/utils/register.py:10-11)
        self.__name__ = "GpuTransform"

    def __call__(self, wav):
        w = np.asarray(wav)
        if w.ndim != 1:
            raise ValueError("expected a mono 1-D waveform (run 'avgchannel' first, datatrans.py:10-14)")
        n = w.shape[0]
        i16 = w.dtype == np.int16
        npad = (n + 7) // 8 * 8
        tdt = torch.int16 if i16 else torch.float32
        if self._pin is None or self._pin.dtype != tdt or self._pin.numel() < npad:
            if self._pin_free is not None:
                self._pin_free.synchronize()
            self._pin = torch.empty((max(2 * npad, 1 << 16),), dtype=tdt, pin_memory=True)
        elif self._pin_free is not None:
            self._pin_free.synchronize()        # the previous call's DMA has read the staging buffer
        view = self._pin.numpy()
        view[:n] = w                            # float64 -> float32 here (the `.float()` of datatrans.py:73)
        view[n:npad] = 0
        dw = self._pin[:npad].to(self.device, non_blocking=True)
        self._pin_free = torch.cuda.Event()
        self._pin_free.record(torch.cuda.current_stream(self.device))
        feats, _ = self.frontend(dw.unsqueeze(0), np.array([n], dtype=np.int64))
        feats = feats[0]
        return feats if self.return_tensor else feats.cpu().numpy()


def install(register_trans, device="cuda:0", return_tensor=False, **fbank_kwargs):
    """Overrides ``fbank:80`` in the reference's registry and adds fused chains:

    ``fbank:80``                 -> GPU fbank (same defaults as WavToKaldiFbank)
    ``b200:norm+fbank:80``       -> peak norm + fbank in one launch (replaces ["norm", "fbank:80"])
    ``b200:norm+fbank:80+specaug`` -> + the reference's SpecAugment: PIL-exact time warp, then 2 frequency / 2 time masks (mean fill)
    """
    base = dict(fbank_kwargs)
    table = {
        "fbank:80": GpuTransform(device, return_tensor, **base),
        "b200:norm+fbank:80": GpuTransform(device, return_tensor, peak_norm=True, **base),
        # the reference's `specaug` always warps first (datatrans.py:136): time_warp=True is the exact drop-in
        "b200:norm+fbank:80+specaug": GpuTransform(device, return_tensor, peak_norm=True, specaug=True, time_warp=True, **base),
    }
    for key, fn in table.items():
        register_trans.register(key)(fn)
    return table


class B200Collate:
    """Batched replacement of ``AudioDataSet.MergeBatch``'s transform loop (dataset.py:190-206).

    ``__call__(list_of_waveforms)`` -> dict(wav_array=(B,Tmax,80) float32, wav_len=(B,) int64),
    the two entries ``LightModelFace.pack_data`` reads (bin/train_lighting.py:104-126).  The list holds
    what the reference's loop receives: 1-D float64 ndarrays from ``soundfile.read`` (reader.py:24);
    float32 arrays and int16 PCM (``soundfile.read(dtype="int16")``, half the PCIe bytes) are taken as
    they are.  ``to_host=True`` returns pinned HOST tensors like the reference's collate (they are slots
    of a ring: valid until ``ring`` further calls); otherwise the features stay on the GPU (Lightning's
    batch transfer is then a no-op) and are ordered on the current stream.

    ``prefetch(iterable)`` yields the batch of item k while item k + 1 is already being packed and
    copied (what a DataLoader's prefetching does for the reference's CPU workers)."""

    def __init__(self, device="cuda:0", to_host=False, ring=3, threads=None, out_dtype=torch.float32, input_rate=None, speed_perturb=None,
                 res_type="kaiser_fast", **frontend_kwargs):
        from .host_pipeline import HostPipeline
        from . import resample as _rs
        self.device = torch.device(device)
        self.to_host = to_host
        self.frontend = GpuFbankFrontend(**frontend_kwargs)
        # out_dtype=torch.bfloat16: features in the precision of the encoder's first convolution under autocast (F2)
        self.pipeline = HostPipeline(self.frontend, self.device, ring=ring, threads=threads, out_dtype=out_dtype)
        # F3 ingest on the device: `resample:16k` (datatrans.py:16-20) when the files' rate differs from the front end's, and
        # `soxspeed` (datatrans.py:29-39) with the reference's ratio list, e.g. speed_perturb=(1, 1.1, 0.9)
        rate = self.frontend.opts["sample_frequency"]
        if input_rate is not None and float(input_rate) != float(rate):
            # res_type="kaiser_fast": librosa's filter and index arithmetic, what `resample:16k` calls (datatrans.py:19); "poly": scipy's
            self.pipeline.resampler = _rs.Resampler(int(input_rate), int(rate), res_type=res_type)
        if speed_perturb:
            self.pipeline.speed = _rs.SpeedPerturb(speed_perturb)

    def __call__(self, wavs):
        feats, flen = self.pipeline.run(wavs, to_host=self.to_host)
        return {"wav_array": feats, "wav_len": flen}

    def prefetch(self, batches):
        pending = None
        for wavs in batches:
            h = self.pipeline.submit(wavs, to_host=self.to_host)
            if pending is not None:
                feats, flen = self.pipeline.result(pending)
                yield {"wav_array": feats, "wav_len": flen}
            pending = h
        if pending is not None:
            feats, flen = self.pipeline.result(pending)
            yield {"wav_array": feats, "wav_len": flen}


def make_dataset_class():
    """Builds ``B200BatchAudioDataSet`` on top of the reference's ``BatchAudioDataSet`` (only
    possible where the ``lasr`` package and its audio I/O dependencies are importable).  Use it
    from config.yaml as ``name: "lasr_b200.lasr_plugin:B200BatchAudioDataSet"`` after calling
    this once, with ``audio_trans: [avgchannel]`` semantics: waveforms are read by the reference
    reader, everything after is one fused GPU launch per batch."""
    import lasr.data.reader as reader
    from lasr.data.dataset import BatchAudioDataSet, batch_list
    from lasr.data.datatrans import register_trans

    class B200BatchAudioDataSet(BatchAudioDataSet):
        def __init__(self, wav_list=None, text_list=None, feats_list=None, tokenizer="char", audio_trans=("avgchannel",),
                     feats_trans=None, pad_audio=0, pad_feats=0, batch_sort=True, batch_size=32, batch_duration=320,
                     batch_bin=32 * 500 * 80, batch_type="size", max_duration=30, min_duration=0.3, text_freq=0.08,
                     min_token=0, max_token=5000, device="cuda:0", peak_norm=True, cmvn="none", specaug=False, time_warp=True, pcm16=True):
            super().__init__(wav_list, text_list, feats_list, tokenizer, list(audio_trans), feats_trans, pad_audio, pad_feats,
                             batch_sort, batch_size, batch_duration, batch_bin, batch_type, max_duration, min_duration,
                             text_freq, min_token, max_token)
            self._collate = B200Collate(device, peak_norm=peak_norm, cmvn=cmvn, specaug=specaug, time_warp=time_warp)
            # pcm16: mono 16-bit files are read as int16 PCM (soundfile.read(dtype="int16")): a quarter of the host bytes of the
            # reader's float64 and half the PCIe bytes, bit-identical features ((float)s16 == float sample * 2^15 exactly)
            self._pcm16 = bool(pcm16)

        def collate_fn(self, batch):
            items = [x for b in batch for x in b]
            wavs = []
            for it in items:
                w = sr = None
                if self._pcm16 and str(it["wav"]).lower().endswith((".wav", ".flac")):
                    import soundfile
                    info = soundfile.info(it["wav"])
                    if info.subtype == "PCM_16" and info.channels == 1 and info.samplerate == 16000:
                        w, sr = soundfile.read(it["wav"], dtype="int16")
                if w is None:
                    w, sr = reader.read_audio(it["wav"])            # float64, R/lasr/data/reader.py:15-29
                    w = register_trans["avgchannel"](w)
                    if sr != 16000:
                        w = register_trans["resample:16k"](w, sr)
                wavs.append(w)
            out = {k: [it[k] for it in items] for k in items[0]}
            out.update(self._collate(wavs))
            # the dummy (B, 1, 1) feature block MergeBatch builds when no feats scp is given (dataset.py:208-215)
            out["feats_array"] = torch.from_numpy(batch_list([np.zeros((1, 1)) for _ in items], pad_value=self.pad_feats))
            out["token_id"] = torch.from_numpy(batch_list(out["token_id"], pad_value=self.tokenizer.ID_VALUE_PAD, dtype=np.int64))
            out["token_len"] = torch.from_numpy(np.array(out["token_len"], dtype=np.int64))
            return out

    globals()["B200BatchAudioDataSet"] = B200BatchAudioDataSet
    return B200BatchAudioDataSet

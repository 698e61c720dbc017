"""Chunked (streaming) front end for S concurrent streams in lock step -- BASELINE config 5.

The reference has no streaming fbank: ASRProcess.frontend (lasr/process/asrprocess.py:49-56) runs the
whole-utterance transforms, and its "online" models chunk already-computed features
(lasr/modules/net/online_transformer/encoder.py:143-176).  Correctness is therefore defined as:
the concatenation of the chunk outputs equals the offline fbank of the whole stream, frame for frame
(SURVEY.md 8(d) C5).  Here it is bit-identical, because a frame's arithmetic does not depend on which
launch or tile computes it.

State per stream: the samples the next chunk's first frames still need (between window - shift and window - 1 of
them).  All streams advance with the same chunk size, so the state is one (S, W) CUDA tensor used as a sliding
window: a push appends the chunk, launches the fused kernel on the unconsumed span IN PLACE (the C ABI takes a
base pointer and a row stride, so no data moves between pushes) and advances the start; the span is moved back to
the front only when the buffer is exhausted.  One push = one device copy + one fused launch; every stream yields
the same number of frames, so the kernel packs several streams into one 32-frame tile (`uniform_frames`).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .frontend import GpuFbankFrontend


class StreamingFbank:
    def __init__(self, n_streams, device="cuda:0", max_chunk=4096, **frontend_kwargs):
        for k in ("specaug", "peak_norm"):
            if frontend_kwargs.get(k):
                raise ValueError("%s needs the whole utterance and is not available in streaming mode" % k)
        if frontend_kwargs.get("cmvn", "none") not in ("none", "global"):
            raise ValueError("utterance CMVN needs the whole utterance; use cmvn='global' when streaming")
        frontend_kwargs.setdefault("compact_tiles", False)      # one launch per push
        self.fe = GpuFbankFrontend(**frontend_kwargs)
        self.device = torch.device(device)
        sf = self.fe.opts["sample_frequency"]
        self.win = int(sf * self.fe.opts["frame_length"] * 0.001)
        self.shift = int(sf * self.fe.opts["frame_shift"] * 0.001)
        self.S = n_streams
        self.max_chunk = max_chunk
        # sliding window: room for ~16 pushes of max_chunk (at least 64 k samples) before the span moves back to the front
        self.width = (self.win + max(16 * max_chunk, 1 << 16) + 3) // 4 * 4
        self.buf = torch.zeros((n_streams, self.width), dtype=torch.float32, device=self.device)
        self.start = 0                      # first unconsumed sample
        self.end = 0                        # one past the last valid sample
        self.frames_out = 0
        self._len_dev = {}                  # device copies of the (uniform) sample count, keyed by its value
        self._plan = self.fe.plan(self.device)
        # dither draws noise per call through the general path; everything else goes straight to the C ABI
        self._direct = self.fe.dither == 0.0

    @property
    def fill(self):
        return self.end - self.start

    def reset(self):
        self.start = self.end = 0
        self.frames_out = 0

    def _launch(self, n, T):
        """Fused launch on buf[:, start : start + n] -> (S, T, D), all arguments prebuilt (no per-push host work beyond
        the ctypes call)."""
        dev, D = self.device, self.fe.num_mel_bins
        len_dev = self._len_dev.get(n)
        if len_dev is None:
            len_dev = self._len_dev[n] = torch.full((self.S,), n, dtype=torch.int64, device=dev)
        out = torch.empty((self.S, T, D), dtype=torch.float32, device=dev)
        a = _lib.FbankArgs()
        a.d_wav = self.buf.data_ptr() + 4 * self.start
        a.wav_stride = self.width
        a.d_nsamp = len_dev.data_ptr()
        a.batch = self.S
        a.d_out = out.data_ptr()
        a.max_frames = T
        a.uniform_frames = 1
        if self.fe.cmvn == "global":
            a.d_cmvn_mean = self.fe.cmvn_mean.data_ptr()
            a.d_cmvn_istd = self.fe.cmvn_istd.data_ptr()
            a.cmvn_stride = 0
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self._plan.lib.b200fe_fbank_fused(self._plan.handle, C.byref(a), stream), "b200fe_fbank_fused")
        self.fe.launch_count += 1
        return out

    @torch.no_grad()
    def push(self, chunk):
        """chunk: float32 CUDA (S, C).  Returns (S, T_new, D) features of the frames completed by it
        (T_new may be 0)."""
        if chunk.shape[0] != self.S or chunk.dim() != 2:
            raise ValueError("chunk must be (n_streams, C)")
        Cn = chunk.shape[1]
        if Cn > self.max_chunk:
            raise ValueError("chunk larger than max_chunk")
        if self.end + Cn > self.width:
            # out of room: move the unconsumed span back to the front (once every ~16 pushes)
            rest = self.end - self.start
            self.buf[:, :rest].copy_(self.buf[:, self.start:self.end].clone())
            self.start, self.end = 0, rest
        self.buf[:, self.end:self.end + Cn].copy_(chunk)
        self.end += Cn
        n = self.end - self.start
        if n < self.win:
            return torch.empty((self.S, 0, self.fe.num_mel_bins), dtype=torch.float32, device=self.device)
        T = 1 + (n - self.win) // self.shift
        if self._direct and self.fe.cmvn_mean is not None and self.fe.cmvn_mean.device != self.device:
            self.fe.cmvn_mean = self.fe.cmvn_mean.to(self.device)
            self.fe.cmvn_istd = self.fe.cmvn_istd.to(self.device)
        if self._direct:
            feats = self._launch(n, T)
        else:
            feats, _ = self.fe(self.buf[:, self.start:self.end], np.full(self.S, n, dtype=np.int64), max_frames=T, uniform_frames=True)
        # frames of the next push start T * shift samples further on
        self.start += T * self.shift
        self.frames_out += T
        return feats

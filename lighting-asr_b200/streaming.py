"""Chunked (streaming) front end for S concurrent streams in lock step -- BASELINE config 5.

The reference has no streaming fbank: ASRProcess.frontend (lasr/process/asrprocess.py:49-56) runs the
whole-utterance transforms, and its "online" models chunk already-computed features
(lasr/modules/net/online_transformer/encoder.py:143-176).  Correctness is therefore defined as:
the concatenation of the chunk outputs equals the offline fbank of the whole stream, frame for frame
(SURVEY.md 8(d) C5).  Here it is bit-identical, because a frame's arithmetic does not depend on which
launch or tile computes it.

State per stream: the last ``N - T*shift`` samples (between window-shift and window samples) that the
next chunk's first frames still need.  All streams advance with the same chunk size, so the carry
length is uniform and the state is one (S, carry) CUDA tensor.
"""
import numpy as np
import torch

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
        width = (self.win + max_chunk + 3) // 4 * 4
        self.buf = torch.zeros((n_streams, width), dtype=torch.float32, device=self.device)
        self.fill = 0                       # valid samples currently held per stream
        self.frames_out = 0

    def reset(self):
        self.fill = 0
        self.frames_out = 0

    @torch.no_grad()
    def push(self, chunk):
        """chunk: float32 CUDA (S, C).  Returns (S, T_new, D) features of the frames completed by it
        (T_new may be 0)."""
        if chunk.shape[0] != self.S or chunk.dim() != 2:
            raise ValueError("chunk must be (n_streams, C)")
        C = chunk.shape[1]
        if self.fill + C > self.buf.shape[1]:
            raise ValueError("chunk larger than max_chunk")
        self.buf[:, self.fill:self.fill + C].copy_(chunk)
        n = self.fill + C
        if n < self.win:
            self.fill = n
            return torch.empty((self.S, 0, self.fe.num_mel_bins), dtype=torch.float32, device=self.device)
        T = 1 + (n - self.win) // self.shift
        feats, _ = self.fe(self.buf, np.full(self.S, n, dtype=np.int64), max_frames=T)
        used = T * self.shift
        rest = n - used
        # keep the tail: frames of the next push start at sample `used`
        self.buf[:, :rest].copy_(self.buf[:, used:n].clone())
        self.fill = rest
        self.frames_out += T
        return feats

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


class IndependentStreams:
    """S independent streams behind the C ABI's ``b200fe_stream_*`` handle (include/b200fe.h): every stream keeps its own carry on
    the device, a push may feed any subset of the streams with chunks of different lengths (a decode server's requests arrive
    asynchronously, lasr/process/asrprocess.py:49-74 serves one utterance per call).

    ``push(ids, chunks)``: host-side convenience -- the chunks (list of 1-D float arrays) go through one pinned staging buffer
    and one H2D copy, the push itself is a CUDA graph of three kernels replayed per call (``graph=True``), the completed frames
    come back as a list of (T_i, D) CUDA tensors (views of the handle's output buffer, valid until the next push).
    ``push_device(...)`` is the raw device-pointer call."""

    def __init__(self, n_streams, device="cuda:0", max_chunk=4096, graph=True, **frontend_kwargs):
        for k in ("specaug", "peak_norm"):
            if frontend_kwargs.get(k):
                raise ValueError("%s needs the whole utterance and is not available in streaming mode" % k)
        if frontend_kwargs.get("cmvn", "none") not in ("none", "global"):
            raise ValueError("utterance CMVN needs the whole utterance; use cmvn='global' when streaming")
        if frontend_kwargs.get("dither", 0.0) != 0.0:
            raise ValueError("dither is not available through the stream handle")
        self.fe = GpuFbankFrontend(**frontend_kwargs)
        self.device = torch.device(device)
        self.S, self.max_chunk = int(n_streams), int(max_chunk)
        self._plan = self.fe.plan(self.device)
        self.lib = self._plan.lib
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200fe_stream_create(self._plan.handle, self.S, self.max_chunk, C.byref(h)), "b200fe_stream_create")
        self.handle = h
        self.max_frames = self.lib.b200fe_stream_max_frames(h)
        D = self.fe.num_mel_bins
        self.D = D
        dev = self.device
        # static buffers: a graph replays fixed pointers; row i of a push <-> entry i of these
        self.d_ids = torch.zeros((self.S,), dtype=torch.int32, device=dev)
        self.d_len = torch.zeros((self.S,), dtype=torch.int32, device=dev)
        self.d_chunks = torch.zeros((self.S, self.max_chunk), dtype=torch.float32, device=dev)
        self.d_out = torch.zeros((self.S, max(self.max_frames, 1), D), dtype=torch.float32, device=dev)
        self.d_frames = torch.zeros((self.S,), dtype=torch.int64, device=dev)
        self.h_meta = torch.zeros((2, self.S), dtype=torch.int32, pin_memory=True)          # ids | lengths
        self.h_chunks = torch.zeros((self.S, self.max_chunk), dtype=torch.float32, pin_memory=True)
        self.h_frames = torch.zeros((self.S,), dtype=torch.int64, pin_memory=True)
        self.d_meta = torch.zeros((2, self.S), dtype=torch.int32, device=dev)
        self.use_graph = bool(graph)
        self._graphs = {}
        if self.fe.cmvn == "global":
            self.fe.cmvn_mean = self.fe.cmvn_mean.to(dev)
            self.fe.cmvn_istd = self.fe.cmvn_istd.to(dev)
        # one empty push (zero-length chunk: no state changes) loads the three kernels outside any graph capture
        self.push_device(1, self.d_meta[0].data_ptr(), self.d_chunks.data_ptr(), self.max_chunk, self.d_meta[1].data_ptr(), self.d_out.data_ptr(),
                         self.d_out.shape[1], self.d_frames.data_ptr())
        torch.cuda.current_stream(dev).synchronize()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200fe_stream_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass

    def push_device(self, n, d_ids, d_chunks, chunk_stride, d_len, d_out, max_out_frames, d_frames, stream=None):
        cm = self.fe.cmvn_mean.data_ptr() if self.fe.cmvn == "global" else 0
        ci = self.fe.cmvn_istd.data_ptr() if self.fe.cmvn == "global" else 0
        st = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.b200fe_stream_push(self.handle, C.c_void_p(d_ids), int(n), C.c_void_p(d_chunks), int(chunk_stride), C.c_void_p(d_len),
                                               C.c_void_p(cm), C.c_void_p(ci), C.c_void_p(d_out), int(max_out_frames), C.c_void_p(d_frames), C.c_void_p(st)),
                   "b200fe_stream_push")
        self.fe.launch_count += 3

    def _launch(self, n):
        """The push over rows 0 .. n-1 of the static buffers: a captured graph per n (replay), or three direct launches."""
        args = (n, self.d_meta[0].data_ptr(), self.d_chunks.data_ptr(), self.max_chunk, self.d_meta[1].data_ptr(), self.d_out.data_ptr(),
                self.d_out.shape[1], self.d_frames.data_ptr())
        if not self.use_graph:
            self.push_device(*args)
            return
        g = self._graphs.get(n)
        if g is None:
            # warm up outside the capture (module loading, attribute set-up), then capture the three launches
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self.push_device(*args, stream=side.cuda_stream)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graphs[n] = g
        g.replay()
        self.fe.launch_count += 3

    @torch.no_grad()
    def push(self, ids, chunks, sync=True):
        """ids: stream indices (each at most once); chunks: list of 1-D float arrays (len <= max_chunk, may be empty).
        Returns a list of (T_i, D) CUDA tensors, one per id."""
        n = len(ids)
        if n == 0:
            return []
        if n > self.S or len(chunks) != n or len(set(int(i) for i in ids)) != n:
            raise ValueError("one chunk per listed stream, each stream at most once per push")
        meta = self.h_meta.numpy()
        hc = self.h_chunks.numpy()
        for i, (sid, c) in enumerate(zip(ids, chunks)):
            c = np.asarray(c, dtype=np.float32).reshape(-1)
            if not 0 <= int(sid) < self.S:
                raise ValueError("stream id out of range")
            if c.shape[0] > self.max_chunk:
                raise ValueError("chunk larger than max_chunk")
            meta[0, i], meta[1, i] = int(sid), c.shape[0]
            hc[i, : c.shape[0]] = c
        self.d_meta[:, :n].copy_(self.h_meta[:, :n], non_blocking=True)
        self.d_chunks[:n].copy_(self.h_chunks[:n], non_blocking=True)
        self._launch(n)
        self.h_frames[:n].copy_(self.d_frames[:n], non_blocking=True)
        if not sync:
            return None
        torch.cuda.current_stream(self.device).synchronize()
        T = self.h_frames[:n].tolist()
        return [self.d_out[i, : int(t)] for i, t in enumerate(T)]

    def reset(self, ids=None):
        if ids is None:
            _lib.check(self.lib.b200fe_stream_reset(self.handle, C.c_void_p(0), self.S, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "b200fe_stream_reset")
            return
        t = torch.as_tensor(list(ids), dtype=torch.int32).to(self.device)
        _lib.check(self.lib.b200fe_stream_reset(self.handle, C.c_void_p(t.data_ptr()), int(t.numel()), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "b200fe_stream_reset")
        torch.cuda.current_stream(self.device).synchronize()

    def flags(self):
        f = C.c_int(0)
        _lib.check(self.lib.b200fe_stream_flags(self.handle, C.byref(f), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "b200fe_stream_flags")
        return f.value

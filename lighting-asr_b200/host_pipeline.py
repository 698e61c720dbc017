"""Host-to-host pipeline behind ``B200Collate``: a LIST of 1-D host waveforms in, the reference's batch tensors out.

The reference's collate loop receives the utterances one by one as float64 ndarrays from ``soundfile.read``
(lasr/data/reader.py:24, lasr/data/dataset.py:190-206), runs the transform chain per utterance and pads with
``batch_list`` (dataset.py:8-22).  Here one call moves the whole batch:

    list of ndarrays --(C thread pool: pack + float64->float32, non-temporal stores)--> pinned staging ring
      --(one DMA per ~32 MB utterance group, stream s_in)--> packed device buffer
      --(fused fbank / CMVN / SpecAugment launches, current stream)--> (B, Tmax, D) device features
      --(one copy kernel per group writing mapped pinned memory, stream s_out)--> pinned (B, Tmax, D) host batch
    padding rows of the host batch are zero-filled by the thread pool meanwhile.

All staging is capacity based and grow-only (pinned host rings, device buffers, streams are created once): batches of
different shapes reuse it, nothing is keyed on the batch geometry.  Returned host tensors are slots of a ring: a batch
stays valid until ``ring`` further calls (the trainer consumes batch k before it asks for batch k + ring).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

_SRC_CODE = {np.dtype(np.float32): 0, np.dtype(np.int16): 1, np.dtype(np.float64): 2}
_DST_TORCH = {0: torch.float32, 1: torch.int16, 2: torch.float32}


def default_threads():
    """Host threads for packing: the process's share of the cores (torchrun starts one process per GPU) MINUS ONE -- the calling
    thread issues the device work and packs too while it waits for a group, and a pool as large as the core count gets one of its
    threads preempted every few calls (measured on a 16-core box, float64 C2 lists: 15 threads 9.55 ms per call, 16 threads
    10.6 ms with a 11.2 ms p90; tools/pipe_threads.py)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(32, n // max(local, 1) - 1))


def stale_ranges(dirty, valid, view_bytes):
    """Byte ranges of a recycled batch buffer that must be cleared: what an earlier batch left there (``dirty``: sorted,
    disjoint [start, stop) pairs; everything else is known to be zero) inside the current view [0, view_bytes) minus what
    this batch overwrites (``valid``, same form).  Returns (ranges to zero, dirty set after this batch)."""
    zero, keep = [], []
    j, nv = 0, len(valid)
    for a, b in dirty:
        if a >= view_bytes:
            keep.append((a, b))
            continue
        if b > view_bytes:
            keep.append((view_bytes, b))
            b = view_bytes
        while j < nv and valid[j][1] <= a:
            j += 1
        k, cur = j, a
        while k < nv and valid[k][0] < b:
            if valid[k][0] > cur:
                zero.append((cur, valid[k][0]))
            cur = max(cur, valid[k][1])
            k += 1
        if cur < b:
            zero.append((cur, b))
    return zero, sorted(list(valid) + keep)


class _Grow:
    """Grow-only flat buffer (pinned host or device); growing synchronises the device first so that no stream still
    touches the buffer that is dropped."""

    def __init__(self, dtype, device=None, zero=False):
        self.dtype, self.device, self.buf, self.zero = dtype, device, None, zero
        self.dirty = []            # zero=True: byte ranges that may hold non-zero data (everything else is zero)

    def get(self, numel):
        if self.buf is None or self.buf.numel() < numel:
            if self.buf is not None:
                torch.cuda.synchronize()
            cap = int(numel * 1.25) + 1024
            if self.device is None:
                self.buf = (torch.zeros if self.zero else torch.empty)((cap,), dtype=self.dtype, pin_memory=True)
                self.dirty = []
            else:
                self.buf = torch.empty((cap,), dtype=self.dtype, device=self.device)
        return self.buf


class HostPipeline:
    def __init__(self, frontend, device="cuda:0", ring=3, threads=None, group_bytes=32 << 20, out_dtype=torch.float32):
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
        self.out_dtype = out_dtype          # bfloat16: converted inside the D2H copy kernel, half the D2H bytes
        self.fe = frontend
        self.dev = torch.device(device)
        self.lib = _lib.load()
        self.ring = max(2, int(ring))
        self.group_bytes = int(group_bytes)
        h = C.c_void_p()
        _lib.check(self.lib.b200fe_host_pool_create(int(threads or default_threads()), C.byref(h)), "b200fe_host_pool_create")
        self.pool = h
        self.threads = self.lib.b200fe_host_pool_threads(h)
        self._in = {}          # dst dtype -> [(_Grow pinned, _Grow device)] * 2
        self._hout = [_Grow(torch.uint8, zero=True) for _ in range(self.ring)]
        self._dout = [_Grow(torch.uint8, self.dev) for _ in range(self.ring)]
        self._dout16 = [_Grow(torch.uint8, self.dev) for _ in range(self.ring)]
        self._hlen = [torch.empty((0,), dtype=torch.int64)] * self.ring
        self._dlen = [None] * self.ring
        self._in_free = [None, None]        # H2D of the call that last used the pinned / device input slot has completed
        self._comp_done = [None, None]      # ... and its kernels no longer read the device staging buffer
        self._out_done = [None] * self.ring
        self._turn = 0
        self._streams = None
        self._events = []                   # per-group upload events, recorded by the pool thread that issues the group's DMA
        self.h2d_bytes = self.d2h_bytes = 0
        self.resampler = None               # resample.Resampler: applied on the device to every group before the fused launch (F3)
        self.speed = None                   # resample.SpeedPerturb: one numpy.random.choice per utterance, then resampling by 1 / ratio
        # "kernel": one copy kernel per group writes the pinned batch; "dma": one cudaMemcpyAsync per utterance; "kernel_end" (A/B
        # only, tools/pipe_ablate3.py): one copy kernel for the whole batch after the last group's kernels, i.e. no D2H overlap
        self.d2h_mode = os.environ.get("B200FE_D2H_MODE", "kernel")       # + "dma_block": one DMA per group over the padded block (padding rows included)
        self.taper = os.environ.get("B200FE_TAPER", "1") != "0"
        self.head_taper = os.environ.get("B200FE_HEAD_TAPER", "1") != "0"       # measured: float64 call -0.8 %, int16 call -4..7 % (profiles/r03_host_simd.txt)
        # float64 lists that HOLD 16-bit PCM values (soundfile.read of PCM_16 files, the reference's reader) go up as int16: half the
        # staging and PCIe bytes, bit-identical features.  A cheap probe decides per batch; a batch that fails the full check inside
        # the packing is repeated as float32 and the next `_pcm16_backoff` calls skip the attempt.
        self.pcm16_auto = os.environ.get("B200FE_PCM16_AUTO", "1") != "0"
        self._pcm16_skip = 0
        self._pcm16_backoff = 16
        self.pcm16_batches = 0              # batches that went up as int16 through the automatic path
        self._fast_ptrs = None              # data pointers read by the library from the ndarray objects (verified on first use)
        self.trace = None                   # set to [] to collect (label, perf_counter) stamps of every call (tools/pipe_trace.py)

    def _stamp(self, label):
        if self.trace is not None:
            import time
            self.trace.append((label, time.perf_counter()))

    def _dev_stamp(self, label, stream):
        """tracing only: a timing event on `stream`; tools/pipe_trace.py prints when the device reached it"""
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.dev_trace.append((label, ev))

    def __del__(self):
        try:
            if getattr(self, "pool", None):
                self.lib.b200fe_host_pool_destroy(self.pool)
                self.pool = None
        except Exception:  # noqa: BLE001
            pass

    def _data_pointers(self, arrs):
        """uint64 ndarray of the arrays' data addresses.  Asking every ndarray through the interpreter (``__array_interface__`` 2.2 us,
        ``.ctypes.data`` 1.8 us each) cost 0.5 ms of a 256-utterance call before the first packing job could start; the library reads
        them from the objects themselves (``id(a)`` + the offset of PyArrayObject's ``data`` field), once that has been verified
        on probe arrays of this interpreter."""
        B = len(arrs)
        if self._fast_ptrs is None and os.environ.get("B200FE_FAST_PTRS", "1") == "0":
            self._fast_ptrs = False                     # A/B runs
        if self._fast_ptrs is None:
            probe = [np.zeros(8), np.zeros(64, dtype=np.int16)[5:40], np.zeros((4, 6), dtype=np.float32)[2]]
            got = np.zeros(len(probe), dtype=np.uint64)
            ids = np.fromiter(map(id, probe), dtype=np.int64, count=len(probe))
            ok = self.lib.b200fe_host_ndarray_data(C.c_void_p(ids.ctypes.data), len(probe), C.c_void_p(got.ctypes.data), 16) == 0
            self._fast_ptrs = bool(ok and all(int(g) == p.__array_interface__["data"][0] for g, p in zip(got, probe)))
        if self._fast_ptrs and all(type(a) is np.ndarray for a in arrs):
            ids = np.fromiter(map(id, arrs), dtype=np.int64, count=B)
            out = np.empty(B, dtype=np.uint64)
            _lib.check(self.lib.b200fe_host_ndarray_data(C.c_void_p(ids.ctypes.data), B, C.c_void_p(out.ctypes.data), 16), "b200fe_host_ndarray_data")
            return out
        return np.array([a.__array_interface__["data"][0] for a in arrs], dtype=np.uint64)

    # ------------------------------------------------------------------------------------------------------------
    def submit(self, wavs, to_host=True, _allow_pcm16=True):
        """Starts one batch; returns a handle for ``result``.  Packing (host threads) is complete when this returns,
        the copies and kernels are queued on the device."""
        fe, lib, dev = self.fe, self.lib, self.dev
        B = len(wavs)
        if B == 0:
            raise ValueError("empty batch")
        self._stamp("submit")
        if self.trace is not None and not hasattr(self, "dev_trace"):
            self.dev_trace = []
        if self.trace is not None:
            self._dev_stamp("start", torch.cuda.current_stream(self.dev))
        arrs = [np.asarray(w) for w in wavs]
        if any(a.ndim == 2 for a in arrs):
            # multi-channel files: the reference's own `avgchannel` (numpy.average(wav, axis=1), datatrans.py:10-14)
            arrs = [np.average(a, axis=1) if a.ndim == 2 else a for a in arrs]
        dt = arrs[0].dtype
        if dt not in _SRC_CODE or any(a.dtype != dt for a in arrs):
            # mixed or unusual element types: everything becomes float32, int16 PCM scaled to [-1, 1) like the reader does (exact)
            dt = np.dtype(np.float32)
            arrs = [(a.astype(np.float32) * np.float32(1.0 / 32768.0)) if a.dtype == np.int16 else np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
        if any(a.ndim != 1 for a in arrs):
            raise ValueError("expected mono 1-D waveforms (run 'avgchannel' first, datatrans.py:10-14)")
        arrs = [a if a.flags.c_contiguous else np.ascontiguousarray(a) for a in arrs]
        code = _SRC_CODE[dt]
        lens = np.fromiter((a.shape[0] for a in arrs), dtype=np.int64, count=B)
        ptrs = self._data_pointers(arrs)
        pcm_try = False
        if code == 2 and self.pcm16_auto and _allow_pcm16 and self.resampler is None and self.speed is None:
            if self._pcm16_skip > 0:
                self._pcm16_skip -= 1
            else:
                pcm_try = lib.b200fe_host_pcm16_probe(C.c_void_p(ptrs.ctypes.data), C.c_void_p(lens.ctypes.data), B, 8, 64) == 1
        if pcm_try:
            code = 3                        # float64 holding PCM16 values -> int16 staging (checked while it is packed)
        ddt = torch.int16 if code == 3 else _DST_TORCH[code]
        esz = 2 if code in (1, 3) else 4
        al = 16 // esz
        lens_eff, ratios = lens, None
        if self.resampler is not None or self.speed is not None:
            if code == 1:
                raise ValueError("device resampling / speed perturbation takes float waveforms (int16 PCM lists are not resampled)")
            if self.resampler is not None:
                lens_eff = self.resampler.out_lengths(lens_eff)
            if self.speed is not None:
                ratios = self.speed.draw(B)
                lens_eff = np.array([int(self.speed._rs[r].out_lengths(n)) for r, n in zip(ratios, lens_eff)], dtype=np.int64)
        T_host, win = fe.frame_counts(lens_eff)
        if (lens_eff < win).any():
            # torchaudio asserts (TA:142); LASR filters min_duration upstream (dataset.py:243,272)
            raise AssertionError("choose a window size {} that is [2, {}]".format(win, int(lens_eff.min())))
        offs = np.zeros(B, dtype=np.int64)
        np.cumsum((lens[:-1] + al - 1) // al * al, out=offs[1:])
        total = int(offs[-1] + (lens[-1] + al - 1) // al * al)
        Tmax, D = int(T_host.max()), fe.num_mel_bins
        turn = self._turn
        self._turn += 1
        si, so = turn & 1, turn % self.ring
        if self._streams is None:
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._streams
        main = torch.cuda.current_stream(dev)
        if ddt not in self._in:
            self._in[ddt] = [(_Grow(ddt), _Grow(ddt, dev)) for _ in range(2)]
        if self._in_free[si] is not None:
            self._in_free[si].synchronize()            # the DMA that last read this pinned slot (two calls ago) is done
        hin = self._in[ddt][si][0].get(total + 64)
        dwav = self._in[ddt][si][1].get(total + 64)
        # ---- utterance groups of ~group_bytes of audio; one pack job per group, queued in order ----
        # the last groups taper (1/2, 1/4 of a group): what follows the last group's packing -- its upload, kernels and D2H -- is
        # the tail of the call that nothing overlaps
        bounds, acc = [0], 0
        csum = np.cumsum(lens * esz)
        total_b = int(csum[-1])
        taper = self.taper and total_b >= 2 * self.group_bytes      # a batch of one or two groups gains nothing from more launches
        head = self.head_taper and taper
        lens_b, csum_l = (lens * esz).tolist(), csum.tolist()      # plain ints: indexing numpy scalars cost 0.2 ms per 256 utterances
        for b in range(B):
            acc += lens_b[b]
            left = total_b - csum_l[b]
            target = self.group_bytes
            if head and len(bounds) <= 2:
                # ... and so do the first groups (1/4, 1/2): nothing is on the link until the first group is packed and nothing comes
                # back until it has gone up and through the kernels
                target = self.group_bytes // (4 if len(bounds) == 1 else 2)
            if taper:
                if left < self.group_bytes // 4:
                    target = self.group_bytes // 4
                elif left < self.group_bytes:
                    target = self.group_bytes // 2
            if acc >= target:
                bounds.append(b + 1)
                acc = 0
        if bounds[-1] != B:
            bounds.append(B)
        if self._comp_done[si] is not None:
            s_in.wait_event(self._comp_done[si])       # kernels of the call that last read this device staging buffer
        while len(self._events) < len(bounds) - 1:
            ev = torch.cuda.Event()
            ev.record(s_in)                            # materialises the cudaEvent_t the pool threads re-record
            self._events.append(ev)
        tickets = []
        self.h2d_bytes = 0
        for g, (b0, b1) in enumerate(zip(bounds[:-1], bounds[1:])):
            o0 = int(offs[b0])
            o1 = int(offs[b1]) if b1 < B else total
            # pack + convert on the pool; the thread that finishes the group issues its DMA on s_in and records the group's event
            tk = lib.b200fe_host_pack_copy_begin(self.pool, C.c_void_p(ptrs.ctypes.data + 8 * b0), C.c_void_p(lens.ctypes.data + 8 * b0), b1 - b0, code,
                                                 C.c_void_p(hin.data_ptr()), C.c_void_p(offs.ctypes.data + 8 * b0), hin.numel(),
                                                 C.c_void_p(dwav.data_ptr()), o1 - o0, dev.index or 0, C.c_void_p(s_in.cuda_stream),
                                                 C.c_void_p(self._events[g].cuda_event))
            if tk <= 0:
                _lib.check(int(tk), "b200fe_host_pack_copy_begin")
            tickets.append(tk)
            self.h2d_bytes += (o1 - o0) * esz
        # ---- host output slot: only what an earlier batch left in this batch's padding rows has to be cleared ----
        obytes = B * Tmax * D * 4
        bf16 = self.out_dtype == torch.bfloat16
        osz = 2 if bf16 else 4
        hbytes = B * Tmax * D * osz
        zero_ticket = None
        hfeats = hlen = None
        if to_host:
            hbuf = self._hout[so].get(hbytes)
            hfeats = hbuf[:hbytes].view(self.out_dtype).view(B, Tmax, D)
            row0 = np.arange(B, dtype=np.int64) * (Tmax * D * osz)
            valid = [(int(a), int(a + t * D * osz)) for a, t in zip(row0, T_host) if t > 0]
            zr, self._hout[so].dirty = stale_ranges(self._hout[so].dirty, valid, hbytes)
            if self.d2h_mode == "dma_block" and not bf16:
                zr = []                                        # the block copies bring the device's zero rows along
            if zr:
                zo = np.array([r[0] for r in zr], dtype=np.int64)
                zn = np.array([r[1] - r[0] for r in zr], dtype=np.int64)
                zero_ticket = lib.b200fe_host_zero_ranges_begin(self.pool, C.c_void_p(hbuf.data_ptr()), C.c_void_p(zo.ctypes.data), C.c_void_p(zn.ctypes.data), len(zr))
                if zero_ticket <= 0:
                    _lib.check(int(zero_ticket), "b200fe_host_zero_ranges_begin")
                self.zero_bytes = int(zn.sum())
            else:
                self.zero_bytes = 0
        # ---- device output slot ----
        dfeats = self._dout[so].get(obytes)[:obytes].view(torch.float32).view(B, Tmax, D)
        if self._dlen[so] is None or self._dlen[so].numel() < B:
            self._dlen[so] = torch.empty((max(B, 256),), dtype=torch.int64, device=dev)
        dlen = self._dlen[so][:B]
        if to_host:
            if self._hlen[so].numel() < B:
                self._hlen[so] = torch.empty((max(B, 256),), dtype=torch.int64, pin_memory=True)
            hlen = self._hlen[so][:B]
            # device-readable (offset, bytes) of every utterance's valid rows, same offsets on both sides (padded layout)
            tab = np.stack([np.arange(B, dtype=np.int64) * (Tmax * D * 4), T_host.astype(np.int64) * (D * 4),
                            np.arange(B, dtype=np.int64) * (Tmax * D * osz)])          # source row offsets | float32 bytes per row block | host row offsets
            tab_dev = torch.from_numpy(tab).to(dev, non_blocking=True)
            ev_tab = torch.cuda.Event()
            ev_tab.record(main)
            s_out.wait_event(ev_tab)
        if self._out_done[so] is not None:
            main.wait_event(self._out_done[so])        # D2H of the call that last used this device feature slot
        self.d2h_bytes = 0
        self._stamp("prepared")
        pcm_failed = False
        flag = C.c_int(0)
        for g, ((b0, b1), tk) in enumerate(zip(zip(bounds[:-1], bounds[1:]), tickets)):
            if pcm_try:
                _lib.check(lib.b200fe_host_wait_flag(self.pool, tk, C.byref(flag)), "b200fe_host_wait_flag")
                pcm_failed = pcm_failed or flag.value != 0
                if pcm_failed:
                    continue                # a sample was not a PCM16 value: nothing more is launched, the batch is repeated as float32 below
            else:
                _lib.check(lib.b200fe_host_wait(self.pool, tk), "b200fe_host_wait")      # packed, DMA issued, event recorded
            self._stamp("packed")
            self._dev_stamp("h2d %d done" % g, s_in)
            main.wait_event(self._events[g])
            gw, gl, go = dwav, lens[b0:b1], offs[b0:b1]
            if self.resampler is not None:
                gw, gl, go = self.resampler(gw, gl, go)
            if self.speed is not None:
                gw, gl, go, _ = self.speed(gw, gl, go, ratios=ratios[b0:b1])
            fe.forward(gw, gl, max_frames=Tmax, out=dfeats[b0:b1], out_len=dlen[b0:b1], wav_offsets=go)
            self._dev_stamp("compute %d done" % g, main)
            if to_host:
                ev_c = torch.cuda.Event()
                ev_c.record(main)
                s_out.wait_event(ev_c)
                if self.d2h_mode == "dma_block" and not bf16:
                    # ONE DMA per group: the group's utterances are one contiguous block of the padded layout on both sides; the
                    # padding rows travel too (zeros written by the fused launch), so the pool has nothing to clear on the host
                    with torch.cuda.stream(s_out):
                        hfeats[b0:b1].copy_(dfeats[b0:b1], non_blocking=True)
                    self.d2h_bytes += (b1 - b0) * Tmax * D * 4
                    self._dev_stamp("d2h %d done" % g, s_out)
                    self._stamp("issued")
                    continue
                if self.d2h_mode == "dma" and not bf16:
                    rows = np.ascontiguousarray(T_host[b0:b1].astype(np.int64))
                    _lib.check(lib.b200fe_d2h_ragged(C.c_void_p(dfeats.data_ptr() + b0 * Tmax * D * 4), D, Tmax, C.c_void_p(rows.ctypes.data), b1 - b0,
                                                     C.c_void_p(hfeats.data_ptr() + b0 * Tmax * D * 4), C.c_void_p(s_out.cuda_stream)), "b200fe_d2h_ragged")
                    self.d2h_bytes += int(T_host[b0:b1].sum()) * D * osz
                    self._stamp("issued")
                    continue
                copy = lib.b200fe_copy_ragged_bf16 if bf16 else lib.b200fe_copy_ragged
                if self.d2h_mode == "kernel_end" and b1 < B:
                    self._stamp("issued")
                    continue                                   # A/B: ONE copy kernel for the whole batch after the last group's kernels
                if self.d2h_mode == "kernel_end":
                    b0 = 0
                _lib.check(copy(C.c_void_p(dfeats.data_ptr()), C.c_void_p(tab_dev.data_ptr() + 8 * b0), C.c_void_p(hfeats.data_ptr()),
                                C.c_void_p(tab_dev.data_ptr() + 8 * (2 * B + b0)), C.c_void_p(tab_dev.data_ptr() + 8 * (B + b0)), b1 - b0,
                                int(T_host[b0:b1].max()) * D * 4, C.c_void_p(s_out.cuda_stream)), "b200fe_copy_ragged")
                fe.launch_count += 1
                self.d2h_bytes += int(T_host[b0:b1].sum()) * D * osz
                self._dev_stamp("d2h %d done" % g, s_out)
            self._stamp("issued")
        self._in_free[si] = torch.cuda.Event()
        self._in_free[si].record(s_in)
        self._comp_done[si] = torch.cuda.Event()
        self._comp_done[si].record(main)
        done = None
        if bf16 and not to_host:
            # the encoder consumes the batch on the device: one conversion pass over the padded batch (zero rows stay zero)
            d16 = self._dout16[so].get(hbytes)[:hbytes].view(torch.bfloat16).view(B, Tmax, D)
            _lib.check(lib.b200fe_cast_bf16(C.c_void_p(dfeats.data_ptr()), C.c_void_p(d16.data_ptr()), B * Tmax * D, C.c_void_p(main.cuda_stream)), "b200fe_cast_bf16")
            fe.launch_count += 1
            dfeats = d16
        if to_host:
            with torch.cuda.stream(s_out):
                hlen.copy_(dlen, non_blocking=True)
            self.d2h_bytes += B * 8
            done = torch.cuda.Event()
            done.record(s_out)
            self._out_done[so] = done
            tab_dev.record_stream(s_out)
        if pcm_failed:
            # the events above cover whatever the abandoned attempt queued; its ring slots are simply overwritten by later calls
            if zero_ticket:
                _lib.check(lib.b200fe_host_wait(self.pool, zero_ticket), "b200fe_host_wait")
            self._pcm16_skip = self._pcm16_backoff
            # the repeat takes the SAME ring slots (its work is queued behind the abandoned attempt's on the same streams and its
            # staging buffers are the float32 ones): a batch handed out earlier stays valid for `ring` further CALLS, as promised
            self._turn -= 1
            return self.submit(wavs, to_host, _allow_pcm16=False)
        if pcm_try:
            self.pcm16_batches += 1
        return dict(to_host=to_host, done=done, zero=zero_ticket, hfeats=hfeats, hlen=hlen, dfeats=dfeats, dlen=dlen, keep=arrs)

    def result(self, h):
        """Blocks until the batch of ``submit`` is complete on the host (``to_host``) and returns (feats, feat_len)."""
        if not h["to_host"]:
            return h["dfeats"], h["dlen"]
        self._stamp("result")
        h["done"].synchronize()
        self._stamp("device done")
        if h["zero"]:
            _lib.check(self.lib.b200fe_host_wait(self.pool, h["zero"]), "b200fe_host_wait")
            h["zero"] = None
        h["keep"] = None
        self._stamp("returned")
        return h["hfeats"], h["hlen"]

    def run(self, wavs, to_host=True):
        return self.result(self.submit(wavs, to_host))

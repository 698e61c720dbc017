"""Batched GPU front end: the drop-in for the reference's per-utterance feature extraction.

Mirrors the call signature of ``WavToKaldiFbank`` (lasr/data/datatrans.py:42-71) in its
constructor and produces the batch layout of ``AudioDataSet.MergeBatch``
(lasr/data/dataset.py:181-220): ``feats (B, Tmax, num_mel_bins) float32`` zero padded and
``feat_len (B,) int64`` frame counts.  All arithmetic runs in the C-ABI CUDA library; torch is
used for device memory and streams only.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, specaug as _specaug

_WINDOWS = {"povey": 0, "hanning": 1, "hamming": 2, "rectangular": 3, "blackman": 4}
_CMVN_MODES = ("none", "utt_mean", "utt_meanvar", "global")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_PAD_ROWS = 256       # rows zeroed by one padding tile (kPadTileRows in fbank_kernel.cuh)


def _tile_table(T, ft, Tmax=None):
    """Compact work list of the fused launch (same order as b200fe_build_tile_table[_padded]): for every utterance its frame
    tiles (utterance, first frame) and, when ``Tmax`` is given, its padding tiles (utterance, -(row0 + 1)) right after them, so
    the zero fill of the padded rows is interleaved with the frame tiles."""
    T = np.asarray(T, dtype=np.int64)
    B = len(T)
    nt = (T + ft - 1) // ft
    npad = (np.maximum(Tmax - T, 0) + _PAD_ROWS - 1) // _PAD_ROWS if Tmax is not None else np.zeros(B, dtype=np.int64)
    per = nt + npad
    tot = int(per.sum())
    tab = np.empty((tot, 2), dtype=np.int32)
    utt = np.repeat(np.arange(B, dtype=np.int64), per)
    k = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(per) - per, per)          # index of the tile inside its utterance
    ntu, Tu = nt[utt], T[utt]
    tab[:, 0] = utt
    tab[:, 1] = np.where(k < ntu, k * ft, -(Tu + (k - ntu) * _PAD_ROWS) - 1)
    return tab


_TORCH_OF = {np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
             np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}


def _h2d_many(arrs, dev):
    """ONE upload for several small host arrays (each pageable transfer costs ~80 us of host time): returns device tensors
    that are 16-byte aligned views of a single buffer, in the order given (None entries pass through)."""
    live = [(i, np.ascontiguousarray(a)) for i, a in enumerate(arrs) if a is not None]
    out = [None] * len(arrs)
    if not live:
        return out
    offs, total = [], 0
    for _, a in live:
        offs.append(total)
        total += (a.nbytes + 15) // 16 * 16
    host = np.empty(max(total, 16), dtype=np.uint8)
    for (_, a), o in zip(live, offs):
        host[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
    d = torch.from_numpy(host).to(dev, non_blocking=True)
    for (i, a), o in zip(live, offs):
        out[i] = d[o:o + a.nbytes].view(_TORCH_OF[a.dtype]).reshape(a.shape) if a.nbytes else torch.empty(a.shape, dtype=_TORCH_OF[a.dtype], device=dev)
    return out


def _h2d(arr, dev):
    """Small host -> device upload (length vectors, tile tables, mask descriptors).  Pageable on purpose: the driver
    stages < 64 kB copies inline in the compute stream's command buffer, whereas a pinned source goes through the H2D
    copy engine -- measured +15 us per step of DMA latency in front of the fused launch, and a stall behind the 32 MB
    waveform copies inside extract_host (profiles/r01_e2e_copy_variants.txt)."""
    return torch.from_numpy(np.ascontiguousarray(arr)).to(dev, non_blocking=True)


def _torch_window(window_type, size, blackman_coeff):
    """Same torch expressions as torchaudio (TA:86-113) so the table matches bit for bit."""
    import math
    if window_type == "hanning":
        return torch.hann_window(size, periodic=False)
    if window_type == "hamming":
        return torch.hamming_window(size, periodic=False, alpha=0.54, beta=0.46)
    if window_type == "povey":
        return torch.hann_window(size, periodic=False).pow(0.85)
    if window_type == "rectangular":
        return torch.ones(size)
    if window_type == "blackman":
        a = 2 * math.pi / (size - 1)
        n = torch.arange(size, dtype=torch.float32)
        return blackman_coeff - 0.5 * torch.cos(a * n) + (0.5 - blackman_coeff) * torch.cos(2 * a * n)
    raise Exception("Invalid window type " + window_type)


def _torch_mel_banks(num_bins, padded, sample_freq, low_freq, high_freq):
    """torch restatement of get_mel_banks (TA:436-511, vtln_warp == 1) evaluated with the same
    float32 tensor operations, so the weights equal torchaudio's bit for bit on this torch."""
    import math
    num_fft_bins = padded / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down)).float().contiguous()


class FbankPlan:
    """Owns a ``b200fe_plan`` (device tables for one option set on the current device)."""

    def __init__(self, num_mel_bins=80, sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0,
                 low_freq=20.0, high_freq=0.0, preemphasis_coefficient=0.97, remove_dc_offset=True,
                 use_power=True, use_log_fbank=True, window_type="povey", blackman_coeff=0.42, audio_bit=16,
                 dither=0.0, torch_tables=True):
        if window_type not in _WINDOWS:
            raise Exception("Invalid window type " + window_type)
        self.lib = _lib.load()
        o = _lib.Opts()
        self.lib.b200fe_default_opts(C.byref(o))
        o.sample_frequency = sample_frequency
        o.frame_length_ms = frame_length
        o.frame_shift_ms = frame_shift
        o.num_mel_bins = num_mel_bins
        o.low_freq = low_freq
        o.high_freq = high_freq
        o.preemphasis_coefficient = preemphasis_coefficient
        o.remove_dc_offset = int(bool(remove_dc_offset))
        o.use_power = int(bool(use_power))
        o.use_log_fbank = int(bool(use_log_fbank))
        o.window_type = _WINDOWS[window_type]
        o.blackman_coeff = blackman_coeff
        o.audio_bit = audio_bit
        o.dither = dither
        keep = []
        if torch_tables:
            win = int(sample_frequency * frame_length * 0.001)
            padded = 1 if win == 0 else 2 ** (win - 1).bit_length()
            w = _torch_window(window_type, win, blackman_coeff).float().contiguous()
            m = _torch_mel_banks(num_mel_bins, padded, sample_frequency, low_freq, high_freq)
            keep = [w, m]
            o.window = w.data_ptr()
            o.mel_weights = m.data_ptr()
        h = C.c_void_p()
        _lib.check(self.lib.b200fe_plan_create(C.byref(o), C.byref(h)), "b200fe_plan_create")
        del keep
        self.handle = h
        self.num_mel_bins = num_mel_bins
        self.window_size = self.lib.b200fe_window_size(h)
        self.window_shift = self.lib.b200fe_window_shift(h)
        self.tile_frames = self.lib.b200fe_plan_info(h, 5)
        self.uses_ws = bool(self.lib.b200fe_plan_info(h, 6))
        self.has_apply_tiles = self.lib.b200fe_plan_info(h, 7) == 1      # utterance CMVN inside the fused launch
        self.apply_rows = self.lib.b200fe_plan_info(h, 8)                # rows normalised by one CMVN-apply tile
        self.sample_frequency = sample_frequency

    def num_frames(self, n):
        return int(self.lib.b200fe_num_frames(self.handle, int(n)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200fe_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass


class GpuFbankFrontend(torch.nn.Module):
    """fbank (+ peak norm) (+ CMVN) (+ SpecAugment masks) over a padded batch, on one GPU.

    forward(wav, wav_len) -> (feats, feat_len)
      wav      float32 CUDA tensor (B, Nmax), zero padded -- what the unmodified reference dataset
               yields with ``audio_trans: [avgchannel]`` (dataset.py:196-206)
      wav_len  int64 (B,) sample counts, CPU (preferred, no sync) or CUDA
      feats    float32 CUDA (B, Tmax, num_mel_bins), rows >= feat_len[b] are 0
      feat_len int64 CUDA (B,)  -- the ``wav_len`` entry of the reference batch dict
    """

    def __init__(self, num_mel_bins=80, dither=0.0, energy_floor=1.0, frame_length=25.0, frame_shift=10.0,
                 high_freq=0.0, low_freq=20.0, preemphasis_coefficient=0.97, remove_dc_offset=True,
                 round_to_power_of_two=True, sample_frequency=16000.0, snip_edges=True, use_energy=False,
                 use_log_fbank=True, use_power=True, vtln_warp=1.0, window_type="povey", blackman_coeff=0.42,
                 audio_bit=16, peak_norm=False, cmvn="none", cmvn_stats=None, specaug=False,
                 max_freq_width=27, n_freq_mask=2, max_time_width=40, n_time_mask=2, replace_with_zero=False,
                 consume_time_warp_draws=False, time_warp=False, max_time_warp=5, l2_chunk_bytes=None, compact_tiles=True):
        super().__init__()
        if not snip_edges or use_energy or vtln_warp != 1.0 or not round_to_power_of_two:
            raise ValueError("snip_edges=False, use_energy=True, vtln_warp != 1 and round_to_power_of_two=False "
                             "are not reachable from LASR configs and are not implemented")
        if cmvn not in _CMVN_MODES:
            raise ValueError("cmvn must be one of %s" % (_CMVN_MODES,))
        self.opts = dict(num_mel_bins=num_mel_bins, sample_frequency=sample_frequency, frame_length=frame_length,
                         frame_shift=frame_shift, low_freq=low_freq, high_freq=high_freq,
                         preemphasis_coefficient=preemphasis_coefficient, remove_dc_offset=remove_dc_offset,
                         use_power=use_power, use_log_fbank=use_log_fbank, window_type=window_type,
                         blackman_coeff=blackman_coeff, audio_bit=audio_bit, dither=dither)
        self.dither = dither
        self.dither_seed = 0            # advanced on every call; set it (or torch.manual_seed-derived) for reproducibility
        self.num_mel_bins = num_mel_bins
        self.peak_norm = peak_norm
        self.cmvn = cmvn
        self.specaug = specaug
        self.replace_with_zero = replace_with_zero
        # time_warp=True runs the registry transform `specaug` exactly as the reference does (warp, then masks,
        # datatrans.py:136-150); the default is the masks-only scope of BASELINE.json's north_star
        self.time_warp = bool(time_warp) and bool(specaug)
        self.sa = dict(max_freq_width=max_freq_width, n_freq_mask=n_freq_mask, max_time_width=max_time_width,
                       n_time_mask=n_time_mask, consume_time_warp_draws=consume_time_warp_draws or self.time_warp,
                       max_time_warp=max_time_warp)
        self._pre = None
        if self.time_warp:
            # features before SpecAugment come from a child front end with the same options; the warp is
            # out of place, so its output buffer is the final one
            self._pre = GpuFbankFrontend(num_mel_bins=num_mel_bins, dither=dither, frame_length=frame_length, frame_shift=frame_shift,
                                         high_freq=high_freq, low_freq=low_freq, preemphasis_coefficient=preemphasis_coefficient,
                                         remove_dc_offset=remove_dc_offset, sample_frequency=sample_frequency, use_log_fbank=use_log_fbank,
                                         use_power=use_power, window_type=window_type, blackman_coeff=blackman_coeff, audio_bit=audio_bit,
                                         peak_norm=peak_norm, cmvn=cmvn, cmvn_stats=cmvn_stats, specaug=False,
                                         l2_chunk_bytes=l2_chunk_bytes, compact_tiles=compact_tiles)
        self.l2_chunk_bytes = l2_chunk_bytes      # None: one launch per batch (measured fastest); else utterance groups
        self.compact_tiles = compact_tiles
        # extract_host: one copy kernel per group over pinned host memory instead of one DMA per utterance
        self.kernel_h2d = True
        self.pad_tiles = True           # padded rows are zeroed by padding tiles inside the fused launch (False: separate zero-fill kernel)
        # Utterance CMVN applied inside the fused launch by CMVN-apply tiles instead of the post-pass launch; the tiles of an utterance
        # are queued `apply_lag` utterances after its frame tiles.  Opt-in: measured on B200 (C2) at 0.380 ms per step against
        # 0.382 ms for the post pass -- an apply tile costs its CTA about 5 us without FFT work.
        self.inlaunch_cmvn = False
        # time-warp path: masks applied inside the warp launch (False, or B200FE_FUSE_WARP_MASKS=0 for A/B runs: finalize + post pass launches)
        self.fuse_warp_masks = os.environ.get("B200FE_FUSE_WARP_MASKS", "1") != "0"
        self._warp_done = {}             # device -> int32 completion counters of the warp launch (left zero by every call)
        self._warp_stats = {}            # (device, classes, bins) -> float64 statistics workspace of the warp launch (left zero by every call)
        self.apply_lag = 64
        # SpecAugment mean fills (global / no CMVN) by completion tiles inside the fused launch instead of finalize + post pass
        # Opt-in: measured on B200 (C3) at 0.55 ms per step with the completion tiles behind all frame tiles against 0.50 ms for
        # finalize + post pass -- one CTA per utterance cannot retire the scattered partial-sector stores of the masks as fast as a
        # pass of the whole GPU can.
        self.inlaunch_fills = os.environ.get("B200FE_INLAUNCH_FILLS", "0") != "0"
        self.fill_lag = int(os.environ.get("B200FE_FILL_LAG", "600"))
        self.overlap_calls = True       # extract_host: the H2D copies of the next call may start before this call's D2H tail ends
        self.kernel_d2h = True
        self._plans = {}
        self._host_cache = {}
        self.launch_count = 0           # kernels launched by this object (bench.py reports it)
        self._apply_flags = None
        self.profile_events = None      # set to [] to collect (start, stop) CUDA events around every fused launch
        self.register_buffer("cmvn_mean", None, persistent=False)
        self.register_buffer("cmvn_istd", None, persistent=False)
        if cmvn == "global":
            if cmvn_stats is None:
                raise ValueError("cmvn='global' needs cmvn_stats ([2, D+1] Kaldi statistics or a (mean, istd) pair)")
            self.set_global_cmvn(cmvn_stats)

    # -- plumbing ---------------------------------------------------------------------------
    def plan(self, device):
        key = torch.device(device).index or 0
        if key not in self._plans:
            with torch.cuda.device(key):
                self._plans[key] = FbankPlan(**self.opts)
        return self._plans[key]

    def set_global_cmvn(self, stats, norm_vars=True):
        if isinstance(stats, (tuple, list)) and len(stats) == 2:
            mean, istd = (torch.as_tensor(np.asarray(s), dtype=torch.float32) for s in stats)
        else:
            st = np.ascontiguousarray(np.asarray(stats, dtype=np.float64))
            d = self.num_mel_bins
            if st.shape != (2, d + 1):
                raise ValueError("cmvn statistics must have shape (2, %d)" % (d + 1))
            mean_np = np.zeros(d, dtype=np.float32)
            istd_np = np.zeros(d, dtype=np.float32)
            lib = _lib.load()
            _lib.check(lib.b200fe_cmvn_from_stats(st.ctypes.data_as(C.POINTER(C.c_double)), d, int(norm_vars),
                                                  mean_np.ctypes.data_as(_lib.c_fp), istd_np.ctypes.data_as(_lib.c_fp)),
                       "b200fe_cmvn_from_stats")
            mean, istd = torch.from_numpy(mean_np), torch.from_numpy(istd_np)
        self.cmvn_mean = mean.contiguous()
        self.cmvn_istd = istd.contiguous()

    def frame_counts(self, wav_len):
        p = self.opts
        win = int(p["sample_frequency"] * p["frame_length"] * 0.001)
        shift = int(p["sample_frequency"] * p["frame_shift"] * 0.001)
        n = np.asarray(wav_len, dtype=np.int64)
        return np.where(n >= win, 1 + (n - win) // shift, 0), win

    # -- the hot path ------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, wav, wav_len, max_frames=None, masks=None, out=None, out_len=None, wav_offsets=None, dither_noise=None,
                uniform_frames=False, packed_out=False, out_dtype=None):
        """``wav_offsets`` (int64 host array, multiples of 4) switches to the packed layout: ``wav`` is then
        a 1-D CUDA tensor and utterance b occupies ``wav[wav_offsets[b] : wav_offsets[b] + wav_len[b]]``.
        ``out_dtype=torch.bfloat16`` returns the features in the precision the encoder's first convolution runs
        in under autocast (subsampling.py:53-57): computed in float32 as always, rounded once (nearest even) by one extra pass."""
        if not wav.is_cuda:
            raise RuntimeError("GpuFbankFrontend has no CPU path: wav must be a CUDA tensor")
        if out_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
        if out_dtype == torch.bfloat16:
            if out is not None and (out.dtype != torch.bfloat16 or not out.is_contiguous()):
                raise ValueError("out must be a contiguous bfloat16 tensor when out_dtype is bfloat16")
            f32, flen = self.forward(wav, wav_len, max_frames=max_frames, masks=masks, out_len=out_len, wav_offsets=wav_offsets,
                                     dither_noise=dither_noise, uniform_frames=uniform_frames, packed_out=packed_out)
            dst = out if out is not None else torch.empty(f32.shape, dtype=torch.bfloat16, device=f32.device)
            if tuple(dst.shape) != tuple(f32.shape):
                raise ValueError("out must have shape %s" % (tuple(f32.shape),))
            _lib.check(_lib.load().b200fe_cast_bf16(_ptr(f32), _ptr(dst), f32.numel(), C.c_void_p(torch.cuda.current_stream(f32.device).cuda_stream)),
                       "b200fe_cast_bf16")
            self.launch_count += 1
            return dst, flen
        if self.time_warp:
            if packed_out:
                raise ValueError("packed_out is not available together with time_warp")
            return self._forward_time_warp(wav, wav_len, max_frames, masks, out, out_len, wav_offsets, dither_noise)
        packed = wav_offsets is not None
        if wav.dtype not in (torch.float32, torch.int16) or wav.dim() != (1 if packed else 2):
            raise ValueError("wav must be float32 or int16 PCM, (B, Nmax) or 1-D with wav_offsets")
        i16 = wav.dtype == torch.int16
        esz = 2 if i16 else 4
        if not packed and wav.stride(1) != 1:
            wav = wav.contiguous()
        dev = wav.device
        B = len(wav_len) if packed else wav.shape[0]
        if packed:
            off_host = np.ascontiguousarray(wav_offsets, dtype=np.int64)
            if (off_host % (16 // esz) != 0).any():
                raise ValueError("wav_offsets must be multiples of 16 bytes")
            row_stride = int(wav.numel())
            row_elems = int(wav.numel())
        else:
            off_host = None
            row_stride = wav.stride(0)
            row_elems = wav.shape[1]
        plan = self.plan(dev)
        lib = plan.lib
        if torch.is_tensor(wav_len) and wav_len.is_cuda:
            len_host = wav_len.cpu().numpy() if (self.specaug or max_frames is None) else None
            len_dev = wav_len.to(torch.int64).contiguous()
        else:
            len_host = np.asarray(wav_len, dtype=np.int64).reshape(-1)
            len_dev = None                       # uploaded below together with the other small tables
        if len_host is not None:
            T_host, win = self.frame_counts(len_host)
            if (len_host < win).any():
                # torchaudio asserts (TA:142); LASR filters min_duration upstream (dataset.py:243,272)
                raise AssertionError("choose a window size {} that is [2, {}]".format(win, int(len_host.min())))
            if (len_host > row_elems).any():
                raise ValueError("wav_len exceeds the padded width")
            Tmax = int(T_host.max()) if max_frames is None else int(max_frames)
        else:
            T_host, Tmax = None, int(max_frames)
        D = self.num_mel_bins
        ooff_dev = ooff_host = None
        if packed_out:
            # packed features: utterance b occupies rows [offsets[b], offsets[b] + T_b) of a (sum T, D) tensor
            if len_host is None:
                raise ValueError("packed_out needs host lengths (the row offsets are their prefix sums)")
            ooff_host = np.zeros(B, dtype=np.int64)
            np.cumsum(T_host[:-1], out=ooff_host[1:])
            oshape = (int(T_host.sum()), D)
        else:
            oshape = (B, Tmax, D)
        if out is not None:
            if tuple(out.shape) != oshape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
                raise ValueError("out must be a contiguous float32 %s tensor on the input's device" % (oshape,))
            feats = out
        else:
            feats = torch.empty(oshape, dtype=torch.float32, device=dev)
        feat_len = out_len if out_len is not None else torch.empty((B,), dtype=torch.int64, device=dev)
        cur_stream = torch.cuda.current_stream(dev)
        stream = C.c_void_p(cur_stream.cuda_stream)

        n_f = n_t = 0
        if self.specaug:
            warp_np = None
            if masks is None:
                if self.time_warp:
                    m_np, b_np, warp_np = _specaug.plan_batch(T_host, D, return_warp=True, **self.sa)
                else:
                    m_np, b_np = _specaug.plan_batch(T_host, D, **self.sa)
            else:
                m_np = np.ascontiguousarray(masks, dtype=np.int32)
                n_f_, n_t_ = self.sa["n_freq_mask"], self.sa["n_time_mask"]
                b_np = np.sort(m_np[:, n_f_:].reshape(B, -1), axis=1).astype(np.int32)
            n_f, n_t = self.sa["n_freq_mask"], self.sa["n_time_mask"]
        else:
            m_np = b_np = None
        mean_fill = self.specaug and not self.replace_with_zero
        utt_cmvn = self.cmvn in ("utt_mean", "utt_meanvar")
        need_post = mean_fill or utt_cmvn
        # replace_with_zero together with utterance CMVN: the statistics must see the unmasked features and the zeros go in
        # after normalisation (CMVN first, masks second -- the order of the mean-fill and time-warp paths), so the post
        # pass writes them (fill_zero) instead of the fused launch
        zero_in_post = self.specaug and self.replace_with_zero and utt_cmvn
        n_cls = (2 * n_t + 1) if mean_fill else 1

        gm = gi = None
        if self.cmvn == "global":
            gm = self.cmvn_mean.to(dev)
            gi = self.cmvn_istd.to(dev)
            self.cmvn_mean, self.cmvn_istd = gm, gi

        # Utterance groups sized so that a group's features stay L2-resident between the fused
        # launch and the in-place post pass.
        if need_post and self.l2_chunk_bytes:
            per_utt = Tmax * D * 4
            group = max(1, min(B, self.l2_chunk_bytes // max(per_utt, 1)))
        else:
            group = B
        # ONE host -> device transfer per call for the small tables; the work list of the fused launch is built on the
        # device from the sample counts (b200fe_build_tile_table_device), the experimental kernel keeps the host-built list
        tab0 = None
        pads = self.pad_tiles and not packed_out and not plan.uses_ws
        dev_tables = self.compact_tiles and not plan.uses_ws
        # a batch whose utterances all fill the padded width (BASELINE config 1: 16 x 10 s) has no ragged tail to skip and no padding
        # rows to zero: the fused launch walks the (batch x max_frames) grid directly, without the list-builder launch in front of it
        # -- only where the launch itself is what the step costs: the grid walk is a STATIC round-robin, and on a large batch the
        # dynamic work list balances better than the 5 us list builder costs (BASELINE config 3, 512 x 10 s, global CMVN:
        # 0.366 ms per step on the grid walk against 0.347 ms through the work list)
        full_grid = (len_host is not None and not need_post and not packed_out and not uniform_frames and bool((T_host == Tmax).all())
                     and B * ((Tmax + plan.tile_frames - 1) // plan.tile_frames) <= 1024)
        if full_grid:
            dev_tables = False
        # utterance CMVN by apply tiles of the same launch: default option set, float32 input, padded layout, no SpecAugment
        apply_tiles = (utt_cmvn and self.inlaunch_cmvn and dev_tables and plan.has_apply_tiles and not mean_fill and not self.specaug
                       and not packed_out and not i16 and not self.peak_norm and not uniform_frames and self.dither == 0.0
                       and feats.data_ptr() % 16 == 0)
        # SpecAugment mean fills by completion tiles of the same launch (global / no CMVN): no finalize launch, no mask pass
        fill_tiles = (mean_fill and not utt_cmvn and self.inlaunch_fills and dev_tables and plan.has_apply_tiles and not packed_out and not i16
                      and not self.peak_norm and not uniform_frames and feats.data_ptr() % 16 == 0)
        if self.compact_tiles and plan.uses_ws and len_host is not None and group >= B:
            tab0 = _tile_table(T_host, plan.tile_frames, None)
        up = _h2d_many([off_host, len_host if len_dev is None else None, m_np, b_np, ooff_host, tab0], dev)
        off_dev, masks_dev, bounds_dev, ooff_dev, tab0_dev = up[0], up[2], up[3], up[4], up[5]
        if len_dev is None:
            len_dev = up[1]

        peak = None
        if self.peak_norm:
            peak = torch.empty((B,), dtype=torch.float32, device=dev)
            absmax = lib.b200fe_peak_absmax_i16 if i16 else lib.b200fe_peak_absmax
            _lib.check(absmax(plan.handle, _ptr(wav), row_stride if not packed else int(len_host.max()) if len_host is not None else row_stride,
                              _ptr(off_dev), _ptr(len_dev), B, _ptr(peak), stream), "b200fe_peak_absmax")
            self.launch_count += 2          # memset + abs-max kernel

        stats = cm = ci = fills = None
        if need_post:
            # with a device-built work list the builder kernel clears the accumulators of its utterance group (no memset node)
            stats = (torch.empty if dev_tables else torch.zeros)((B, n_cls + 1, D), dtype=torch.float64, device=dev)
            if utt_cmvn:
                cm = torch.empty((B, D), dtype=torch.float32, device=dev)
                ci = torch.empty((B, D), dtype=torch.float32, device=dev)
            if mean_fill or zero_in_post:
                fills = torch.empty((B, n_f + n_t), dtype=torch.float32, device=dev)

        def off(t, b0, per):
            return C.c_void_p(t.data_ptr() + b0 * per) if t is not None else C.c_void_p(0)

        for b0 in range(0, B, group):
            nb = min(group, B - b0)
            a = _lib.FbankArgs()
            a.wav_dtype = 1 if i16 else 0
            a.uniform_frames = 1 if uniform_frames else 0      # lock-step streaming: every utterance yields exactly max_frames frames
            if packed:
                a.d_wav = _ptr(wav)
                a.wav_stride = row_stride
                a.d_wav_offsets = off(off_dev, b0, 8)
                a.offsets_aligned = 1
            else:
                a.d_wav = off(wav, b0, wav.stride(0) * esz)
                a.wav_stride = wav.stride(0)
            a.d_nsamp = off(len_dev, b0, 8)
            a.batch = nb
            a.d_peak = off(peak, b0, 4)
            if packed_out:
                a.d_out = _ptr(feats)
                a.d_out_offsets = off(ooff_dev, b0, 8)
            else:
                a.d_out = off(feats, b0, Tmax * D * 4)
            a.d_out_len = off(feat_len, b0, 8)
            a.max_frames = Tmax
            if self.dither != 0.0:
                if dither_noise is not None:
                    # (B, Tmax, window) standard-normal noise, e.g. torch.randn under the reference's seed
                    if dither_noise.shape != (B, Tmax, plan.window_size) or dither_noise.dtype != torch.float32 or not dither_noise.is_contiguous():
                        raise ValueError("dither_noise must be contiguous float32 (B, Tmax, window_size)")
                    a.d_dither_noise = off(dither_noise, b0, Tmax * plan.window_size * 4)
                self.dither_seed += 1
                a.dither_seed = self.dither_seed
            if gm is not None:
                a.d_cmvn_mean, a.d_cmvn_istd, a.cmvn_stride = _ptr(gm), _ptr(gi), 0
            if self.specaug and not zero_in_post:
                a.d_masks = off(masks_dev, b0, (n_f + n_t) * 2 * 4)
                a.n_freq_masks, a.n_time_masks = n_f, n_t
                a.mask_zero = int(self.replace_with_zero)
            if dev_tables:
                # ragged batch: only tiles with valid frames (+ padding tiles), consumed through an atomic counter; the list,
                # its length and the counter reset come from one small kernel on the device-resident sample counts
                cap = lib.b200fe_tile_table_capacity(plan.handle, nb, Tmax, 1 if pads else 0)
                if apply_tiles:
                    cap += nb * ((Tmax + plan.apply_rows - 1) // plan.apply_rows)
                if fill_tiles:
                    cap += nb                                    # one completion tile per utterance
                any_apply = apply_tiles or fill_tiles
                work = torch.empty((2 * cap + 2 + (nb + 1 if any_apply else 0),), dtype=torch.int32, device=dev)   # table | n_tiles | counter | done[nb] | error
                done_ptr = C.c_void_p(work.data_ptr() + 8 * cap + 8) if any_apply else C.c_void_p(0)
                zero_ptr, zero_bytes = (off(stats, b0, (n_cls + 1) * D * 8), nb * (n_cls + 1) * D * 8) if need_post else (C.c_void_p(0), 0)
                _lib.check(lib.b200fe_build_work_list_device(plan.handle, a.d_nsamp, nb, Tmax, 1 if pads else 0,
                                                             (max(1, int(self.apply_lag)) if apply_tiles else -max(1, int(self.fill_lag)) if fill_tiles else 0),
                                                             _ptr(work), cap, C.c_void_p(work.data_ptr() + 8 * cap),
                                                             C.c_void_p(work.data_ptr() + 8 * cap + 4), done_ptr, zero_ptr, zero_bytes, stream),
                           "b200fe_build_work_list_device")
                if apply_tiles:
                    a.apply_cmvn_mode = 1 if self.cmvn == "utt_mean" else 2
                    a.d_utt_done = done_ptr
                    a.d_utt_mean, a.d_utt_istd = off(cm, b0, D * 4), off(ci, b0, D * 4)
                    self._apply_flags = (work, 2 * cap + 2 + nb)          # tests read the error flag
                if fill_tiles:
                    a.apply_cmvn_mode = 3
                    a.d_utt_done = done_ptr
                    a.d_fills = off(fills, b0, (n_f + n_t) * 4)
                    self._apply_flags = (work, 2 * cap + 2 + nb)
                a.d_tile_table, a.n_tiles = _ptr(work), cap
                a.d_n_tiles, a.d_work_counter = C.c_void_p(work.data_ptr() + 8 * cap), C.c_void_p(work.data_ptr() + 8 * cap + 4)
                a.tile_table_pads = 1 if pads else 0
                self.launch_count += 1 if pads else 2      # table kernel (+ zero-pad kernel)
            elif self.compact_tiles and len_host is not None and not full_grid:
                if tab0_dev is not None:
                    tab_dev, tot = tab0_dev, tab0.shape[0]
                else:
                    tab = _tile_table(T_host[b0:b0 + nb], plan.tile_frames, None)
                    tot = tab.shape[0]
                    tab_dev = _h2d(tab, dev)
                counter = torch.empty((1,), dtype=torch.int32, device=dev)
                a.d_tile_table, a.n_tiles, a.d_work_counter = _ptr(tab_dev), tot, _ptr(counter)
                self.launch_count += 2                     # counter memset + zero-pad kernel
            if need_post:
                a.d_stats = off(stats, b0, (n_cls + 1) * D * 8)
                a.stats_stride = (n_cls + 1) * D
                a.n_row_classes = n_cls
                if mean_fill:
                    a.d_row_bounds = off(bounds_dev, b0, 2 * n_t * 4)
            if self.profile_events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(cur_stream)
            _lib.check(lib.b200fe_fbank_fused(plan.handle, C.byref(a), stream), "b200fe_fbank_fused")
            self.launch_count += 1
            if self.profile_events is not None:
                e1.record(cur_stream)
                self.profile_events.append((e0, e1))
            if need_post and not apply_tiles and not fill_tiles:
                q = _lib.PostArgs()
                q.d_feats = a.d_out
                q.d_feat_offsets = a.d_out_offsets
                q.d_nsamp = a.d_nsamp
                q.batch = nb
                q.max_frames = Tmax
                q.d_stats = a.d_stats
                q.stats_stride = a.stats_stride
                q.d_row_bounds = a.d_row_bounds
                q.n_row_classes = n_cls
                q.cmvn_mode = {"utt_mean": 1, "utt_meanvar": 2}.get(self.cmvn, 0)
                q.d_cmvn_mean = off(cm, b0, D * 4)
                q.d_cmvn_istd = off(ci, b0, D * 4)
                if mean_fill or zero_in_post:
                    q.d_masks = off(masks_dev, b0, (n_f + n_t) * 2 * 4)
                    q.n_freq_masks, q.n_time_masks = n_f, n_t
                    q.d_fills = off(fills, b0, (n_f + n_t) * 4)
                    q.fill_zero = int(zero_in_post)
                _lib.check(lib.b200fe_postpass(plan.handle, C.byref(q), stream), "b200fe_postpass")
                self.launch_count += 2 if (mean_fill or zero_in_post) else 1      # (finalize +) in-place post pass
        self.last = dict(stats=stats, fills=fills, masks=masks_dev, utt_mean=cm, utt_istd=ci, peak=peak, feat_offsets=ooff_host,
                         apply_flags=self._apply_flags if (apply_tiles or fill_tiles) else None)
        return feats, feat_len

    def _forward_time_warp(self, wav, wav_len, max_frames, masks, out, out_len, wav_offsets, dither_noise):
        """norm -> fbank -> CMVN (child front end) -> time warp -> frequency / time masks: the registry
        transform `specaug` exactly as the reference applies it (datatrans.py:136-150)."""
        if masks is not None:
            raise ValueError("explicit masks cannot be combined with time_warp (the warp draws precede them)")
        dev = wav.device
        len_host = np.asarray(wav_len.cpu() if torch.is_tensor(wav_len) else wav_len, dtype=np.int64).reshape(-1)
        B, D = len(len_host), self.num_mel_bins
        pre, flen = self._pre(wav, len_host, max_frames=max_frames, out_len=out_len, wav_offsets=wav_offsets, dither_noise=dither_noise)
        self.launch_count += self._pre.launch_count
        self._pre.launch_count = 0
        Tmax = pre.shape[1]
        # (drawing and uploading the small tables BEFORE the fbank launches, so that the warp launch follows them directly, measured
        # 7-10 us slower per C3 step: the list builder then no longer overlaps the previous step's tail)
        T_host, _ = self.frame_counts(len_host)
        m_np, b_np, w_np = _specaug.plan_batch(T_host, D, return_warp=True, **self.sa)
        n_f, n_t = self.sa["n_freq_mask"], self.sa["n_time_mask"]
        n_cls = 2 * n_t + 1
        masks_dev, bounds_dev, warp_dev, len_dev = _h2d_many([m_np, b_np, w_np, len_host], dev)     # one upload (each costs ~80 us of host time)
        fills = torch.empty((B, n_f + n_t), dtype=torch.float32, device=dev)
        if self.fuse_warp_masks:
            key = (dev, n_cls, D)
            stats = self._warp_stats.get(key)                 # workspace: the launch leaves it zero
            if stats is None or stats.shape[0] < B:
                stats = self._warp_stats[key] = torch.zeros((max(B, 256), n_cls + 1, D), dtype=torch.float64, device=dev)
        else:
            stats = torch.zeros((B, n_cls + 1, D), dtype=torch.float64, device=dev)
        if out is not None:
            if out.shape != pre.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.data_ptr() == pre.data_ptr():
                raise ValueError("out must be a distinct contiguous float32 (B, Tmax, D) tensor")
            feats = out
        else:
            feats = torch.empty_like(pre)
        plan = self.plan(dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        w = _lib.WarpArgs()
        w.d_in, w.d_out, w.d_nsamp, w.batch, w.max_frames = _ptr(pre), _ptr(feats), _ptr(len_dev), B, Tmax
        w.d_warp, w.d_stats, w.stats_stride, w.d_row_bounds, w.n_row_classes = _ptr(warp_dev), _ptr(stats), (n_cls + 1) * D, _ptr(bounds_dev), n_cls
        if self.fuse_warp_masks:
            # finalize + mask fill inside the warp launch (the CTA that completes an utterance applies its masks)
            done = self._warp_done.get(dev)
            if done is None or done.numel() < B:
                done = self._warp_done[dev] = torch.zeros((max(B, 1024),), dtype=torch.int32, device=dev)
            w.d_masks, w.n_freq_masks, w.n_time_masks, w.d_fills = _ptr(masks_dev), n_f, n_t, _ptr(fills)
            w.fill_zero, w.d_utt_done = int(self.replace_with_zero), _ptr(done)
            _lib.check(plan.lib.b200fe_time_warp(plan.handle, C.byref(w), stream), "b200fe_time_warp")
            self.launch_count += 1
            self.last = dict(stats=None, fills=fills, masks=masks_dev, warp=warp_dev, pre=pre)
            return feats, flen
        _lib.check(plan.lib.b200fe_time_warp(plan.handle, C.byref(w), stream), "b200fe_time_warp")
        q = _lib.PostArgs()
        q.d_feats, q.d_nsamp, q.batch, q.max_frames = _ptr(feats), _ptr(len_dev), B, Tmax
        q.d_stats, q.stats_stride, q.d_row_bounds, q.n_row_classes = _ptr(stats), (n_cls + 1) * D, _ptr(bounds_dev), n_cls
        q.cmvn_mode = 0
        q.d_masks, q.n_freq_masks, q.n_time_masks, q.d_fills = _ptr(masks_dev), n_f, n_t, _ptr(fills)
        q.fill_zero = int(self.replace_with_zero)
        _lib.check(plan.lib.b200fe_postpass(plan.handle, C.byref(q), stream), "b200fe_postpass")
        self.launch_count += 3          # warp + finalize + mask fill
        self.last = dict(stats=stats, fills=fills, masks=masks_dev, warp=warp_dev, pre=pre)
        return feats, flen

    # -- host-to-host path: the drop-in for the reference's collate_fn (features back on the host) --
    @staticmethod
    def pack_host(wavs, dtype=torch.float32, pin=True, out=None):
        """Packs a list of 1-D utterances (what the reference's collate loop receives, dataset.py:190-206) into ONE
        host buffer with 16-byte aligned starts -- the layout extract_host moves with a single DMA per group.
        ``out``: a (pinned) 1-D tensor to reuse when it is large enough (pinning costs milliseconds; a loader keeps one).
        Returns (packed CPU tensor, lengths int64, offsets int64)."""
        esz = torch.empty((), dtype=dtype).element_size()
        al = 16 // esz
        lens = np.array([len(w) for w in wavs], dtype=np.int64)
        offs = np.zeros(len(wavs), dtype=np.int64)
        np.cumsum((lens[:-1] + al - 1) // al * al, out=offs[1:])
        total = int(offs[-1] + (lens[-1] + al - 1) // al * al) if len(wavs) else 0
        if out is not None and out.dtype == dtype and out.dim() == 1 and out.numel() >= total:
            buf = out[:total]
        else:
            buf = torch.empty((total,), dtype=dtype)
            if pin:
                buf = buf.pin_memory()
        view = buf.numpy()
        for w, o, k, nxt in zip(wavs, offs, lens, list(offs[1:]) + [total]):
            view[o:o + k] = w
            view[o + k:nxt] = 0                       # alignment gap
        return buf, lens, offs

    @torch.no_grad()
    def extract_host(self, wav_host, wav_len, device="cuda:0", group_bytes=32 << 20, return_host=True, wav_offsets=None,
                     packed_out=False):
        """Pipelined H2D copy -> fused kernels -> D2H copy over utterance groups on three streams.

        wav_host: zero-padded float32 CPU tensor (B, Nmax), ideally pinned (what batch_list builds).
        Only the VALID samples of each utterance cross PCIe: they are staged into a packed device
        buffer (b200fe_h2d_ragged) that the fused kernel reads through per-utterance offsets.
        Returns (feats, feat_len): pinned CPU tensors when ``return_host`` (what
        AudioDataSet.collate_fn hands to the trainer, dataset.py:222-232), else CUDA tensors (the
        case where the encoder consumes them in place).

        ``wav_offsets`` (from ``pack_host``): ``wav_host`` is the 1-D packed pinned buffer, moved with one DMA per group.
        ``packed_out``: features come back packed, ``(sum T, D)`` with no padding rows, as ``(feats, feat_len,
        row_offsets)`` -- one DMA per group in both directions."""
        dev = torch.device(device)
        packed_in = wav_offsets is not None
        if wav_host.dtype not in (torch.float32, torch.int16) or wav_host.stride(-1) != 1 or wav_host.dim() != (1 if packed_in else 2):
            raise ValueError("wav_host must be a float32 or int16 tensor, (B, Nmax) with contiguous rows or 1-D with wav_offsets")
        esz = wav_host.element_size()
        al = 16 // esz                                    # utterance starts are 16-byte aligned in the packed buffer
        len_host = np.ascontiguousarray(np.asarray(wav_len, dtype=np.int64).reshape(-1))
        B = len(len_host)
        Nmax = int(wav_host.shape[-1])
        T_host, win = self.frame_counts(len_host)
        if (len_host < win).any():
            raise AssertionError("choose a window size {} that is [2, {}]".format(win, int(len_host.min())))
        Tmax, D = int(T_host.max()), self.num_mel_bins
        if packed_in:
            # packed host input (pack_host): the device buffer mirrors it, one DMA per utterance group
            offs = np.ascontiguousarray(np.asarray(wav_offsets, dtype=np.int64).reshape(-1))
            if (offs % al != 0).any() or (np.diff(offs) < len_host[:-1]).any() or offs[-1] + len_host[-1] > Nmax:
                raise ValueError("wav_offsets must be ascending, 16-byte aligned and leave room for every utterance")
            total = Nmax
        else:
            offs = np.zeros(B, dtype=np.int64)
            np.cumsum((len_host[:-1] + al - 1) // al * al, out=offs[1:])
            total = int(offs[-1] + (len_host[-1] + al - 1) // al * al)
        foff = np.zeros(B + 1, dtype=np.int64)                     # first feature row of every utterance in the packed layout
        np.cumsum(T_host, out=foff[1:])
        rows_total = int(foff[-1])
        key = (B, Nmax, Tmax, total, dev.index or 0, wav_host.dtype, packed_in, packed_out, rows_total if packed_out else 0)
        c = self._host_cache.get(key)
        if c is None:
            # new geometry: the old staging buffers may still be read by queued kernels / copies of the previous call, and the
            # allocator can hand their memory straight back -- drain the device before dropping them (HostPipeline, the
            # path behind B200Collate, keeps grow-only capacity buffers instead and never gets here)
            if self._host_cache:
                torch.cuda.synchronize(dev)
            self._host_cache.clear()
            # two staging buffers: the H2D copies of call k+1 overlap the compute / D2H tail of call k
            c = dict(wavs=[torch.zeros((total + 64,), dtype=wav_host.dtype, device=dev) for _ in range(2)],
                     wav_free=[None, None], turn=0,
                     feats=torch.empty((rows_total, D) if packed_out else (B, Tmax, D), dtype=torch.float32, device=dev),
                     flen=torch.empty((B,), dtype=torch.int64, device=dev),
                     hfeats=torch.zeros((rows_total, D) if packed_out else (B, Tmax, D), dtype=torch.float32).pin_memory(),
                     hlen=torch.empty((B,), dtype=torch.int64, pin_memory=True),
                     hrows=np.zeros(B, dtype=np.int64),      # rows of the host buffer that are not known to be zero
                     s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev))
            self._host_cache[key] = c
        lib = _lib.load()
        # rows to bring back per utterance: valid now, or valid in the previous batch held by the host buffer
        d2h_rows = np.ascontiguousarray(np.maximum(T_host, c["hrows"]).astype(np.int64))
        main = torch.cuda.current_stream(dev)
        # Pinned host buffers are read / written by ONE copy kernel per group (b200fe_copy_ragged) instead of one DMA
        # set-up per utterance; byte offsets and sizes of every row, device resident:
        #   0: host row start   1: packed device start   2: valid bytes   3: feature block start   4: feature bytes to bring back
        kh2d = self.kernel_h2d and wav_host.is_pinned() and not packed_in
        kd2h = self.kernel_d2h and return_host and not packed_out
        if kh2d or kd2h:
            tab = np.stack([np.arange(B, dtype=np.int64) * ((0 if packed_in else wav_host.stride(0)) * esz), offs * esz, len_host * esz,
                            np.arange(B, dtype=np.int64) * (Tmax * D * 4), d2h_rows * (D * 4)])
            tab_dev = torch.from_numpy(np.ascontiguousarray(tab)).to(dev, non_blocking=True)
            tptr = tab_dev.data_ptr()
            c["tabs"] = (c.get("tabs", ()) + (tab_dev,))[-2:]        # alive until the copy kernels of this call have run
            ev_tab = torch.cuda.Event()
            ev_tab.record(main)
        s_in, s_out = c["s_in"], c["s_out"]
        slot = c["turn"] & 1
        c["turn"] += 1
        dwav = c["wavs"][slot]
        if not self.overlap_calls:
            s_in.wait_stream(main)
        elif c["wav_free"][slot] is not None:
            s_in.wait_event(c["wav_free"][slot])          # the call that last read this staging buffer has finished computing
        else:
            s_in.wait_stream(main)                        # fresh buffers: their zero fill was queued on the main stream
        s_out.wait_stream(main)
        if kh2d:
            s_in.wait_event(ev_tab)
        # utterance groups of ~group_bytes of valid audio (shrinking the last groups to shorten the non-overlapped tail was
        # measured slower: with overlapped calls the tail already hides behind the next call's copies)
        bounds = [0]
        acc = 0
        for b in range(B):
            acc += int(len_host[b]) * esz
            if acc >= group_bytes:
                bounds.append(b + 1)
                acc = 0
        if bounds[-1] != B:
            bounds.append(B)
        self.h2d_bytes = self.d2h_bytes = 0
        for b0, b1 in zip(bounds[:-1], bounds[1:]):
            if packed_in:
                o0 = int(offs[b0])
                o1 = int(offs[b1]) if b1 < B else int(offs[-1] + len_host[-1])
                with torch.cuda.stream(s_in):
                    dwav[o0:o1].copy_(wav_host[o0:o1], non_blocking=True)
            elif kh2d:
                _lib.check(lib.b200fe_copy_ragged(C.c_void_p(wav_host.data_ptr()), C.c_void_p(tptr + (0 * B + b0) * 8), _ptr(dwav),
                                                  C.c_void_p(tptr + (1 * B + b0) * 8), C.c_void_p(tptr + (2 * B + b0) * 8), b1 - b0,
                                                  int(len_host[b0:b1].max()) * esz, C.c_void_p(s_in.cuda_stream)), "b200fe_copy_ragged")
                self.launch_count += 1
            else:
                _lib.check(lib.b200fe_h2d_ragged(C.c_void_p(wav_host.data_ptr() + b0 * wav_host.stride(0) * esz), wav_host.stride(0),
                                                 C.c_void_p(len_host.ctypes.data + b0 * 8), C.c_void_p(offs.ctypes.data + b0 * 8), b1 - b0,
                                                 _ptr(dwav), esz, C.c_void_p(s_in.cuda_stream)), "b200fe_h2d_ragged")
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
            self.h2d_bytes += ((o1 - o0) * esz if packed_in else int(len_host[b0:b1].sum()) * esz) + (b1 - b0) * 16
            main.wait_event(ev_in)
            r0, r1 = int(foff[b0]), int(foff[b1])
            self.forward(dwav, len_host[b0:b1], max_frames=Tmax, out=c["feats"][r0:r1] if packed_out else c["feats"][b0:b1],
                         out_len=c["flen"][b0:b1], wav_offsets=offs[b0:b1], packed_out=packed_out)
            if return_host:
                ev_c = torch.cuda.Event()
                ev_c.record(main)
                s_out.wait_event(ev_c)
                if packed_out:
                    with torch.cuda.stream(s_out):
                        c["hfeats"][r0:r1].copy_(c["feats"][r0:r1], non_blocking=True)
                    self.d2h_bytes += (r1 - r0) * D * 4
                    continue
                if kd2h:
                    _lib.check(lib.b200fe_copy_ragged(_ptr(c["feats"]), C.c_void_p(tptr + (3 * B + b0) * 8), C.c_void_p(c["hfeats"].data_ptr()),
                                                      C.c_void_p(tptr + (3 * B + b0) * 8), C.c_void_p(tptr + (4 * B + b0) * 8), b1 - b0,
                                                      int(d2h_rows[b0:b1].max()) * D * 4, C.c_void_p(s_out.cuda_stream)), "b200fe_copy_ragged")
                    self.launch_count += 1
                else:
                    _lib.check(lib.b200fe_d2h_ragged(C.c_void_p(c["feats"].data_ptr() + b0 * Tmax * D * 4), D, Tmax,
                                                     C.c_void_p(d2h_rows.ctypes.data + b0 * 8), b1 - b0,
                                                     C.c_void_p(c["hfeats"].data_ptr() + b0 * Tmax * D * 4), C.c_void_p(s_out.cuda_stream)),
                               "b200fe_d2h_ragged")
                self.d2h_bytes += int(d2h_rows[b0:b1].sum()) * D * 4
        ev_free = torch.cuda.Event()
        ev_free.record(main)
        c["wav_free"][slot] = ev_free
        if return_host:
            with torch.cuda.stream(s_out):
                c["hlen"].copy_(c["flen"], non_blocking=True)
            self.d2h_bytes += B * 8
            main.wait_stream(s_out)
            c["hrows"] = T_host.astype(np.int64).copy()
            return (c["hfeats"], c["hlen"], foff[:-1].copy()) if packed_out else (c["hfeats"], c["hlen"])
        return (c["feats"], c["flen"], foff[:-1].copy()) if packed_out else (c["feats"], c["flen"])

    # -- global CMVN statistics (Kaldi compute-cmvn-stats), one fused pass without feature output --
    @torch.no_grad()
    def accumulate_stats(self, wav, wav_len, stats=None):
        """Adds this batch's [sum | count ; sumsq | 0] to ``stats`` (float64 CUDA [2, D+1])."""
        if not wav.is_cuda or wav.dim() != 2 or wav.dtype not in (torch.float32, torch.int16):
            raise ValueError("accumulate_stats: wav must be a float32 or int16 CUDA tensor (B, Nmax)")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        dev = wav.device
        plan = self.plan(dev)
        B, D = wav.shape[0], self.num_mel_bins
        len_host = np.asarray(wav_len.cpu() if torch.is_tensor(wav_len) else wav_len, dtype=np.int64).reshape(-1)
        T_host, win = self.frame_counts(len_host)
        if len(len_host) != B:
            raise ValueError("accumulate_stats: one length per utterance")
        if (len_host < win).any():
            raise AssertionError("choose a window size {} that is [2, {}]".format(win, int(len_host.min())))
        if (len_host > wav.shape[1]).any():
            raise ValueError("wav_len exceeds the padded width")
        len_dev = _h2d(len_host, dev)
        if stats is None:
            stats = torch.zeros((2, D + 1), dtype=torch.float64, device=dev)
        peak = None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if self.peak_norm:
            peak = torch.empty((B,), dtype=torch.float32, device=dev)
            absmax = plan.lib.b200fe_peak_absmax_i16 if wav.dtype == torch.int16 else plan.lib.b200fe_peak_absmax
            _lib.check(absmax(plan.handle, _ptr(wav), wav.stride(0), C.c_void_p(0), _ptr(len_dev), B, _ptr(peak), stream), "b200fe_peak_absmax")
        # per-utterance accumulators (no atomic contention on one 2 x D block), compact tile list, then one
        # fp64 reduction over the batch
        dev_list = self.compact_tiles and not plan.uses_ws        # the list builder clears the accumulators as well
        acc = (torch.empty if dev_list else torch.zeros)((B, 2, D), dtype=torch.float64, device=dev)
        a = _lib.FbankArgs()
        a.d_wav, a.wav_stride, a.d_nsamp, a.batch = _ptr(wav), wav.stride(0), _ptr(len_dev), B
        a.wav_dtype = 1 if wav.dtype == torch.int16 else 0
        a.d_peak = _ptr(peak)
        a.max_frames = int(T_host.max())
        a.d_stats, a.stats_stride, a.n_row_classes = _ptr(acc), 2 * D, 1
        if self.dither != 0.0:
            self.dither_seed += 1
            a.dither_seed = self.dither_seed
        if self.compact_tiles and not plan.uses_ws:
            cap = plan.lib.b200fe_tile_table_capacity(plan.handle, B, a.max_frames, 0)
            work = torch.empty((2 * cap + 2,), dtype=torch.int32, device=dev)              # table | n_tiles | counter
            _lib.check(plan.lib.b200fe_build_work_list_device(plan.handle, _ptr(len_dev), B, a.max_frames, 0, 0, _ptr(work), cap,
                                                              C.c_void_p(work.data_ptr() + 8 * cap), C.c_void_p(work.data_ptr() + 8 * cap + 4),
                                                              C.c_void_p(0), _ptr(acc), B * 2 * D * 8, stream),
                       "b200fe_build_work_list_device")
            a.d_tile_table, a.n_tiles = _ptr(work), cap
            a.d_n_tiles, a.d_work_counter = C.c_void_p(work.data_ptr() + 8 * cap), C.c_void_p(work.data_ptr() + 8 * cap + 4)
        elif self.compact_tiles:
            tab = _tile_table(T_host, plan.tile_frames)
            tot = tab.shape[0]
            tab_dev = _h2d(tab, dev)
            counter = torch.empty((1,), dtype=torch.int32, device=dev)
            a.d_tile_table, a.n_tiles, a.d_work_counter = _ptr(tab_dev), tot, _ptr(counter)
        _lib.check(plan.lib.b200fe_fbank_fused(plan.handle, C.byref(a), stream), "b200fe_fbank_fused")
        stats[:, :D] += acc.sum(0)
        stats[0, D] += float(T_host.sum())
        return stats

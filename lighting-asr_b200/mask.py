"""Encoder masks on the device: the consumer side of the front end.

The reference builds the source mask on the CPU from ``xlen.tolist()`` -- a host synchronisation per step
(lasr/model/e2e_ctc_att/e2e_base.py:19-20 -> lasr/utils/mask.py:5-45) -- and ``Conv2dSubsampling`` subsamples it
with ``x_mask[:, :, :-2:2][:, :, :-2:2]`` (lasr/modules/net/transformer/subsampling.py:60).  Here both masks and the
subsampled lengths (``E2E_CTC_ATT.subfunction``, e2e_base.py:47-49) come from one small kernel on the frame counts the
front end already left on the device.  Names and argument meaning follow the reference.
"""
import ctypes as C

import torch

from . import _lib


def _launch(lengths, max_length, subsample, plan, want_len):
    if not lengths.is_cuda:
        raise RuntimeError("lasr_b200.mask has no CPU path: lengths must be a CUDA tensor")
    lengths = lengths.to(torch.int64).contiguous()
    B = lengths.numel()
    Tout = max_length if subsample == 1 else (((max_length - 1) // 2 - 1) // 2 if max_length >= 7 else 0)
    mask = torch.empty((B, 1, max(Tout, 0)), dtype=torch.bool, device=lengths.device)
    out_len = torch.empty((B,), dtype=torch.int64, device=lengths.device) if want_len else None
    lib = _lib.load()
    stream = C.c_void_p(torch.cuda.current_stream(lengths.device).cuda_stream)
    _lib.check(lib.b200fe_src_mask(plan.handle if plan is not None else None, C.c_void_p(lengths.data_ptr()), 1 if plan is not None else 0,
                                   B, int(max_length), subsample, C.c_void_p(mask.data_ptr()),
                                   C.c_void_p(out_len.data_ptr()) if want_len else None, stream), "b200fe_src_mask")
    return mask, out_len


def make_pad_mask(lengths, xs=None, length_dim=-1, max_length=-1):
    """Device counterpart of ``lasr.utils.mask.make_pad_mask`` for the (B, Tmax) case: True on the padded part.
    ``lengths`` is a CUDA int tensor; ``max_length`` (or ``xs.size(length_dim)``) must be given -- computing the
    maximum on the host is exactly the synchronisation this function removes."""
    if length_dim == 0:
        raise ValueError('length_dim cannot be 0: {}'.format(length_dim))
    if xs is not None:
        if xs.dim() != 2 and not (xs.dim() == 3 and length_dim in (1, -2)):
            raise NotImplementedError("only (B, T) and (B, T, D) shapes are mirrored")
        maxlen = xs.size(length_dim)
    else:
        if max_length < 0:
            raise ValueError("max_length is required on the device (the reference takes max(lengths) on the host)")
        maxlen = int(max_length)
    m, _ = _launch(lengths, maxlen, 1, None, False)
    pad = ~m.squeeze(1)
    if xs is not None and xs.dim() == 3:
        pad = pad.unsqueeze(-1).expand_as(xs)
    return pad


def src_mask(feat_len, max_frames):
    """``(~make_pad_mask(xlen.tolist(), max_length=T)).unsqueeze(-2)`` of e2e_base.py:19-20: bool (B, 1, T)."""
    return _launch(feat_len, int(max_frames), 1, None, False)[0]


def subsampled_mask(feat_len, max_frames):
    """Mask and lengths after ``Conv2dSubsampling``: ``(src_mask[:, :, :-2:2][:, :, :-2:2], hs_len)`` where
    ``hs_len = sum(mask)`` (subsampling.py:60, e2e_base.py:47-49)."""
    return _launch(feat_len, int(max_frames), 4, None, True)


def src_mask_from_samples(plan, wav_len, max_frames, subsample=1):
    """The same masks straight from SAMPLE counts (frames derived with the plan's window / shift, TA:63-67)."""
    return _launch(wav_len, int(max_frames), subsample, plan, True)

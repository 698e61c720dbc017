"""ctypes binding of include/b200fe.h.  Fails loudly when the CUDA library is absent."""
import ctypes as C
import os

from . import build as _build

c_ll = C.c_longlong
c_fp = C.POINTER(C.c_float)


class Opts(C.Structure):
    _fields_ = [
        ("sample_frequency", C.c_float), ("frame_length_ms", C.c_float), ("frame_shift_ms", C.c_float),
        ("num_mel_bins", C.c_int), ("low_freq", C.c_float), ("high_freq", C.c_float),
        ("preemphasis_coefficient", C.c_float), ("remove_dc_offset", C.c_int), ("use_power", C.c_int),
        ("use_log_fbank", C.c_int), ("window_type", C.c_int), ("blackman_coeff", C.c_float),
        ("audio_bit", C.c_int), ("dither", C.c_float), ("window", C.c_void_p), ("mel_weights", C.c_void_p),
    ]


class _Sized(C.Structure):
    """Argument structs open with ``struct_size`` (= sizeof of the header's declaration): the library rejects a stale binding."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.struct_size = C.sizeof(self)


class FbankArgs(_Sized):
    _fields_ = [
        ("struct_size", C.c_uint), ("d_wav", C.c_void_p), ("wav_stride", c_ll), ("d_nsamp", C.c_void_p), ("batch", C.c_int),
        ("d_peak", C.c_void_p), ("d_out", C.c_void_p), ("d_out_len", C.c_void_p), ("max_frames", C.c_int),
        ("d_cmvn_mean", C.c_void_p), ("d_cmvn_istd", C.c_void_p), ("cmvn_stride", c_ll),
        ("d_masks", C.c_void_p), ("n_freq_masks", C.c_int), ("n_time_masks", C.c_int), ("mask_zero", C.c_int),
        ("d_stats", C.c_void_p), ("stats_stride", c_ll), ("d_row_bounds", C.c_void_p), ("n_row_classes", C.c_int),
        ("d_tile_table", C.c_void_p), ("n_tiles", C.c_int), ("d_work_counter", C.c_void_p),
        ("d_wav_offsets", C.c_void_p), ("offsets_aligned", C.c_int),
        ("dither_seed", C.c_ulonglong), ("d_dither_noise", C.c_void_p), ("wav_dtype", C.c_int), ("uniform_frames", C.c_int),
        ("d_out_offsets", C.c_void_p), ("tile_table_pads", C.c_int), ("d_n_tiles", C.c_void_p),
        ("apply_cmvn_mode", C.c_int), ("d_utt_done", C.c_void_p), ("d_utt_mean", C.c_void_p), ("d_utt_istd", C.c_void_p), ("d_fills", C.c_void_p),
    ]


class PostArgs(_Sized):
    _fields_ = [
        ("struct_size", C.c_uint), ("d_feats", C.c_void_p), ("d_nsamp", C.c_void_p), ("batch", C.c_int), ("max_frames", C.c_int),
        ("d_stats", C.c_void_p), ("stats_stride", c_ll), ("d_row_bounds", C.c_void_p), ("n_row_classes", C.c_int),
        ("cmvn_mode", C.c_int), ("d_cmvn_mean", C.c_void_p), ("d_cmvn_istd", C.c_void_p),
        ("d_masks", C.c_void_p), ("n_freq_masks", C.c_int), ("n_time_masks", C.c_int), ("d_fills", C.c_void_p),
        ("fill_zero", C.c_int), ("d_feat_offsets", C.c_void_p),
    ]


class WarpArgs(_Sized):
    _fields_ = [
        ("struct_size", C.c_uint), ("d_in", C.c_void_p), ("d_out", C.c_void_p), ("d_nsamp", C.c_void_p), ("batch", C.c_int), ("max_frames", C.c_int),
        ("d_warp", C.c_void_p), ("d_stats", C.c_void_p), ("stats_stride", c_ll), ("d_row_bounds", C.c_void_p),
        ("n_row_classes", C.c_int),
        ("d_masks", C.c_void_p), ("n_freq_masks", C.c_int), ("n_time_masks", C.c_int), ("d_fills", C.c_void_p), ("fill_zero", C.c_int), ("d_utt_done", C.c_void_p),
    ]


EXPORTS = [
    "b200fe_default_opts", "b200fe_plan_create", "b200fe_plan_destroy", "b200fe_last_error",
    "b200fe_window_size", "b200fe_window_shift", "b200fe_padded_window_size", "b200fe_num_frames", "b200fe_plan_info", "b200fe_build_tile_table", "b200fe_build_tile_table_padded", "b200fe_tile_table_capacity", "b200fe_build_tile_table_device", "b200fe_build_work_list_device",
    "b200fe_peak_absmax", "b200fe_peak_absmax_i16", "b200fe_fbank_fused", "b200fe_h2d_ragged", "b200fe_d2h_ragged", "b200fe_copy_ragged", "b200fe_src_mask", "b200fe_specaug_plan", "b200fe_postpass", "b200fe_time_warp", "b200fe_cmvn_from_stats",
    "b200fe_cast_bf16", "b200fe_copy_ragged_bf16", "b200fe_resample_poly", "b200fe_avg_channels",
    "b200fe_stream_create", "b200fe_stream_destroy", "b200fe_stream_max_frames", "b200fe_stream_reset", "b200fe_stream_push", "b200fe_stream_flags",
    "b200fe_host_pool_create", "b200fe_host_pool_destroy", "b200fe_host_pool_threads", "b200fe_host_isa", "b200fe_host_ndarray_data", "b200fe_host_pack_begin", "b200fe_host_pack_copy_begin", "b200fe_host_zero_rows_begin", "b200fe_host_zero_ranges_begin", "b200fe_host_wait", "b200fe_host_wait_flag", "b200fe_host_pcm16_probe",
]

_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Returns the loaded CDLL.  Raises RuntimeError if the CUDA library cannot be found or built:
    the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(path):
                raise RuntimeError("libb200fe.so is missing and could not be built (no CPU fallback exists): %s" % e)
    if not os.path.exists(path):
        raise RuntimeError("libb200fe.so is missing (run `python -m __graft_entry__` or lighting-asr_b200/build.py); "
                           "the front end has no CPU fallback")
    lib = C.CDLL(path)
    lib.b200fe_default_opts.argtypes = [C.POINTER(Opts)]
    lib.b200fe_default_opts.restype = None
    lib.b200fe_plan_create.argtypes = [C.POINTER(Opts), C.POINTER(C.c_void_p)]
    lib.b200fe_plan_create.restype = C.c_int
    lib.b200fe_plan_destroy.argtypes = [C.c_void_p]
    lib.b200fe_plan_destroy.restype = None
    lib.b200fe_last_error.argtypes = []
    lib.b200fe_last_error.restype = C.c_char_p
    for f in ("b200fe_window_size", "b200fe_window_shift", "b200fe_padded_window_size"):
        getattr(lib, f).argtypes = [C.c_void_p]
        getattr(lib, f).restype = C.c_int
    lib.b200fe_plan_info.argtypes = [C.c_void_p, C.c_int]
    lib.b200fe_plan_info.restype = C.c_int
    lib.b200fe_num_frames.argtypes = [C.c_void_p, c_ll]
    lib.b200fe_num_frames.restype = c_ll
    lib.b200fe_build_tile_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200fe_build_tile_table.restype = C.c_int
    lib.b200fe_build_tile_table_padded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.b200fe_build_tile_table_padded.restype = C.c_int
    lib.b200fe_tile_table_capacity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.b200fe_tile_table_capacity.restype = C.c_int
    lib.b200fe_build_tile_table_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200fe_build_tile_table_device.restype = C.c_int
    lib.b200fe_build_work_list_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_ll, C.c_void_p]
    lib.b200fe_build_work_list_device.restype = C.c_int
    lib.b200fe_peak_absmax.argtypes = [C.c_void_p, C.c_void_p, c_ll, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.b200fe_peak_absmax.restype = C.c_int
    lib.b200fe_peak_absmax_i16.argtypes = lib.b200fe_peak_absmax.argtypes
    lib.b200fe_peak_absmax_i16.restype = C.c_int
    lib.b200fe_fbank_fused.argtypes = [C.c_void_p, C.POINTER(FbankArgs), C.c_void_p]
    lib.b200fe_fbank_fused.restype = C.c_int
    lib.b200fe_h2d_ragged.argtypes = [C.c_void_p, c_ll, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200fe_h2d_ragged.restype = C.c_int
    lib.b200fe_d2h_ragged.argtypes = [C.c_void_p, c_ll, c_ll, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.b200fe_d2h_ragged.restype = C.c_int
    lib.b200fe_copy_ragged.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_ll, C.c_void_p]
    lib.b200fe_copy_ragged.restype = C.c_int
    lib.b200fe_src_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200fe_src_mask.restype = C.c_int
    lib.b200fe_specaug_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 7 + [C.c_void_p] * 3
    lib.b200fe_specaug_plan.restype = C.c_int
    lib.b200fe_postpass.argtypes = [C.c_void_p, C.POINTER(PostArgs), C.c_void_p]
    lib.b200fe_postpass.restype = C.c_int
    lib.b200fe_time_warp.argtypes = [C.c_void_p, C.POINTER(WarpArgs), C.c_void_p]
    lib.b200fe_time_warp.restype = C.c_int
    lib.b200fe_cmvn_from_stats.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, c_fp, c_fp]
    lib.b200fe_cmvn_from_stats.restype = C.c_int
    lib.b200fe_resample_poly.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_ll, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.b200fe_resample_poly.restype = C.c_int
    lib.b200fe_avg_channels.argtypes = [C.c_void_p, C.c_void_p, c_ll, C.c_int, C.c_void_p]
    lib.b200fe_avg_channels.restype = C.c_int
    lib.b200fe_cast_bf16.argtypes = [C.c_void_p, C.c_void_p, c_ll, C.c_void_p]
    lib.b200fe_cast_bf16.restype = C.c_int
    lib.b200fe_copy_ragged_bf16.argtypes = lib.b200fe_copy_ragged.argtypes
    lib.b200fe_copy_ragged_bf16.restype = C.c_int
    lib.b200fe_stream_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.b200fe_stream_create.restype = C.c_int
    lib.b200fe_stream_destroy.argtypes = [C.c_void_p]
    lib.b200fe_stream_destroy.restype = None
    lib.b200fe_stream_max_frames.argtypes = [C.c_void_p]
    lib.b200fe_stream_max_frames.restype = C.c_int
    lib.b200fe_stream_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200fe_stream_reset.restype = C.c_int
    lib.b200fe_stream_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, c_ll, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.b200fe_stream_push.restype = C.c_int
    lib.b200fe_stream_flags.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    lib.b200fe_stream_flags.restype = C.c_int
    lib.b200fe_host_pool_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.b200fe_host_pool_create.restype = C.c_int
    lib.b200fe_host_pool_destroy.argtypes = [C.c_void_p]
    lib.b200fe_host_pool_destroy.restype = None
    lib.b200fe_host_pool_threads.argtypes = [C.c_void_p]
    lib.b200fe_host_pool_threads.restype = C.c_int
    lib.b200fe_host_isa.argtypes = []
    lib.b200fe_host_isa.restype = C.c_int
    lib.b200fe_host_ndarray_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200fe_host_ndarray_data.restype = C.c_int
    lib.b200fe_host_pack_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_ll]
    lib.b200fe_host_pack_begin.restype = c_ll
    lib.b200fe_host_pack_copy_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_ll,
                                                C.c_void_p, c_ll, C.c_int, C.c_void_p, C.c_void_p]
    lib.b200fe_host_pack_copy_begin.restype = c_ll
    lib.b200fe_host_zero_rows_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_ll, c_ll, C.c_void_p, C.c_int]
    lib.b200fe_host_zero_rows_begin.restype = c_ll
    lib.b200fe_host_zero_ranges_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.b200fe_host_zero_ranges_begin.restype = c_ll
    lib.b200fe_host_wait.argtypes = [C.c_void_p, c_ll]
    lib.b200fe_host_wait.restype = C.c_int
    lib.b200fe_host_wait_flag.argtypes = [C.c_void_p, c_ll, C.c_void_p]
    lib.b200fe_host_wait_flag.restype = C.c_int
    lib.b200fe_host_pcm16_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.b200fe_host_pcm16_probe.restype = C.c_int
    _lib = lib
    return lib


class B200feError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        msg = load().b200fe_last_error().decode("utf-8", "replace")
        raise B200feError("%s failed (%d): %s" % (what, status, msg))

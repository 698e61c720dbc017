"""lighting-asr_b200: B200-native (sm_100a) acoustic front end for gaochangfeng/lighting-asr.

The directory name contains a hyphen (repository layout contract); import it with
``importlib.import_module("lighting-asr_b200")`` or through the ``lasr_b200`` alias module at
the repository root.  The hot path lives in ``csrc/`` (hand-written CUDA behind the C ABI of
``include/b200fe.h``); the Python here mirrors the reference's interfaces around it.
"""
from . import _lib, build, cmvn, lasr_plugin, mask, resample, specaug  # noqa: F401
from .frontend import FbankPlan, GpuFbankFrontend  # noqa: F401
from .streaming import IndependentStreams, StreamingFbank  # noqa: F401

__all__ = ["GpuFbankFrontend", "FbankPlan", "StreamingFbank", "IndependentStreams", "specaug", "cmvn", "mask", "build"]

// Instantiation list of fbank_fused_kernel<NLOAD, kStaticMel, kPeak, kDual, kI16, kMulti, kLean, kApply>.
// The list is split over several translation units (fbank_inst.cu compiled once per group with
// -DB200FE_INST_GROUP=g) so that the library builds in parallel; b200fe.cu sees the same list as
// `extern template` declarations.  X(group, NLOAD, static, peak, dual, i16, multi, lean, apply)
#pragma once

#define B200FE_FBANK_INSTANCES(X)                                   \
    /* lean default option set (BASELINE configs 2/4) + apply tiles */ \
    X(0, 13, true, false, false, false, false, true, false)         \
    X(0, 13, true, false, false, false, false, true, true)          \
    /* default option set, full epilogue (CMVN / masks / packed output), peak norm, multi-utterance tiles */ \
    X(1, 13, true, false, false, false, false, false, false)        \
    X(1, 13, true, true, false, false, false, false, false)         \
    X(2, 13, true, false, false, false, true, false, false)         \
    X(2, 13, true, false, false, true, false, false, false)         \
    X(3, 13, true, true, false, true, false, false, false)          \
    /* generic mel tables, 512-point family */                      \
    X(3, 13, false, false, false, false, false, false, false)       \
    X(4, 13, false, true, false, false, false, false, false)        \
    /* default option set, full epilogue + completion tiles (SpecAugment mean fills inside the launch) */ \
    X(4, 13, true, false, false, false, false, false, true)         \
    X(4, 16, false, false, false, false, false, false, false)       \
    X(5, 16, false, true, false, false, false, false, false)        \
    X(5, 13, false, false, false, true, false, false, false)        \
    X(6, 13, false, true, false, true, false, false, false)         \
    X(6, 16, false, false, false, true, false, false, false)        \
    X(6, 16, false, true, false, true, false, false, false)         \
    /* 8 kHz family (two real frames per complex FFT) */            \
    X(7, 13, false, false, true, false, false, false, false)        \
    X(7, 13, false, true, true, false, false, false, false)         \
    X(7, 16, false, false, true, false, false, false, false)        \
    X(7, 16, false, true, true, false, false, false, false)         \
    X(7, 13, false, false, true, false, true, false, false)

#define B200FE_INST_GROUPS 8

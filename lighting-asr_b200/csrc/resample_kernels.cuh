// Waveform ingest on the device: rational-ratio polyphase resampling of a ragged batch.
//
// The reference resamples per utterance on the CPU with librosa's `kaiser_fast` (R/lasr/data/datatrans.py:16-20 -- resampy's
// precomputed Kaiser-windowed sinc table) and perturbs speed with sox (datatrans.py:29-39: `speed` = resample by 1 / ratio and
// keep the nominal rate).  Neither library is in this image, so their outputs cannot be pinned; the filter comes from the host
// (lighting-asr_b200/resample.py): either librosa's path restated -- resampy's interpolation with its `kaiser_fast` window, which
// for a rational ratio is one fixed FIR per output phase -- or the one of scipy.signal.resample_poly (firwin, Kaiser beta = 5,
// half length 10 * max(up, down)), checked against scipy itself (tests/test_gpu_resample.py).  Index arithmetic of
// scipy.signal.upfirdn / resample_poly:
//     out[m] = sum_j h[(m + pre_remove) * down - j * up] * x[j]        (h zero outside [0, len))
// One thread per output sample; the filter phase (t mod up) is warp-divergent only in its start index, every lane walks
// taps = ceil(len / up) input samples of the same neighbourhood (coalesced), the taps come through L1.  HBM-bound:
// 4 B in per input sample + 4 B out per output sample.
#pragma once
#include "b200fe_common.cuh"

namespace b200fe {

__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ in, const long long* __restrict__ in_off,
                                                            const long long* __restrict__ n_in, float* __restrict__ out,
                                                            const long long* __restrict__ out_off, const long long* __restrict__ n_out,
                                                            const float* __restrict__ h, int hlen, int up, int down, int pre_remove, float scale)
{
    const int u = blockIdx.y;
    const long long ni = n_in[u], no = n_out[u];
    const float* x = in + in_off[u];
    float* y = out + out_off[u];
    for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < no; m += (long long)gridDim.x * 256) {
        const long long t = (m + pre_remove) * (long long)down;
        // j ranges over t - j * up in [0, hlen): j <= t / up, j >= (t - hlen + 1 + up - 1) / up
        long long j_hi = t / up;
        long long j_lo = (t - hlen + up) / up;               // ceil((t - hlen + 1) / up) for non-negative numerators
        if (t - hlen + 1 <= 0) j_lo = 0;
        if (j_hi > ni - 1) j_hi = ni - 1;
        float acc = 0.f;
        for (long long j = j_lo; j <= j_hi; ++j) acc = fmaf(__ldg(h + (t - j * up)), x[j], acc);
        y[m] = acc * scale;
    }
}

// Channel average of interleaved (N, C) input on the device: y[i] = mean_c x[i][c]  (AverageChanl, R/lasr/data/datatrans.py:10-14).
__global__ void __launch_bounds__(256) avg_channels_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int channels)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double s = 0.0;
        for (int c = 0; c < channels; ++c) s += (double)in[i * channels + c];
        out[i] = (float)(s / channels);
    }
}

}  // namespace b200fe

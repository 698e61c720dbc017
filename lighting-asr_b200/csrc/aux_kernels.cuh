// Auxiliary kernels: peak abs-max, statistics finalisation (utterance CMVN vectors and
// SpecAugment mean fills) and the in-place post pass (CMVN apply + mask fill).
#pragma once
#include "b200fe_common.cuh"
#include "fbank_kernel.cuh"

namespace b200fe {

// ---- VoiceNorm abs-max (R/lasr/data/datatrans.py:24): one float per utterance -----------------
// |x| >= 0, so the IEEE bit pattern orders like an unsigned integer and atomicMax is exact.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ wav, long long stride,
                                                     const long long* __restrict__ nsamp, float* __restrict__ peak)
{
    const int utt = blockIdx.y;
    const long long n = nsamp[utt];
    const float* x = wav + (long long)utt * stride;
    const long long chunk = 256LL * 4 * 8;
    long long i0 = (long long)blockIdx.x * chunk;
    if (i0 >= n) return;
    float m = 0.f;
    const bool al = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (al) {
        for (int r = 0; r < 8; ++r) {
            long long i = i0 + ((long long)r * 256 + threadIdx.x) * 4;
            if (i + 3 < n) {
                float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            } else {
                for (long long q = i; q < n && q < i + 4; ++q) m = fmaxf(m, fabsf(__ldg(x + q)));
            }
        }
    } else {
        for (long long i = i0 + threadIdx.x; i < n && i < i0 + chunk; i += 256) m = fmaxf(m, fabsf(__ldg(x + i)));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
        atomicMax(reinterpret_cast<unsigned int*>(peak + utt), __float_as_uint(m));
    }
}

struct PostArgs {
    float* feats;
    const long long* nsamp;
    int B, Tmax, nmel, win, shift;
    const double* stats;
    long long stats_stride;
    const int* row_bounds;
    int n_cls;
    int cmvn_mode;
    float* cm_mean;
    float* cm_istd;
    const int* masks;
    int n_fmask, n_tmask;
    float* fills;
    int rows_per_cta;
};

// One CTA (128 threads, thread d = mel column d) per utterance.
__global__ void __launch_bounds__(128) finalize_kernel(const PostArgs a)
{
    const int utt = blockIdx.x, d = threadIdx.x;
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int nb = a.n_cls - 1;
    const int* bounds = (a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
    const double* sb = a.stats + (long long)utt * a.stats_stride;
    __shared__ double red[4];
    __shared__ int cls_lo[kMaxRowClasses + 1];
    if (d == 0) {
        cls_lo[0] = 0;
        for (int c = 0; c < nb; ++c) cls_lo[c + 1] = min(max(bounds ? bounds[c] : T, 0), T);
        cls_lo[a.n_cls] = T;
    }
    __syncthreads();
    double S[kMaxRowClasses];
    double mean = 0.0, istd = 1.0;
    const bool col = d < a.nmel;
    if (col) {
        double tot = 0.0;
        for (int c = 0; c < a.n_cls; ++c) { S[c] = sb[(long long)c * a.nmel + d]; tot += S[c]; }
        if (a.cmvn_mode != 0 && T > 0) {
            mean = tot / T;
            if (a.cmvn_mode == 2) {
                double var = sb[(long long)a.n_cls * a.nmel + d] / T - mean * mean;
                istd = 1.0 / sqrt(var > 1e-20 ? var : 1e-20);
            }
            a.cm_mean[(long long)utt * a.nmel + d] = (float)mean;
            a.cm_istd[(long long)utt * a.nmel + d] = (float)istd;
            // class sums in the normalised domain, using the float32 vectors the apply pass uses
            const double mf = (double)(float)mean, sf = (double)(float)istd;
            for (int c = 0; c < a.n_cls; ++c) S[c] = (S[c] - (double)(cls_lo[c + 1] - cls_lo[c]) * mf) * sf;
        }
    } else {
        for (int c = 0; c < a.n_cls; ++c) S[c] = 0.0;
    }
    if (a.masks == nullptr) return;

    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((d & 31) == 0) red[d >> 5] = v;
        __syncthreads();
        return red[0] + red[1] + red[2] + red[3];
    };
    double part = 0.0;
    for (int c = 0; c < a.n_cls; ++c) part += S[c];
    double total = block_sum(part);
    const int nm = a.n_fmask + a.n_tmask;
    const int* mk = a.masks + (long long)utt * nm * 2;
    const double cells = (double)T * (double)a.nmel;
    for (int i = 0; i < nm; ++i) {
        // numpy: x[...] = x.mean() evaluated on the current array (specaugment.py:74,105)
        const double fill = cells > 0 ? (double)(float)(total / cells) : 0.0;
        int lo = mk[2 * i], hi = mk[2 * i + 1];
        double delta = 0.0;
        if (i < a.n_fmask) {
            lo = max(lo, 0); hi = min(hi, a.nmel);
            if (col && d >= lo && d < hi)
                for (int c = 0; c < a.n_cls; ++c) {
                    const double nv = fill * (double)(cls_lo[c + 1] - cls_lo[c]);
                    delta += nv - S[c]; S[c] = nv;
                }
        } else {
            lo = max(lo, 0); hi = min(hi, T);
            if (col)
                for (int c = 0; c < a.n_cls; ++c)
                    if (cls_lo[c] >= lo && cls_lo[c + 1] <= hi && cls_lo[c + 1] > cls_lo[c]) {
                        const double nv = fill * (double)(cls_lo[c + 1] - cls_lo[c]);
                        delta += nv - S[c]; S[c] = nv;
                    }
        }
        total += block_sum(delta);
        if (d == 0) a.fills[(long long)utt * nm + i] = (float)fill;
    }
}

// In-place post pass: rows of one utterance, thread = (row, column).  Applies utterance CMVN and
// overwrites masked cells with the fill of the LAST mask that covers them (later masks overwrite
// earlier ones, specaugment.py applies them sequentially).
__global__ void __launch_bounds__(256) postpass_kernel(const PostArgs a)
{
    const int utt = blockIdx.y;
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int r0 = blockIdx.x * a.rows_per_cta;
    if (r0 >= T) return;
    const int r1 = min(r0 + a.rows_per_cta, T);
    const int nm = a.n_fmask + a.n_tmask;
    __shared__ int s_mk[2 * (kMaxFreqMasks + kMaxTimeMasks)];
    __shared__ float s_fill[kMaxFreqMasks + kMaxTimeMasks];
    __shared__ float s_mean[kMaxMel], s_istd[kMaxMel];
    if (threadIdx.x < 2 * nm && a.masks) s_mk[threadIdx.x] = a.masks[(long long)utt * nm * 2 + threadIdx.x];
    if (threadIdx.x < nm && a.masks) s_fill[threadIdx.x] = a.fills[(long long)utt * nm + threadIdx.x];
    if (a.cmvn_mode != 0 && threadIdx.x < a.nmel) {
        s_mean[threadIdx.x] = a.cm_mean[(long long)utt * a.nmel + threadIdx.x];
        s_istd[threadIdx.x] = a.cm_istd[(long long)utt * a.nmel + threadIdx.x];
    }
    __syncthreads();
    float* base = a.feats + (long long)utt * a.Tmax * a.nmel;
    const int total = (r1 - r0) * a.nmel;
    for (int e = threadIdx.x; e < total; e += 256) {
        const int r = r0 + e / a.nmel, d = e % a.nmel;
        int hit = -1;
        if (a.masks) {
            for (int i = 0; i < a.n_fmask; ++i) if (d >= s_mk[2 * i] && d < s_mk[2 * i + 1]) hit = i;
            for (int i = a.n_fmask; i < nm; ++i) if (r >= s_mk[2 * i] && r < s_mk[2 * i + 1]) hit = i;
        }
        float* ptr = base + (long long)r * a.nmel + d;
        if (hit >= 0) *ptr = s_fill[hit];
        else if (a.cmvn_mode != 0) *ptr = (*ptr - s_mean[d]) * s_istd[d];
    }
}

}  // namespace b200fe

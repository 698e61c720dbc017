// Auxiliary kernels: peak abs-max, statistics finalisation (utterance CMVN vectors and
// SpecAugment mean fills) and the in-place post pass (CMVN apply + mask fill).
#pragma once
#include <cuda_bf16.h>
#include "b200fe_common.cuh"
#include "fbank_kernel.cuh"

namespace b200fe {

// ---- VoiceNorm abs-max (R/lasr/data/datatrans.py:24): one float per utterance -----------------
// |x| >= 0, so the IEEE bit pattern orders like an unsigned integer and atomicMax is exact.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ wav, long long stride, const long long* __restrict__ offsets,
                                                     const long long* __restrict__ nsamp, float* __restrict__ peak)
{
    const int utt = blockIdx.y;
    const long long n = nsamp[utt];
    const float* x = wav + (offsets ? offsets[utt] : (long long)utt * stride);
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

// Rows past an utterance's frame count are zero (pad_audio = 0, dataset.py:18,205): float4 stores over
// [T_u * nmel, Tmax * nmel) of every utterance (used with the compact tile list).
__global__ void __launch_bounds__(256) zero_pad_kernel(float* __restrict__ out, const long long* __restrict__ nsamp, int Tmax, int nmel,
                                                       int win, int shift)
{
    const int utt = blockIdx.y;
    const long long n = nsamp[utt];
    const long long T = n >= win ? (1 + (n - win) / shift) : 0;
    const long long lo = T * nmel, hi = (long long)Tmax * nmel;
    const long long c0 = (long long)blockIdx.x * (256 * 4 * 8);
    if (c0 + 256 * 4 * 8 <= lo || lo >= hi) return;
    float* base = out + (long long)utt * hi;
    if ((nmel & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long e = c0 + ((long long)r * 256 + threadIdx.x) * 4;
            if (e >= lo && e + 3 < hi) *reinterpret_cast<float4*>(base + e) = make_float4(0.f, 0.f, 0.f, 0.f);
            else for (long long q = e; q < e + 4; ++q) if (q >= lo && q < hi) base[q] = 0.f;
        }
    } else {
        for (long long e = c0 + threadIdx.x; e < c0 + 256 * 4 * 8 && e < hi; e += 256) if (e >= lo) base[e] = 0.f;
    }
}

// Ragged row copy by a kernel: row u = nbytes[u] bytes from src + src_off[u] to dst + dst_off[u].  Either side may be
// pinned (mapped) host memory: the SMs then read / write it over PCIe directly, so a whole batch of variable-length
// utterances moves in ONE launch instead of one DMA set-up per utterance (measured on B200 / PCIe 5 x16: 256 x 1.15 MB
// cudaMemcpyAsync 46.7 GB/s, one contiguous copy 55.5 GB/s).  Persistent grid over (row, 32 kB chunk) pairs.
constexpr int kCopyThreads = 256;
constexpr int kCopyChunk = kCopyThreads * 16 * 8;
__global__ void __launch_bounds__(kCopyThreads) ragged_copy_kernel(const char* __restrict__ src, const long long* __restrict__ src_off,
                                                                   char* __restrict__ dst, const long long* __restrict__ dst_off,
                                                                   const long long* __restrict__ nbytes, int B, int chunks_per_row)
{
    const long long total = (long long)B * chunks_per_row;
    for (long long v = blockIdx.x; v < total; v += gridDim.x) {
        const int u = (int)(v / chunks_per_row);
        const long long c0 = (v - (long long)u * chunks_per_row) * kCopyChunk;
        const long long nb = nbytes[u];
        if (c0 >= nb) continue;
        const char* s = src + src_off[u];
        char* d = dst + dst_off[u];
        const long long lim = min(nb, c0 + (long long)kCopyChunk);
        if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
            uint4 vreg[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const long long off = c0 + ((long long)r * kCopyThreads + threadIdx.x) * 16;
                if (off + 16 <= lim) vreg[r] = __ldcs(reinterpret_cast<const uint4*>(s + off));
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const long long off = c0 + ((long long)r * kCopyThreads + threadIdx.x) * 16;
                if (off + 16 <= lim) __stcs(reinterpret_cast<uint4*>(d + off), vreg[r]);
                else if (off < lim) for (long long q = off; q < lim; ++q) d[q] = s[q];      // < 16 trailing bytes of the row
            }
        } else if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)nb) & 3) == 0) {
            for (long long off = c0 + threadIdx.x * 4; off < lim; off += kCopyThreads * 4)
                *reinterpret_cast<unsigned*>(d + off) = *reinterpret_cast<const unsigned*>(s + off);
        } else {
            for (long long off = c0 + threadIdx.x; off < lim; off += kCopyThreads) d[off] = s[off];
        }
    }
}

// ---- bfloat16 feature emission (the consumer's precision) --------------------------------------------
// The encoder's first layer (Conv2dSubsampling, R/lasr/modules/net/transformer/subsampling.py:53-57) runs in bfloat16 under
// autocast; handing it bfloat16 features halves the bytes of the hand-off (and of the D2H copy when features go to the host).
// Round to nearest even, the same rounding `tensor.to(torch.bfloat16)` applies.
__device__ __forceinline__ uint2 pack_bf16x4(float4 v)
{
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const unsigned*>(&lo);
    r.y = *reinterpret_cast<const unsigned*>(&hi);
    return r;
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n)
{
    const long long n4 = n >> 2;
    const float4* i4 = reinterpret_cast<const float4*>(in);
    uint2* o4 = reinterpret_cast<uint2*>(out);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) o4[i] = pack_bf16x4(__ldcs(i4 + i));
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = __float2bfloat16_rn(in[i]);
}

// Ragged row copy float32 -> bfloat16: row u = nbytes[u] bytes of float32 from src + src_off[u] to bfloat16 at dst + dst_off[u]
// (byte offsets, 16-byte aligned rows, nbytes a multiple of 16).  The destination may be pinned host memory: the D2H copy of
// the features and their conversion are one pass, and only half the bytes cross PCIe.
__global__ void __launch_bounds__(kCopyThreads) ragged_cast_bf16_kernel(const char* __restrict__ src, const long long* __restrict__ src_off,
                                                                        char* __restrict__ dst, const long long* __restrict__ dst_off,
                                                                        const long long* __restrict__ nbytes, int B, int chunks_per_row)
{
    const long long total = (long long)B * chunks_per_row;
    for (long long v = blockIdx.x; v < total; v += gridDim.x) {
        const int u = (int)(v / chunks_per_row);
        const long long c0 = (v - (long long)u * chunks_per_row) * kCopyChunk;
        const long long nb = nbytes[u];
        if (c0 >= nb) continue;
        const char* s = src + src_off[u];
        char* d = dst + dst_off[u];
        const long long lim = min(nb, c0 + (long long)kCopyChunk);
        float4 vreg[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long off = c0 + ((long long)r * kCopyThreads + threadIdx.x) * 16;
            if (off + 16 <= lim) vreg[r] = __ldcs(reinterpret_cast<const float4*>(s + off));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long off = c0 + ((long long)r * kCopyThreads + threadIdx.x) * 16;
            if (off + 16 <= lim) __stcs(reinterpret_cast<uint2*>(d + (off >> 1)), pack_bf16x4(vreg[r]));
        }
    }
}

// Encoder source mask on the device (replaces the host round trip `make_pad_mask(xlen.tolist(), max_length=T)`,
// lasr/model/e2e_ctc_att/e2e_base.py:19-20 -> lasr/utils/mask.py:5-45): mask[b][0][i] = (i * step < frames[b]) for
// i < Tout, where step = 1 is the encoder input mask and step = 4, Tout = ((T-1)/2-1)/2 is what Conv2dSubsampling
// hands on (x_mask[:, :, :-2:2][:, :, :-2:2], lasr/modules/net/transformer/subsampling.py:60).  out_len[b] = number of
// true cells (E2E_CTC_ATT.subfunction, e2e_base.py:47-49).  frames come from sample counts when win > 0.
__global__ void __launch_bounds__(256) src_mask_kernel(const long long* __restrict__ len, int win, int shift, int Tout, int step,
                                                       unsigned char* __restrict__ mask, long long* __restrict__ out_len)
{
    const int b = blockIdx.y;
    long long T = len[b];
    if (win > 0) T = T >= win ? 1 + (T - win) / shift : 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < Tout; i += gridDim.x * 256)
        mask[(long long)b * Tout + i] = ((long long)i * step < T) ? 1 : 0;
    if (out_len != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        const long long cnt = T <= 0 ? 0 : (T + step - 1) / step;
        out_len[b] = cnt < Tout ? cnt : Tout;
    }
}

// Work list of the fused launch built on the device from the sample counts (same order and entries as
// b200fe_build_tile_table[_padded]): per utterance its frame tiles (utt, first frame) and, with pads, its padding tiles
// (utt, -(row0 + 1)).  One CTA: block-wide exclusive scan of the per-utterance entry counts, then every thread writes the
// entries of its utterances.  Also publishes the tile count and resets the work counter of the dynamic scheduler, so a call
// needs no host-side table, no table upload and no memset node.  Launched with a few CTAs: every CTA repeats the (cheap) scan and
// writes its share of the entries; the optional buffer `zero16` (the statistics accumulators of the same call) is cleared by all
// of them, which saves the memset node in front of the fused launch.
// apply_lag > 0 (b200fe_build_work_list_device): slot u additionally carries the CMVN-apply tiles (utt | apply_bit, row0) of
// utterance u - apply_lag, apply_lag extra slots close the list, and the per-utterance completion counters are zeroed.
__global__ void __launch_bounds__(1024) build_tile_table_kernel(const long long* __restrict__ nsamp, int B, int win, int shift, int ft,
                                                                int Tmax, int pads, int pad_rows, int2* __restrict__ table, int capacity,
                                                                int* __restrict__ ntiles_out, int* __restrict__ counter,
                                                                int apply_lag, int apply_bit, int apply_rows, int* __restrict__ utt_done,
                                                                uint4* __restrict__ zero16, long long n_zero16)
{
    __shared__ int warp_sums[32];
    __shared__ int s_off[1025];          // exclusive offsets of the chunk's utterances inside the chunk (+ total)
    __shared__ int s_nt[1024], s_T[1024], s_np[1024];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_launch_dependents();      // the fused launch may stage its plan constants while the list is being built
    pdl_wait();                   // the buffers written below may still be read by the previous call's post pass
    if (tid == 0) { s_base = 0; if (counter && blockIdx.x == 0) *counter = 0; }
    if (utt_done && blockIdx.x == 0) for (int u = tid; u <= B; u += 1024) utt_done[u] = 0;      // [B] = error flag of the apply tiles
    for (long long i = (long long)blockIdx.x * 1024 + tid; i < n_zero16; i += 1024LL * gridDim.x) zero16[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    auto frames_of = [&](int u) -> int {
        const long long n = nsamp[u];
        long long Tl = n >= win ? 1 + (n - win) / shift : 0;
        if (Tl > Tmax) Tl = Tmax;
        return (int)Tl;
    };
    const int nslots = B + (apply_lag > 0 ? apply_lag : 0);
    for (int u0 = 0; u0 < nslots; u0 += 1024) {
        const int u = u0 + tid;
        int T = 0, nt = 0, np_ = 0, na = 0;
        if (u < B) {
            T = frames_of(u);
            nt = (T + ft - 1) / ft;
            np_ = pads ? (max(Tmax - T, 0) + pad_rows - 1) / pad_rows : 0;
        }
        if (apply_lag > 0 && u >= apply_lag && u - apply_lag < B) na = (frames_of(u - apply_lag) + apply_rows - 1) / apply_rows;
        const int cnt = nt + np_ + na;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
            warp_sums[lane] = w;                     // inclusive scan of the warp totals
        }
        __syncthreads();
        const int excl = (wid > 0 ? warp_sums[wid - 1] : 0) + incl - cnt;
        s_off[tid] = excl; s_nt[tid] = nt; s_T[tid] = T; s_np[tid] = np_;
        if (tid == 1023) s_off[1024] = excl + cnt;
        __syncthreads();
        // cooperative fill: consecutive threads write consecutive entries (utterance found by binary search over the offsets)
        const int total = s_off[1024], base = s_base;
        for (int e = tid + 1024 * (int)blockIdx.x; e < total; e += 1024 * (int)gridDim.x) {
            int lo = 0, hi = 1023;                   // last j with s_off[j] <= e
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= e) lo = mid; else hi = mid - 1; }
            const int k = e - s_off[lo], ntj = s_nt[lo], npj = s_np[lo];
            const int idx = base + e;
            if (idx < capacity)
                table[idx] = k < ntj ? make_int2(u0 + lo, k * ft)
                           : k < ntj + npj ? make_int2(u0 + lo, -(s_T[lo] + (k - ntj) * pad_rows) - 1)
                                           : make_int2((u0 + lo - apply_lag) | apply_bit, (k - ntj - npj) * apply_rows);
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0 && blockIdx.x == 0) *ntiles_out = min(s_base, capacity);
}

// int16 PCM abs-max: peak = max |s16| / 2^15 (what max |x| is after soundfile's conversion)
__global__ void __launch_bounds__(256) absmax_i16_kernel(const short* __restrict__ wav, long long stride, const long long* __restrict__ offsets,
                                                         const long long* __restrict__ nsamp, float* __restrict__ peak)
{
    const int utt = blockIdx.y;
    const long long n = nsamp[utt];
    const short* x = wav + (offsets ? offsets[utt] : (long long)utt * stride);
    const long long chunk = 256LL * 4 * 8;
    const long long i0 = (long long)blockIdx.x * chunk;
    if (i0 >= n) return;
    int m = 0;
    for (long long i = i0 + threadIdx.x; i < n && i < i0 + chunk; i += 256) { const int v = x[i]; m = max(m, v < 0 ? -v : v); }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ int sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = max(m, sm[w]);
        atomicMax(reinterpret_cast<unsigned int*>(peak + utt), __float_as_uint((float)m * 3.0517578125e-05f));
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
    int fill_zero;               // 1: replace_with_zero (fills are 0), 0: running mean fills
    int rows_per_cta;
    const long long* feat_offsets;   // optional [B] first row of every utterance (packed features)
    int inline_finalize;             // utterance CMVN without masks: the post pass derives the vectors itself (no finalize launch)
};

// One CTA of NT threads (thread d = mel column d; d >= num_mel_bins only helps in the reductions) per utterance.  Also the tail of
// time_warp_kernel (NT = 256), which is why the statistics are read past L1.
template <int NT>
__device__ __forceinline__ void finalize_body(const PostArgs& a, const int utt, const int d)
{
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int nb = a.n_cls - 1;
    const int* bounds = (a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
    const double* sb = a.stats + (long long)utt * a.stats_stride;
    __shared__ double red[NT / 32];
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
        for (int c = 0; c < a.n_cls; ++c) { S[c] = __ldcg(sb + (long long)c * a.nmel + d); tot += S[c]; }
        if (a.cmvn_mode != 0 && T > 0) {
            mean = tot / T;
            if (a.cmvn_mode == 2) {
                double var = __ldcg(sb + (long long)a.n_cls * a.nmel + d) / T - mean * mean;
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
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) t += red[w];
        return t;
    };
    double part = 0.0;
    for (int c = 0; c < a.n_cls; ++c) part += S[c];
    double total = block_sum(part);
    const int nm = a.n_fmask + a.n_tmask;
    const int* mk = a.masks + (long long)utt * nm * 2;
    const double cells = (double)T * (double)a.nmel;
    for (int i = 0; i < nm; ++i) {
        // numpy: x[...] = x.mean() evaluated on the current array (specaugment.py:74,105)
        const double fill = (cells > 0 && !a.fill_zero) ? (double)(float)(total / cells) : 0.0;
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

__global__ void __launch_bounds__(128) finalize_kernel(const PostArgs a) { finalize_body<128>(a, blockIdx.x, threadIdx.x); }

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
    float* base = a.feats + (a.feat_offsets != nullptr ? a.feat_offsets[utt] : (long long)utt * a.Tmax) * a.nmel;
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

// ---- SpecAugment time warp ----------------------------------------------------
// R/lasr/utils/specaugment.py:15-32: rows [0, center) are resized to `warped` rows and rows [center, T)
// to T - warped rows with PIL's BICUBIC filter.  Pillow (12.2, src/libImaging/Resample.c) evaluates this
// for mode 'F' as: per output row, float64 coefficients bicubic((y + ymin - c + 0.5) / filterscale),
// normalised by their sum; float64 accumulation of pixel * coefficient in source order; one rounding to
// float32.  The kernel repeats exactly that arithmetic (explicit _rn operations, no FMA contraction), so
// the result is bit-identical to the reference's.
struct WarpArgs {
    const float* in;
    float* out;
    const long long* nsamp;
    int B, Tmax, nmel, win, shift;
    const int* warp;             // [B][2] (center, warped); center < 0 = utterance too short, rows are copied
    double* stats;               // optional row-class column sums + sums of squares of the WARPED features
    long long stats_stride;
    const int* row_bounds;
    int n_cls;
    int win_rows;                // source rows the dynamic shared memory holds as float64 (0 = no staging)
    const int* masks;            // optional: SpecAugment masks applied by the CTA that completes an utterance (finalize + fill in this launch)
    int n_fmask, n_tmask;
    float* fills;
    int fill_zero;
    int* done;                   // [B] completion counters, zero on entry, zero again on exit
    int part_ok;                 // the dynamic shared memory holds 2 * min(1024 / nmel, kWarpRows) * nmel floats for the statistics partials
};

__device__ __forceinline__ double pil_bicubic(double x)
{
    x = fabs(x);
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
    return 0.0;
}

// Tail of time_warp_kernel for the CTA that completes an utterance: fills as finalize_kernel derives them, then the masked cells.
// (As a __noinline__ call it slowed the whole kernel -- C3 full step 0.552 -> 0.596 ms: 240 B of stack frame -- so it is inlined.)
__device__ __forceinline__ void warp_mask_tail(const WarpArgs& a, const int utt, const int T, float* dst, const bool vec)
{
    const int tid = threadIdx.x;
    const int nq = a.nmel >> 2;
    PostArgs q;
    q.feats = dst; q.nsamp = a.nsamp; q.B = a.B; q.Tmax = a.Tmax; q.nmel = a.nmel; q.win = a.win; q.shift = a.shift;
    q.stats = a.stats; q.stats_stride = a.stats_stride; q.row_bounds = a.row_bounds; q.n_cls = a.n_cls;
    q.cmvn_mode = 0; q.cm_mean = nullptr; q.cm_istd = nullptr;
    q.masks = a.masks; q.n_fmask = a.n_fmask; q.n_tmask = a.n_tmask; q.fills = a.fills; q.fill_zero = a.fill_zero;
    q.rows_per_cta = 0; q.feat_offsets = nullptr; q.inline_finalize = 0;
    finalize_body<256>(q, utt, tid);
    __syncthreads();                                               // the fills (written by thread 0) are visible to the CTA
    {   // the statistics are a workspace in this mode: zero on entry, zero again on exit (no fill launch per call)
        double* sbz = a.stats + (long long)utt * a.stats_stride;
        for (int i = tid; i < (a.n_cls + 1) * a.nmel; i += 256) sbz[i] = 0.0;
    }
    const int nm = a.n_fmask + a.n_tmask;
    const int* mk = a.masks + (long long)utt * nm * 2;
    const float* fl = a.fills + (long long)utt * nm;
    int tlo[kMaxTimeMasks], thi[kMaxTimeMasks];
    float tfill[kMaxTimeMasks];
#pragma unroll
    for (int i = 0; i < kMaxTimeMasks; ++i) {
        tlo[i] = 0; thi[i] = 0; tfill[i] = 0.f;
        if (i < a.n_tmask) { tlo[i] = mk[2 * (a.n_fmask + i)]; thi[i] = mk[2 * (a.n_fmask + i) + 1]; tfill[i] = fl[a.n_fmask + i]; }
    }
    if (vec) {
        const int slots = 256 / nq, slot = tid / nq, qd = tid - slot * nq;
        if (slot >= slots) return;
        int fhit[4] = {-1, -1, -1, -1};
        float ffill[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < a.n_fmask; ++i) {
            const int lo = mk[2 * i], hi = mk[2 * i + 1];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * qd + c >= lo && 4 * qd + c < hi) { fhit[c] = i; ffill[c] = fl[i]; }
        }
        const bool anyf = fhit[0] >= 0 || fhit[1] >= 0 || fhit[2] >= 0 || fhit[3] >= 0;
        const bool allf = fhit[0] >= 0 && fhit[1] >= 0 && fhit[2] >= 0 && fhit[3] >= 0;
        for (int r = slot; r < T; r += slots) {
            int thit = -1; float tf = 0.f;
#pragma unroll
            for (int i = 0; i < kMaxTimeMasks; ++i) if (r >= tlo[i] && r < thi[i]) { thit = i; tf = tfill[i]; }
            float* ptr = dst + (long long)r * a.nmel + 4 * qd;
            if (thit >= 0) *reinterpret_cast<float4*>(ptr) = make_float4(tf, tf, tf, tf);
            else if (allf) *reinterpret_cast<float4*>(ptr) = make_float4(ffill[0], ffill[1], ffill[2], ffill[3]);
            else if (anyf) {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (fhit[c] >= 0) ptr[c] = ffill[c];      // partial float4: no read-modify-write, the other cells stay as they are
            }
        }
    } else {
        for (long long e = tid; e < (long long)T * a.nmel; e += 256) {
            const int r = (int)(e / a.nmel), d = (int)(e - (long long)r * a.nmel);
            int hit = -1; float f = 0.f;
            for (int i = 0; i < a.n_fmask; ++i) if (d >= mk[2 * i] && d < mk[2 * i + 1]) { hit = i; f = fl[i]; }
#pragma unroll
            for (int i = 0; i < kMaxTimeMasks; ++i) if (r >= tlo[i] && r < thi[i]) { hit = a.n_fmask + i; f = tfill[i]; }
            if (hit >= 0) dst[e] = f;
        }
    }
}

#ifndef B200FE_WARP_ROWS
#define B200FE_WARP_ROWS 64       // C3 full specaug step on B200: 32 rows 0.595 ms, 48 rows 0.630, 64 rows 0.576 (v5, 5 CTAs per SM)
#endif
#ifndef B200FE_WARP_OCC
#define B200FE_WARP_OCC 5
#endif
constexpr int kWarpRows = B200FE_WARP_ROWS;    // output rows per CTA
constexpr int kWarpTaps = 16;    // taps per output row: xmax - xmin <= min(in_size, 2 * support + 1); with |in - out| <= W (max_time_warp, 5 in the
                                 // reference) a scale above 2 needs out < W, i.e. in < 2 W, so 16 taps cover W <= 8 (larger windows are clipped as before)
#ifndef B200FE_WARP_LAUNCH_DEPENDENTS
#define B200FE_WARP_LAUNCH_DEPENDENTS 0
#endif
#ifndef B200FE_WARP_STAGE
#define B200FE_WARP_STAGE 1
#endif
constexpr bool kWarpStage = B200FE_WARP_STAGE != 0;
constexpr int kWarpWinExtra = 20;   // source rows staged per CTA beyond kWarpRows: 2 x (|center - warped| <= 5 in the reference's setting, + 5 for the support)

// Work split (fifth version).  History: v1 spent 218 thread-instructions per output cell in a coefficient set-up that 32 of 256
// threads ran alone; v2 (parallel set-up, 128-bit taps) 174 us on C3; its ncu capture (profiles/r03_ncu_time_warp.json) shows the
// XU pipe -- the float32 -> float64 converter, one conversion per tap and output cell, 64-bit results at half rate -- as the
// limiter ("xu_realtime" 209 %), which is why halving the set-up / statistics instructions and prefetching taps changed nothing.
// A second capture after those cuts (98 M instead of 125 M warp instructions, yet 184 us) settled it: the kernel waits for its tap
// loads (long-scoreboard stalls 7.3 per issue), and neither more loads in flight per thread (66-80 registers: fewer warps, slower)
// nor a float64 copy of the source rows in shared memory (twice the shared-memory bytes per tap, slower) pays.  v5: (0) thread 0
// requests the CTA's source rows -- [r0 - m, r0 + rows + m), contiguous in memory -- with ONE bulk copy (cp.async.bulk, mbarrier)
// before anything else; (1) one thread per output row derives its tap window (one float64 division per row); (2) one warp
// compacts the live (row, tap) pairs; (3) raw bicubic coefficients, one live pair per thread; (4) one thread per row adds them in
// Pillow's order (the sum's rounding depends on it); (5) one division per live pair; (6) wait for the copy (it has had the whole
// set-up to land); thread = (row slot, 4 columns): 128-bit taps from shared memory, float64 multiply and add per tap in source
// order, one rounding to float32, 128-bit store; (7) statistics from the rows just written (same CTA, after a barrier): float32
// partial sums per (row slot, column), added in float64 in a fixed order, one atomic per (moment, column).
__global__ void __launch_bounds__(256, B200FE_WARP_OCC) time_warp_kernel(const WarpArgs a)
{
#if B200FE_WARP_LAUNCH_DEPENDENTS
    pdl_launch_dependents();
#endif
    pdl_wait();                                                    // launched with programmatic serialisation behind the launch that writes `in`
    const int utt = blockIdx.y;
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int r0 = blockIdx.x * kWarpRows;
    if (r0 >= a.Tmax) return;
    __shared__ __align__(16) double s_k[kWarpRows][kWarpTaps];
    __shared__ double s_c[kWarpRows], s_ss[kWarpRows], s_ww[kWarpRows];
    __shared__ int s_ymin[kWarpRows], s_n[kWarpRows], s_xmin[kWarpRows], s_cls[kWarpRows], s_npairs;
    __shared__ unsigned short s_pair[kWarpRows * kWarpTaps];
    extern __shared__ __align__(16) float s_win[];               // [win_rows][nmel]: the CTA's source rows (one bulk copy); later the statistics partials
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int center = a.warp[2 * utt], warped = a.warp[2 * utt + 1];
    const float* src = a.in + (long long)utt * a.Tmax * a.nmel;
    float* dst = a.out + (long long)utt * a.Tmax * a.nmel;
    const int nb = a.n_cls - 1;
    const int* bounds = (a.stats != nullptr && a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
    const int nvalid = min(max(T - r0, 0), kWarpRows);            // rows of this CTA that hold frames
    const int nq = a.nmel >> 2;
    const bool vec = (a.nmel & 3) == 0 && nq <= 256 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    // Source window of the CTA, requested before anything else so that the copy flies during the coefficient set-up: an output
    // row's taps lie within |center - warped| + 1/2 rows of the row itself plus the filter support (2 x scale, scale ~ 1), so
    // [r0 - m, r0 + rows + m) with m = |center - warped| + 5 holds them; a row whose taps fall outside after all (checked per
    // row) reads global memory.
    int w0 = 0, wn = 0;
    if (vec && a.win_rows > 0 && nvalid > 0) {
        const int m = (center >= 0 ? abs(center - warped) : 0) + 5;
        w0 = max(r0 - m, 0);
        wn = min(min(r0 + nvalid + m, T) - w0, a.win_rows);
    }
    if (tid == 0 && wn > 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t bytes = (uint32_t)wn * a.nmel * sizeof(float);
        mbar_expect_tx(&s_bar, bytes);
        tma_load_1d(s_win, src + (long long)w0 * a.nmel, bytes, &s_bar);
    }
    if (tid < kWarpRows) {
        const int row = r0 + tid;
        int ymin = row, cnt = 1, xmin = 0;
        double c = 0.0, ss = 0.0;
        if (row < T && center >= 0) {
            const bool left = row < warped;
            const int in_size = left ? center : T - center;
            const int out_size = left ? warped : T - warped;
            const int xx = left ? row : row - warped;
            const int base = left ? 0 : center;
            if (in_size == out_size) {
                ymin = base + xx;                                      // Pillow returns a copy when the size is unchanged
            } else {
                const double scale = (double)in_size / (double)out_size;
                const double filterscale = scale < 1.0 ? 1.0 : scale;
                const double support = __dmul_rn(2.0, filterscale);
                c = __dmul_rn((double)xx + 0.5, scale);
                ss = 1.0 / filterscale;
                xmin = (int)__dadd_rn(__dsub_rn(c, support), 0.5);
                if (xmin < 0) xmin = 0;
                int xmax = (int)__dadd_rn(__dadd_rn(c, support), 0.5);
                if (xmax > in_size) xmax = in_size;
                cnt = -min(xmax - xmin, kWarpTaps);                   // negative: coefficients still to be evaluated
                ymin = base + xmin;
            }
        }
        s_ymin[tid] = ymin; s_n[tid] = cnt; s_xmin[tid] = xmin; s_c[tid] = c; s_ss[tid] = ss;
        s_cls[tid] = bounds ? row_class(bounds, nb, row) : 0;
        if (cnt > 0) s_k[tid][0] = 1.0;
    }
    __syncthreads();
    // one warp lists the (row, tap) pairs that need a coefficient, so that the float64 divisions below run on full warps (a
    // row has 5-6 of the 16 tap slots in use; mapping slots to threads directly left two thirds of every warp idle)
    if (tid < 32) {
        int base = 0;
        for (int c = 0; c < kWarpRows; c += 32) {
            const int r = c + tid;
            const int cr = r < kWarpRows ? max(-s_n[r], 0) : 0;
            int inc = cr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
            const int off = base + inc - cr;
            for (int x = 0; x < cr; ++x) s_pair[off + x] = (unsigned short)((r << 4) | x);
            base += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) s_npairs = base;
    }
    __syncthreads();
    const int np = s_npairs;
    for (int i = tid; i < np; i += 256) {
        const int r = s_pair[i] >> 4, x = s_pair[i] & 15;
        s_k[r][x] = pil_bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + s_xmin[r]), s_c[r]), 0.5), s_ss[r]));
    }
    __syncthreads();
    if (tid < kWarpRows && s_n[tid] < 0) {
        double ww = 0.0;
        for (int x = 0; x < -s_n[tid]; ++x) ww = __dadd_rn(ww, s_k[tid][x]);
        s_ww[tid] = ww;
    }
    __syncthreads();
    for (int i = tid; i < np; i += 256) {
        const int r = s_pair[i] >> 4, x = s_pair[i] & 15;
        if (s_ww[r] != 0.0) s_k[r][x] = s_k[r][x] / s_ww[r];
    }
    __syncthreads();
    if (wn > 0) mbar_wait(&s_bar, 0);                              // initialised before the first barrier above; the copy has had the whole set-up to land
    if (vec) {
        const int slots = 256 / nq, slot = tid / nq, q = tid - slot * nq;      // thread = (row slot, 4 columns): no division in the row loop
        if (slot < slots) {
            for (int r = slot; r < kWarpRows; r += slots) {
                const int row = r0 + r;
                if (row >= a.Tmax) break;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < T) {
                    const int cnt = abs(s_n[r]), ymin = s_ymin[r];
                    const double* kp = s_k[r];
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    if (ymin >= w0 && ymin + cnt <= w0 + wn) {
                        const float4* sp = reinterpret_cast<const float4*>(s_win + (ymin - w0) * a.nmel) + q;
                        for (int y = 0; y < cnt; ++y, sp += nq) {
                            const float4 p = *sp;
                            const double k = kp[y];
                            a0 = __dadd_rn(a0, __dmul_rn((double)p.x, k)); a1 = __dadd_rn(a1, __dmul_rn((double)p.y, k));
                            a2 = __dadd_rn(a2, __dmul_rn((double)p.z, k)); a3 = __dadd_rn(a3, __dmul_rn((double)p.w, k));
                        }
                    } else {
                        const float4* sp = reinterpret_cast<const float4*>(src + (long long)ymin * a.nmel) + q;
                        for (int y = 0; y < cnt; ++y, sp += nq) {
                            const float4 p = __ldg(sp);
                            const double k = kp[y];
                            a0 = __dadd_rn(a0, __dmul_rn((double)p.x, k)); a1 = __dadd_rn(a1, __dmul_rn((double)p.y, k));
                            a2 = __dadd_rn(a2, __dmul_rn((double)p.z, k)); a3 = __dadd_rn(a3, __dmul_rn((double)p.w, k));
                        }
                    }
                    v = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
                }
                reinterpret_cast<float4*>(dst + (long long)row * a.nmel)[q] = v;
            }
        }
    } else {
        for (int e = tid; e < kWarpRows * a.nmel; e += 256) {
            const int r = e / a.nmel, col = e - r * a.nmel, row = r0 + r;
            if (row >= a.Tmax) break;
            float v = 0.f;
            if (row < T) {
                const int ymin = s_ymin[r], cnt = abs(s_n[r]);
                double acc = 0.0;
                for (int y = 0; y < cnt; ++y) acc = __dadd_rn(acc, __dmul_rn((double)src[(long long)(ymin + y) * a.nmel + col], s_k[r][y]));
                v = (float)acc;
            }
            dst[(long long)row * a.nmel + col] = v;
        }
    }
    if (a.stats == nullptr) return;
    __syncthreads();                                               // the CTA's rows of `dst` are visible to all of its threads from here on
    double* sb = a.stats + (long long)utt * a.stats_stride;
    const float* orow = dst + (long long)r0 * a.nmel;
    if (nvalid > 0 && vec && a.part_ok && s_cls[0] == s_cls[nvalid - 1]) {
        // all rows of this CTA lie in one SpecAugment row class (row classes ascend with the row, so first == last says so; true
        // for all but a handful of CTAs per utterance): thread = (row slot, 4 columns) as in the main loop -- it reads back its own
        // stores -- float32 partial sums over its <= ceil(rows / slots) rows, partials through the source window (idle since
        // the barrier above), then one thread per (moment, column) adds the slots' partials in float64, in a fixed order.
        const int slots = 256 / nq, slot = tid / nq, q = tid - slot * nq;
        const int used = min(slots, nvalid);                          // slots that own a row
        float* part1 = s_win;                                         // [used][nmel]
        float* part2 = part1 + min(slots, kWarpRows) * a.nmel;        // [used][nmel]; the host sizes the buffer for 2 * min(slots, kWarpRows) * nmel floats
        if (slot < used) {
            float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
            for (int r = slot; r < nvalid; r += slots) {
                const float4 v = reinterpret_cast<const float4*>(orow + (long long)r * a.nmel)[q];
                s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
                s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
            }
            *reinterpret_cast<float4*>(part1 + slot * a.nmel + 4 * q) = s1;
            *reinterpret_cast<float4*>(part2 + slot * a.nmel + 4 * q) = s2;
        }
        __syncthreads();
        const int cls = s_cls[0];
        for (int i = tid; i < 2 * a.nmel; i += 256) {
            const int m = i >= a.nmel, j = i - m * a.nmel;
            const float* pp = (m ? part2 : part1) + j;
            double acc = 0.0;
            for (int k = 0; k < used; ++k) acc += (double)pp[k * a.nmel];
            atomicAdd(sb + (long long)(m ? a.n_cls : cls) * a.nmel + j, acc);
        }
    } else if (nvalid > 0) {
        for (int j = tid; j < a.nmel; j += 256) {                  // a class boundary inside the CTA's rows, or an odd layout: one thread per column
            int cls = s_cls[0];
            double s1 = 0.0, s2 = 0.0;
            for (int fr = 0; fr < nvalid; ++fr) {
                const int cc = s_cls[fr];
                if (cc != cls) { atomicAdd(sb + (long long)cls * a.nmel + j, s1); s1 = 0.0; cls = cc; }
                const double x = (double)orow[(long long)fr * a.nmel + j];
                s1 += x;
                s2 = fma(x, x, s2);
            }
            atomicAdd(sb + (long long)cls * a.nmel + j, s1);
            atomicAdd(sb + (long long)a.n_cls * a.nmel + j, s2);
        }
    }
    if (a.done == nullptr) return;

    // ---- SpecAugment masks in the same launch: the CTA whose arrival completes the utterance (its rows and statistics are all
    // in L2 by then) derives the fills and overwrites the masked cells; later masks win, time masks come after frequency masks
    // (specaugment.py applies them in that order).  Replaces a finalize launch and a mask-fill pass over the whole batch.
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
        // this CTA's rows and statistics before its arrival (cumulative over the barrier); fence.acq_rel, not __threadfence():
        // the latter is a sequentially consistent fence plus an L1 invalidate, several times the cost on every CTA's tail
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        const int prev = atomicAdd(a.done + utt, 1);
        s_last = prev == (int)gridDim.x - 1;
        if (s_last) { a.done[utt] = 0; asm volatile("fence.acq_rel.gpu;" ::: "memory"); }   // every CTA of the utterance has arrived: the counter is free again
    }
    __syncthreads();
    if (!s_last) return;
    warp_mask_tail(a, utt, T, dst, vec);
}

// Vectorised in-place post pass for num_mel_bins % 4 == 0: thread = (row slot, 4 columns), float4
// traffic, per-thread CMVN vectors and frequency-mask winners in registers.
__global__ void __launch_bounds__(256) postpass_vec_kernel(const PostArgs a)
{
    // CTAs are scheduled x-fastest, y ascending: walking utterances and row blocks BACKWARDS starts with the rows the fused launch
    // wrote last, i.e. the ones most likely to be L2-resident still
    pdl_launch_dependents();
    pdl_wait();
    const int utt = (int)(gridDim.y - 1 - blockIdx.y);
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int r0 = (int)(gridDim.x - 1 - blockIdx.x) * a.rows_per_cta;
    if (r0 >= T) return;
    __shared__ __align__(16) float s_mu[1024], s_is[1024];
    if (a.inline_finalize) {
        // same fp64 -> fp32 vectors as finalize_kernel (one row class); the CTA of the utterance's first rows publishes them
        const double* sb = a.stats + (long long)utt * a.stats_stride;
        for (int d = threadIdx.x; d < a.nmel; d += 256) {
            const double mean = sb[d] / T;
            double istd = 1.0;
            if (a.cmvn_mode == 2) {
                const double var = sb[(long long)a.n_cls * a.nmel + d] / T - mean * mean;
                istd = 1.0 / sqrt(var > 1e-20 ? var : 1e-20);
            }
            s_mu[d] = (float)mean; s_is[d] = (float)istd;
            if (r0 == 0) { a.cm_mean[(long long)utt * a.nmel + d] = (float)mean; a.cm_istd[(long long)utt * a.nmel + d] = (float)istd; }
        }
        __syncthreads();
    }
    const int r1 = min(r0 + a.rows_per_cta, T);
    const int nq = a.nmel >> 2;                    // float4 groups per row
    const int slots = 256 / nq;                    // rows processed concurrently
    const int q = threadIdx.x % nq, slot = threadIdx.x / nq;
    if (slot >= slots) return;
    const int nm = a.n_fmask + a.n_tmask;
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {1.f, 1.f, 1.f, 1.f};
    int fhit[4] = {-1, -1, -1, -1};
    float ffill[4] = {0.f, 0.f, 0.f, 0.f};
    int tlo[kMaxTimeMasks], thi[kMaxTimeMasks];
    float tfill[kMaxTimeMasks];
    if (a.cmvn_mode != 0) {
        const float4 m4 = a.inline_finalize ? *reinterpret_cast<const float4*>(s_mu + 4 * q) : *reinterpret_cast<const float4*>(a.cm_mean + (long long)utt * a.nmel + 4 * q);
        const float4 s4 = a.inline_finalize ? *reinterpret_cast<const float4*>(s_is + 4 * q) : *reinterpret_cast<const float4*>(a.cm_istd + (long long)utt * a.nmel + 4 * q);
        mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
        is[0] = s4.x; is[1] = s4.y; is[2] = s4.z; is[3] = s4.w;
    }
#pragma unroll
    for (int i = 0; i < kMaxTimeMasks; ++i) { tlo[i] = 0; thi[i] = 0; tfill[i] = 0.f; }
    if (a.masks) {
        const int* mk = a.masks + (long long)utt * nm * 2;
        const float* fl = a.fills + (long long)utt * nm;
        for (int i = 0; i < a.n_fmask; ++i) {
            const int lo = mk[2 * i], hi = mk[2 * i + 1];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * q + c >= lo && 4 * q + c < hi) { fhit[c] = i; ffill[c] = fl[i]; }
        }
#pragma unroll
        for (int i = 0; i < kMaxTimeMasks; ++i)
            if (i < a.n_tmask) { tlo[i] = mk[2 * (a.n_fmask + i)]; thi[i] = mk[2 * (a.n_fmask + i) + 1]; tfill[i] = fl[a.n_fmask + i]; }
    }
    const bool need_read = a.cmvn_mode != 0;
    float4* base = reinterpret_cast<float4*>(a.feats + (a.feat_offsets != nullptr ? a.feat_offsets[utt] : (long long)utt * a.Tmax) * a.nmel);
    if (need_read && a.masks == nullptr) {
        // plain utterance CMVN: four rows' loads are in flight before the first store (the in-place loop below gives the compiler
        // no licence to hoist a load above the previous row's store)
        for (int rb = r0 + slot; rb < r1; rb += 4 * slots) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int r = rb + u * slots; if (r < r1) v[u] = base[(long long)r * nq + q]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u * slots;
                if (r < r1) base[(long long)r * nq + q] = make_float4((v[u].x - mu[0]) * is[0], (v[u].y - mu[1]) * is[1], (v[u].z - mu[2]) * is[2], (v[u].w - mu[3]) * is[3]);
            }
        }
        return;
    }
    for (int r = r0 + slot; r < r1; r += slots) {
        int thit = -1; float tf = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxTimeMasks; ++i) if (r >= tlo[i] && r < thi[i]) { thit = i; tf = tfill[i]; }   // later time masks win
        float4* ptr = base + (long long)r * nq + q;
        const bool anyf = (fhit[0] | fhit[1] | fhit[2] | fhit[3]) >= 0 || fhit[0] >= 0 || fhit[1] >= 0 || fhit[2] >= 0 || fhit[3] >= 0;
        if (!need_read && thit < 0 && !anyf) continue;                    // nothing changes in this float4
        if (!need_read) {
            // masks only (global / no CMVN): nothing is read -- a float4 that is masked entirely is one 16-byte store, a partly
            // masked one gets scalar stores (the read-modify-write of the earlier version made this a latency-bound pass)
            const bool allf = fhit[0] >= 0 && fhit[1] >= 0 && fhit[2] >= 0 && fhit[3] >= 0;
            if (thit >= 0) *ptr = make_float4(tf, tf, tf, tf);
            else if (allf) *ptr = make_float4(ffill[0], ffill[1], ffill[2], ffill[3]);
            else {
                float* sp = reinterpret_cast<float*>(ptr);
#pragma unroll
                for (int c = 0; c < 4; ++c) if (fhit[c] >= 0) sp[c] = ffill[c];
            }
            continue;
        }
        float4 v = *ptr;
        float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (need_read) x[c] = (x[c] - mu[c]) * is[c];
            if (thit >= 0) x[c] = tf;                                    // time masks are applied after frequency masks
            else if (fhit[c] >= 0) x[c] = ffill[c];
        }
        *ptr = make_float4(x[0], x[1], x[2], x[3]);
    }
}

// Plain utterance CMVN (no SpecAugment fills): the post pass of BASELINE config 2.  Same arithmetic as postpass_vec_kernel's
// CMVN path, as a kernel of its own so that its register count (and with it the number of resident CTAs) is not set by the
// mask logic: the ncu capture of the shared kernel (profiles/r03_ncu_postpass.json) shows a LATENCY-bound pass -- 52 registers,
// 4 CTAs per SM, long-scoreboard stalls 12 per issue, DRAM 52 % -- not a bandwidth-bound one.
#ifndef B200FE_POST_OCC
#define B200FE_POST_OCC 5          // measured on B200 (C2 step): 4 -> +61.5 us, 5 -> +58.0, 6 (32 B of spills) -> +64.2, 8 (spills) -> +92
#endif
__global__ void __launch_bounds__(256, B200FE_POST_OCC) postpass_cmvn_kernel(const PostArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    const int utt = (int)(gridDim.y - 1 - blockIdx.y);          // backwards: the rows the fused launch wrote last are the most likely L2 hits
    const long long n = a.nsamp[utt];
    const int T = n >= a.win ? (int)(1 + (n - a.win) / a.shift) : 0;
    const int r0 = (int)(gridDim.x - 1 - blockIdx.x) * a.rows_per_cta;
    if (r0 >= T) return;
    __shared__ __align__(16) float s_mu[kMaxMel], s_is[kMaxMel];
    {
        const double* sb = a.stats + (long long)utt * a.stats_stride;
        for (int d = threadIdx.x; d < a.nmel; d += 256) {
            const double mean = sb[d] / T;
            double istd = 1.0;
            if (a.cmvn_mode == 2) {
                const double var = sb[(long long)a.n_cls * a.nmel + d] / T - mean * mean;
                istd = 1.0 / sqrt(var > 1e-20 ? var : 1e-20);
            }
            s_mu[d] = (float)mean; s_is[d] = (float)istd;
            if (r0 == 0) { a.cm_mean[(long long)utt * a.nmel + d] = (float)mean; a.cm_istd[(long long)utt * a.nmel + d] = (float)istd; }
        }
    }
    __syncthreads();
    const int r1 = min(r0 + a.rows_per_cta, T);
    const int nq = a.nmel >> 2, slots = 256 / nq;
    const int q = threadIdx.x % nq, slot = threadIdx.x / nq;
    if (slot >= slots) return;
    const float4 mu = *reinterpret_cast<const float4*>(s_mu + 4 * q), is = *reinterpret_cast<const float4*>(s_is + 4 * q);
    float4* base = reinterpret_cast<float4*>(a.feats + (a.feat_offsets != nullptr ? a.feat_offsets[utt] : (long long)utt * a.Tmax) * a.nmel);
    for (int rb = r0 + slot; rb < r1; rb += 4 * slots) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int r = rb + u * slots; if (r < r1) v[u] = base[(long long)r * nq + q]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = rb + u * slots;
            if (r < r1) base[(long long)r * nq + q] = make_float4((v[u].x - mu.x) * is.x, (v[u].y - mu.y) * is.y, (v[u].z - mu.z) * is.z, (v[u].w - mu.w) * is.w);
        }
    }
}

}  // namespace b200fe

// Streaming state of the C ABI (b200fe_stream_*, include/b200fe.h): S INDEPENDENT audio streams per handle, each with its
// own carry of the samples the next chunk's first frames still need (window - shift ... window - 1 of them; TA:63-67 framing
// continued across chunk boundaries).  The reference has no streaming fbank -- ASRProcess.frontend
// (R/lasr/process/asrprocess.py:49-56) transforms whole utterances -- so correctness is "concatenated chunk outputs == offline
// fbank of the whole stream" (SURVEY.md 8(d) C5).  A push is three launches on the caller's stream, none of which depends on
// host-side lengths (capturable in a CUDA graph): append -> b200fe_fbank_fused on the state rows -> advance.
#pragma once
#include "b200fe_common.cuh"

namespace b200fe {

// One CTA per pushed stream: the chunk is appended behind the stream's carry; the fused launch then reads
// state[id][0 : nsamp) through (offsets, nsamp).  flags[0] |= 1 if a chunk had to be clipped, |= 2 if a push completed more
// frames than the output can hold (the excess frames are dropped from the output but still consumed).
__global__ void __launch_bounds__(256) stream_append_kernel(float* __restrict__ state, int* __restrict__ fill, int cap, int n_streams,
                                                            const int* __restrict__ ids, const float* __restrict__ chunks, long long chunk_stride,
                                                            const int* __restrict__ chunk_len, int max_chunk, int win, int shift, int max_out_frames,
                                                            long long* __restrict__ offsets, long long* __restrict__ nsamp, long long* __restrict__ out_frames,
                                                            int* __restrict__ flags)
{
    const int i = blockIdx.x;
    const int id = ids ? ids[i] : i;
    if (id < 0 || id >= n_streams) {            // not a stream of this handle: an empty row (no frames)
        if (threadIdx.x == 0) { offsets[i] = 0; nsamp[i] = 0; if (out_frames) out_frames[i] = 0; atomicOr(flags, 4); }
        return;
    }
    int len = chunk_len ? chunk_len[i] : max_chunk;
    if (len < 0) len = 0;
    const int f = fill[id];
    if (len > max_chunk || f + len > cap) { len = min(max_chunk, cap - f); if (threadIdx.x == 0) atomicOr(flags, 1); }
    float* dst = state + (long long)id * cap + f;
    const float* src = chunks + (long long)i * chunk_stride;
    for (int k = threadIdx.x; k < len; k += blockDim.x) dst[k] = src[k];
    if (threadIdx.x == 0) {
        const int n = f + len;
        int T = n >= win ? 1 + (n - win) / shift : 0;
        if (T > max_out_frames) atomicOr(flags, 2);
        offsets[i] = (long long)id * cap;
        nsamp[i] = n;
        if (out_frames) out_frames[i] = min(T, max_out_frames);
    }
}

// After the fused launch: the samples the next push still needs move to the front of the stream's row.
__global__ void __launch_bounds__(128) stream_advance_kernel(float* __restrict__ state, int* __restrict__ fill, int cap, int n_streams,
                                                             const int* __restrict__ ids, const long long* __restrict__ nsamp, int win, int shift)
{
    const int i = blockIdx.x;
    const int id = ids ? ids[i] : i;
    if (id < 0 || id >= n_streams) return;
    const int n = (int)nsamp[i];
    const int T = n >= win ? 1 + (n - win) / shift : 0;
    const int used = T * shift, rest = n - used;                 // rest <= max(win - 1, carry) < 1024
    float* row = state + (long long)id * cap;
    float keep[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { const int k = threadIdx.x + r * 128; keep[r] = (used > 0 && k < rest) ? row[used + k] : 0.f; }
    __syncthreads();
    if (used > 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) { const int k = threadIdx.x + r * 128; if (k < rest) row[k] = keep[r]; }
    }
    if (threadIdx.x == 0) fill[id] = rest;
}

__global__ void stream_reset_kernel(int* __restrict__ fill, int n_streams, const int* __restrict__ ids, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids ? ids[i] : i;
    if (id >= 0 && id < n_streams) fill[id] = 0;
}

}  // namespace b200fe

// Tensor-core DFT probe: is a 256-point complex FFT built from two 16x16 complex DFT stages on the tensor
// cores (mma.sync.m16n8k8 TF32, 3xTF32 error compensation so that the result keeps ~fp32 accuracy) faster than the CUDA-core
// FFT of fbank_fused_kernel (~184 SM-cycles per frame measured, ~80 at the FMA-pipe floor)?
//
//   part A: raw issue rate of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 on this GPU (MMAs per clock per SM).
//   part B: a working FFT-256: one warp per frame, stage = [F_r -F_i; F_i F_r](32x32, constant A fragments, hi + lo parts in
//           registers) x [D_r; D_i](32x16, B fragments from shared memory, split hi/lo on the fly): 16 MMAs per stage at 1xTF32,
//           48 at 3xTF32 (hi*hi + lo*hi + hi*lo); twiddles W_256^(n1 k2) on the accumulator registers; one shared-memory
//           transposition between the stages -- the same data movement as the CUDA-core kernel's.
//
//   tc_probe <frames.bin> <out.bin> <nframes> [splits=3]      frames: nframes x 256 complex float32 (windowed, packed z[n])
// Prints one JSON line: MMA rate, ns and SM-cycles per frame of the FFT alone (no load/window/power/mel/log around it).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- part A: issue rate ----
__global__ void __launch_bounds__(256) mma_rate_kernel(float* out, int iters)
{
    float c[8][4];
    unsigned a[4] = {tf32(1.0f + threadIdx.x), tf32(0.5f), tf32(0.25f), tf32(2.0f)};
    unsigned b0 = tf32(1.0f), b1 = tf32(-1.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32(c[i], a, b0, b1);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

// ---- part B: FFT-256 = two DFT-16 stages as TF32 matrix products ----
constexpr int kStride = 24;            // words per row of the 32 x 16 B-operand source in shared memory (conflict-free fragment loads)
constexpr int kWarpsPerCta = 8;

struct Consts {
    unsigned ahi[2][4][4], alo[2][4][4];   // [m tile][k step][fragment register] of [F_r -F_i; F_i F_r]
};

template <int kSplits>
__device__ __forceinline__ void stage(const Consts& K, const float* __restrict__ S, float (&acc)[2][2][4], int g, int t)
{
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float x0 = S[(8 * ks + t) * kStride + 8 * nt + g], x1 = S[(8 * ks + t + 4) * kStride + 8 * nt + g];
            const unsigned h0 = tf32(x0), h1 = tf32(x1);
            const unsigned l0 = tf32(x0 - __uint_as_float(h0)), l1 = tf32(x1 - __uint_as_float(h1));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_tf32(acc[mt][nt], K.ahi[mt][ks], h0, h1);
                if (kSplits >= 3) {
                    mma_tf32(acc[mt][nt], K.alo[mt][ks], h0, h1);
                    mma_tf32(acc[mt][nt], K.ahi[mt][ks], l0, l1);
                }
            }
        }
    }
}

template <int kSplits>
__global__ void __launch_bounds__(32 * kWarpsPerCta) fft256_tc_kernel(const float2* __restrict__ in, float2* __restrict__ out, int nframes, int reps)
{
    __shared__ float smem[kWarpsPerCta][32 * kStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    float* S = smem[warp];
    // constant A fragments: Fbig[m][k], m = (re | im) x k2, k = (re | im) x n2;  F[k2][n2] = exp(-2 pi j k2 n2 / 16)
    Consts K;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = 16 * mt + g + 8 * (r & 1), k = 8 * ks + t + 4 * (r >> 1);
                const int k2 = m & 15, n2 = k & 15;
                const double ang = -2.0 * M_PI * ((k2 * n2) & 15) / 16.0;
                const double fr = cos(ang), fi = sin(ang);
                const double v = (m < 16) ? ((k < 16) ? fr : -fi) : ((k < 16) ? fi : fr);
                const float vf = (float)v;
                K.ahi[mt][ks][r] = tf32(vf);
                K.alo[mt][ks][r] = tf32(vf - __uint_as_float(K.ahi[mt][ks][r]));
            }
    // twiddles W_256^(n1 k2) for this lane's accumulator positions: k2 = g + 8 (c >> 1), n1 = 8 nt + 2 t + (c & 1)
    float2 tw[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k2 = g + 8 * (c >> 1), n1 = 8 * nt + 2 * t + (c & 1);
            const double ang = -2.0 * M_PI * ((k2 * n1) & 255) / 256.0;
            tw[nt][c] = make_float2((float)cos(ang), (float)sin(ang));
        }
    const int gw = blockIdx.x * kWarpsPerCta + warp, nw = gridDim.x * kWarpsPerCta;
    for (int rep = 0; rep < reps; ++rep)
        for (int f = gw; f < nframes; f += nw) {
            // z[n1 + 16 n2] -> S[n2][n1] (re rows 0..15, im rows 16..31)
            const float2* zf = in + (size_t)f * 256;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = lane + 32 * j, n1 = i & 15, n2 = i >> 4;
                const float2 z = zf[i];
                S[n2 * kStride + n1] = z.x;
                S[(16 + n2) * kStride + n1] = z.y;
            }
            __syncwarp();
            float acc[2][2][4];
            stage<kSplits>(K, S, acc, g, t);           // acc[0] = Re A[k2][n1], acc[1] = Im A[k2][n1]
            __syncwarp();
            // twiddle, then B[n1][k2] -> S[n1][k2] for the second stage (sum over n1)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float re = acc[0][nt][c], im = acc[1][nt][c];
                    const float2 w = tw[nt][c];
                    const int k2 = g + 8 * (c >> 1), n1 = 8 * nt + 2 * t + (c & 1);
                    S[n1 * kStride + k2] = fmaf(re, w.x, -im * w.y);
                    S[(16 + n1) * kStride + k2] = fmaf(re, w.y, im * w.x);
                }
            __syncwarp();
            stage<kSplits>(K, S, acc, g, t);           // acc[0] = Re X[k1][k2], acc[1] = Im;  bin = k2 + 16 k1
            __syncwarp();
            float2* of = out + (size_t)f * 256;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k1 = g + 8 * (c >> 1), k2 = 8 * nt + 2 * t + (c & 1);
                    of[k2 + 16 * k1] = make_float2(acc[0][nt][c], acc[1][nt][c]);
                }
        }
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: tc_probe frames.bin out.bin nframes [splits]\n"); return 2; }
    const int nframes = atoi(argv[3]), splits = argc > 4 ? atoi(argv[4]) : 3;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float* dsink; CK(cudaMalloc(&dsink, 16));
    // ---- part A ----
    double best_rate = 0; int best_ctas = 0;
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        const int iters = 20000;
        mma_rate_kernel<<<sms * ctas, 256>>>(dsink, 100);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        mma_rate_kernel<<<sms * ctas, 256>>>(dsink, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double mmas = (double)sms * ctas * 8 /*warps*/ * iters * 8;
        const double rate = mmas / (ms * 1e-3) / ((double)khz * 1e3) / sms;     // MMAs per clock per SM at the nominal max clock
        if (rate > best_rate) { best_rate = rate; best_ctas = ctas; }
    }
    // ---- part B ----
    std::vector<float> h((size_t)nframes * 512);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(h.data(), sizeof(float), h.size(), fi) != h.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(fi);
    float2 *din, *dout;
    CK(cudaMalloc(&din, h.size() * 4)); CK(cudaMalloc(&dout, h.size() * 4));
    CK(cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    auto kern = splits >= 3 ? fft256_tc_kernel<3> : fft256_tc_kernel<1>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * kWarpsPerCta, 0));
    const int grid = sms * (occ > 0 ? occ : 1);
    kern<<<grid, 32 * kWarpsPerCta>>>(din, dout, nframes, 1);
    CK(cudaDeviceSynchronize());
    const int reps = 50;
    CK(cudaEventRecord(e0));
    kern<<<grid, 32 * kWarpsPerCta>>>(din, dout, nframes, reps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
    FILE* fo = fopen(argv[2], "wb");
    fwrite(h.data(), sizeof(float), h.size(), fo);
    fclose(fo);
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    const double ns_per_frame = ms * 1e6 / ((double)nframes * reps);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"mma_m16n8k8_tf32_per_clk_per_sm\": %.3f, \"mma_rate_ctas_per_sm\": %d, "
           "\"fft_splits\": %d, \"fft_regs\": %d, \"fft_ctas_per_sm\": %d, \"fft_ns_per_frame\": %.4f, \"fft_sm_cycles_per_frame\": %.1f, "
           "\"mmas_per_frame\": %d, \"frames\": %d}\n",
           prop.name, sms, khz / 1e3, best_rate, best_ctas, splits, fa.numRegs, occ, ns_per_frame, ns_per_frame * 1e-9 * khz * 1e3 * sms,
           splits >= 3 ? 96 : 32, nframes);
    return 0;
}

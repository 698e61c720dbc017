// Pipe-throughput microbenchmark for sm_100a (B200): scalar vs packed (f32x2) FP32,
// MUFU, shared-memory loads and shuffles.  Informs the fbank kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd1(float a, float b){ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float lg2a(float a){ float d; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

enum { OP_FFMA=0, OP_FFMA2, OP_FADD, OP_FADD2, OP_FMUL2, OP_MIX_FFMA_FADD, OP_MIX2, OP_LG2, OP_LDS32, OP_LDS64, OP_LDS128, OP_SHFL, OP_FFMA_LDS64, OP_FFMA2_LDS128, OP_N };
static const char* names[] = {"FFMA","FFMA2","FADD","FADD2","FMUL2","FFMA+FADD","FFMA2+FADD2","MUFU.LG2","LDS.32","LDS.64","LDS.128","SHFL.BFLY","4xFFMA+LDS.64","4xFFMA2+LDS.128"};

template<int OP>
__global__ void __launch_bounds__(1024,1) bench(float* out, int iters, long long* cyc)
{
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0f + i*1e-6f;
    __syncthreads();
    float a[8]; u64 p[8];
    float b = 1.0000001f, c = 1e-9f;
    u64 pb, pc; { float2 t = make_float2(b, b); pb = *reinterpret_cast<u64*>(&t); t = make_float2(c,c); pc = *reinterpret_cast<u64*>(&t); }
    #pragma unroll
    for (int i=0;i<8;i++){ a[i] = 1.0f + threadIdx.x*1e-7f + i; float2 t = make_float2(a[i], a[i]+1); p[i] = *reinterpret_cast<u64*>(&t); }
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    long long t0 = clock64();
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int r=0;r<4;r++) {
        #pragma unroll
        for (int i=0;i<8;i++) {
            if (OP==OP_FFMA) a[i] = ffma1(a[i], b, c);
            else if (OP==OP_FFMA2) p[i] = ffma2(p[i], pb, pc);
            else if (OP==OP_FADD) a[i] = fadd1(a[i], c);
            else if (OP==OP_FADD2) p[i] = fadd2(p[i], pc);
            else if (OP==OP_FMUL2) p[i] = fmul2(p[i], pb);
            else if (OP==OP_MIX_FFMA_FADD) { if (i&1) a[i] = ffma1(a[i], b, c); else a[i] = fadd1(a[i], c); }
            else if (OP==OP_MIX2) { if (i&1) p[i] = ffma2(p[i], pb, pc); else p[i] = fadd2(p[i], pc); }
            else if (OP==OP_LG2) a[i] = lg2a(a[i]);
            else if (OP==OP_LDS32) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr + ((i*128+r*1024) & 8191))); a[i] += v; }
            else if (OP==OP_LDS64) { float v,w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v),"=f"(w) : "r"(saddr + ((i*512+r*4096) & 8191))); a[i] += v+w; }
            else if (OP==OP_LDS128) { float v,w,x,y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v),"=f"(w),"=f"(x),"=f"(y) : "r"(saddr + ((i*512+r*4096) & 8191))); a[i] += v+w+x+y; }
            else if (OP==OP_SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i&3));
            else if (OP==OP_FFMA_LDS64) {
                a[i] = ffma1(a[i], b, c); a[(i+1)&7] = ffma1(a[(i+1)&7], b, c); a[(i+2)&7] = ffma1(a[(i+2)&7], b, c); a[(i+3)&7] = ffma1(a[(i+3)&7], b, c);
                float v,w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v),"=f"(w) : "r"(saddr + ((i*512+r*4096) & 8191))); c += v*1e-30f + w*1e-30f; }
            else if (OP==OP_FFMA2_LDS128) {
                p[i] = ffma2(p[i], pb, pc); p[(i+1)&7] = ffma2(p[(i+1)&7], pb, pc); p[(i+2)&7] = ffma2(p[(i+2)&7], pb, pc); p[(i+3)&7] = ffma2(p[(i+3)&7], pb, pc);
                float v,w,x,y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v),"=f"(w),"=f"(x),"=f"(y) : "r"(saddr + ((i*512+r*4096) & 8191))); c += (v+w+x+y)*1e-30f; }
        }}
    }
    long long t1 = clock64();
    float s = c;
    #pragma unroll
    for (int i=0;i<8;i++){ float2 t = *reinterpret_cast<float2*>(&p[i]); s += a[i] + t.x + t.y; }
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int OP> void run(float* out, long long* cyc, int nsm)
{
    const int iters = 2000;
    bench<OP><<<nsm, 1024>>>(out, 10, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nsm, 1024>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, nsm*sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i=0;i<nsm;i++) avg += h[i]; avg /= nsm;
    double winstr = 32.0 * iters * 32;            // warp-instrs per warp (primary op count)
    if (OP==OP_FFMA_LDS64 || OP==OP_FFMA2_LDS128) winstr *= 5;  // 4 math + 1 LDS
    double per_sm = winstr * 32 /*warps*/ ;
    printf("%-18s cycles=%.0f  warp-instr/clk/SM=%.3f  (per SMSP %.3f)  ms=%.3f  clk=%.0f MHz\n", names[OP], avg, per_sm/avg, per_sm/avg/4, ms, avg/ms/1e3);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1);}
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int nsm = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, nsm);
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)nsm*1024*sizeof(float)); cudaMalloc(&cyc, 1024*sizeof(long long));
    run<OP_FFMA>(out,cyc,nsm); run<OP_FFMA2>(out,cyc,nsm); run<OP_FADD>(out,cyc,nsm); run<OP_FADD2>(out,cyc,nsm); run<OP_FMUL2>(out,cyc,nsm);
    run<OP_MIX_FFMA_FADD>(out,cyc,nsm); run<OP_MIX2>(out,cyc,nsm); run<OP_LG2>(out,cyc,nsm);
    run<OP_LDS32>(out,cyc,nsm); run<OP_LDS64>(out,cyc,nsm); run<OP_LDS128>(out,cyc,nsm); run<OP_SHFL>(out,cyc,nsm);
    run<OP_FFMA_LDS64>(out,cyc,nsm); run<OP_FFMA2_LDS128>(out,cyc,nsm);
    return 0;
}

// Microbenchmark behind csrc/host_simd.cpp: float64 -> float32 packing of ~584 MB by N threads in 256 kB tasks (the host staging
// pool's task size), SSE2 / AVX-512F non-temporal stores with and without software prefetch; best of 12 interleaved rounds.
//   g++ -O3 -std=c++17 -pthread -o tools/pack_bench tools/pack_bench.cpp && tools/pack_bench 8
// Results of the build container (Xeon model 207, 8 vCPUs): profiles/r03_host_simd.txt.
#include <immintrin.h>
#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cmath>
#include <algorithm>
typedef void (*cvt_fn)(const double*, float*, long long);
static void cvt_sse(const double* s, float* d, long long n) {
    long long i = 0;
    for (; i + 8 <= n; i += 8) {
        __m128 a = _mm_cvtpd_ps(_mm_loadu_pd(s + i)), b = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
        __m128 c = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 4)), e = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 6));
        _mm_stream_ps(d + i, _mm_movelh_ps(a, b)); _mm_stream_ps(d + i + 4, _mm_movelh_ps(c, e));
    }
    _mm_sfence();
    for (; i < n; ++i) d[i] = (float)s[i];
}
__attribute__((target("avx512f"))) static void cvt_512(const double* s, float* d, long long n) {
    long long i = 0;
    while (i < n && ((uintptr_t)(d + i) & 63)) { d[i] = (float)s[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i)), b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
        __m512d v = _mm512_insertf64x4(_mm512_castpd256_pd512(_mm256_castps_pd(a)), _mm256_castps_pd(b), 1) ;
        _mm512_stream_ps(d + i, _mm512_castpd_ps(v));
    }
    _mm_sfence();
    for (; i < n; ++i) d[i] = (float)s[i];
}
template <int PF, int HINT>
__attribute__((target("avx512f"))) static void cvt_512_pf(const double* s, float* d, long long n) {
    long long i = 0;
    while (i < n && ((uintptr_t)(d + i) & 63)) { d[i] = (float)s[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        _mm_prefetch((const char*)(s + i) + PF, (_mm_hint)HINT); _mm_prefetch((const char*)(s + i) + PF + 64, (_mm_hint)HINT);
        __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i)), b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
        __m512d v = _mm512_insertf64x4(_mm512_castpd256_pd512(_mm256_castps_pd(a)), _mm256_castps_pd(b), 1) ;
        _mm512_stream_ps(d + i, _mm512_castpd_ps(v));
    }
    _mm_sfence();
    for (; i < n; ++i) d[i] = (float)s[i];
}
__attribute__((target("avx2"))) static void cvt_avx2(const double* s, float* d, long long n) {
    long long i = 0;
    while (i < n && ((uintptr_t)(d + i) & 31)) { d[i] = (float)s[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i)), b = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 4));
        _mm256_stream_ps(d + i, _mm256_set_m128(b, a));
    }
    _mm_sfence();
    for (; i < n; ++i) d[i] = (float)s[i];
}
static void cvt_sse_cached(const double* s, float* d, long long n) {   // regular stores
    long long i = 0;
    for (; i + 4 <= n; i += 4) {
        __m128 a = _mm_cvtpd_ps(_mm_loadu_pd(s + i)), b = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
        _mm_storeu_ps(d + i, _mm_movelh_ps(a, b));
    }
    for (; i < n; ++i) d[i] = (float)s[i];
}
int main(int argc, char** argv) {
    int nt = argc > 1 ? atoi(argv[1]) : 8;
    const long long N = 73000000;   // ~584 MB of doubles
    double* src = (double*)aligned_alloc(64, N * 8);
    float* dst = (float*)aligned_alloc(64, N * 4);
    for (long long i = 0; i < N; ++i) src[i] = (double)(i & 1023) * 1e-3;
    memset(dst, 0, N * 4);
    struct { const char* name; cvt_fn f; } fns[] = {{"sse2_nt", cvt_sse}, {"avx512_nt", cvt_512}, {"avx512_nt_pf2k_t1", cvt_512_pf<2048, _MM_HINT_T1>}, {"avx512_nt_pf4k_t1", cvt_512_pf<4096, _MM_HINT_T1>},
        {"avx512_nt_pf8k_t1", cvt_512_pf<8192, _MM_HINT_T1>}, {"avx512_nt_pf16k_t1", cvt_512_pf<16384, _MM_HINT_T1>}, {"avx512_nt_pf4k_t2", cvt_512_pf<4096, _MM_HINT_T2>}, {"avx512_nt_pf4k_t0", cvt_512_pf<4096, _MM_HINT_T0>}, {"avx512_nt_pf8k_t0", cvt_512_pf<8192, _MM_HINT_T0>}};
    const long long chunk = 65536;    // elements per task (256 kB of destination)
    const int NF = sizeof(fns) / sizeof(fns[0]);
    double best[16]; for (int i = 0; i < 16; ++i) best[i] = 1e9;
    for (int rep = 0; rep < 12; ++rep) {
        for (int k = 0; k < NF; ++k) {
            auto& fn = fns[k];
            std::atomic<long long> next{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back([&] {
                for (;;) { long long c = next.fetch_add(chunk); if (c >= N) break; long long m = std::min(chunk, N - c); fn.f(src + c, dst + c, m); }
            });
            for (auto& t : th) t.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best[k]) best[k] = dt;
        }
    }
    for (int k = 0; k < NF; ++k) printf("%-22s threads %2d  %.2f ms  %.1f GB/s (read+write)\n", fns[k].name, nt, best[k] * 1e3, N * 12.0 / best[k] / 1e9);
    // check
    double err = 0; for (long long i = 0; i < N; i += 9973) err += fabs((double)dst[i] - (double)(float)src[i]);
    printf("check %g\n", err);
}

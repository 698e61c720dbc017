// Streaming primitives of the host staging pool (host_pool.h): float64 -> float32 conversion, copy and zero fill into (pinned)
// buffers that a DMA engine, not a core, reads next.  Plain C++ (compiled by the host compiler, no CUDA), one implementation per
// instruction set, selected once at load time.
//
// Why an AVX-512 variant: on the GPU boxes' Xeons the packing of a float64 utterance list (R/lasr/data/reader.py:24 ->
// the `.float()` of R/lasr/data/datatrans.py:73) is bound by how many cache misses ONE core keeps in flight, not by DRAM: with the
// SSE2 loop 8 threads move 72 GB/s (read + write); 64-byte loads, whole-line non-temporal stores and a software prefetch into L2
// 4 kB ahead of the loads (the L2 queue holds more outstanding lines than the L1 fill buffers) move 112 GB/s on the same cores
// (tools/pack_bench.cpp, measured in the build container, same CPU model as the GPU boxes).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace b200fe_host {

// ---- baseline (SSE2, every x86-64) ------------------------------------------------------------------------------------------
static void cvt_f64_f32_sse2(const double* __restrict__ s, float* __restrict__ d, long long n)
{
    long long i = 0;
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        for (; i + 8 <= n; i += 8) {
            const __m128 a = _mm_cvtpd_ps(_mm_loadu_pd(s + i)), b = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
            const __m128 c = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 4)), e = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 6));
            _mm_stream_ps(d + i, _mm_movelh_ps(a, b));          // the staging buffer is read next by the DMA engine, not by a core
            _mm_stream_ps(d + i + 4, _mm_movelh_ps(c, e));
        }
        _mm_sfence();
    }
    for (; i < n; ++i) d[i] = (float)s[i];
}

static void zero_sse2(void* dst, long long n)
{
    char* d = static_cast<char*>(dst);
    while (n > 0 && (reinterpret_cast<uintptr_t>(d) & 15) != 0) { *d++ = 0; --n; }
    const __m128i z = _mm_setzero_si128();
    long long i = 0;
    for (; i + 64 <= n; i += 64) {
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), z); _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), z); _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), z);
    }
    _mm_sfence();
    d += i; n -= i;
    if (n > 0) memset(d, 0, (size_t)n);
}

static void copy_sse2(const void* src, void* dst, long long n)
{
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    long long i = 0;
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
        }
        _mm_sfence();
    }
    if (i < n) memcpy(d + i, s + i, (size_t)(n - i));
}

// ---- AVX-512F: 64-byte loads, whole-line non-temporal stores, L2 prefetch kPrefetch bytes ahead of the loads -----------------
constexpr int kPrefetch = 4096;

__attribute__((target("avx512f"))) static void cvt_f64_f32_avx512(const double* __restrict__ s, float* __restrict__ d, long long n)
{
    long long i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 63) != 0) { d[i] = (float)s[i]; ++i; }     // (float)double rounds like cvtpd2ps
    for (; i + 16 <= n; i += 16) {
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch, _MM_HINT_T1);
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch + 64, _MM_HINT_T1);
        const __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i)), b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
        const __m512d v = _mm512_insertf64x4(_mm512_castpd256_pd512(_mm256_castps_pd(a)), _mm256_castps_pd(b), 1);
        _mm512_stream_ps(d + i, _mm512_castpd_ps(v));
    }
    _mm_sfence();
    for (; i < n; ++i) d[i] = (float)s[i];
}

__attribute__((target("avx512f"))) static void zero_avx512(void* dst, long long n)
{
    char* d = static_cast<char*>(dst);
    while (n > 0 && (reinterpret_cast<uintptr_t>(d) & 63) != 0) { *d++ = 0; --n; }
    const __m512i z = _mm512_setzero_si512();
    long long i = 0;
    for (; i + 64 <= n; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), z);
    _mm_sfence();
    if (i < n) memset(d + i, 0, (size_t)(n - i));
}

__attribute__((target("avx512f"))) static void copy_avx512(const void* src, void* dst, long long n)
{
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    long long i = 0;
    const long long head = (64 - (long long)(reinterpret_cast<uintptr_t>(d) & 63)) & 63;
    if (head > 0 && head <= n) { memcpy(d, s, (size_t)head); i = head; }
    if ((reinterpret_cast<uintptr_t>(d + i) & 63) == 0) {
        for (; i + 64 <= n; i += 64) {
            _mm_prefetch(s + i + kPrefetch, _MM_HINT_T1);
            _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), _mm512_loadu_si512(reinterpret_cast<const void*>(s + i)));
        }
        _mm_sfence();
    }
    if (i < n) memcpy(d + i, s + i, (size_t)(n - i));
}

// ---- float64 samples that hold 16-bit PCM values (k / 32768: what soundfile.read returns for PCM_16 files, R/lasr/data/reader.py:24)
// -> the int16 k itself: half the staging and PCIe bytes of float32, bit-identical features ((float)k == float(x) * 2^15).  Returns
// false as soon as the range holds a sample that is not such a value (the caller then repeats the batch as float32); d may be null
// (probe only).
static bool pcm16_sse2(const double* __restrict__ s, short* __restrict__ d, long long n)
{
    bool ok = true;
    for (long long i = 0; i < n; ++i) {
        const double v = s[i] * 32768.0;
        const bool in = v >= -32768.0 && v <= 32767.0;                 // NaN fails both
        const int k = in ? (int)__builtin_rint(v) : 0;
        ok = ok && in && (double)k == v;
        if (d) d[i] = (short)k;
    }
    return ok;
}

__attribute__((target("avx512f"))) static bool pcm16_avx512(const double* __restrict__ s, short* __restrict__ d, long long n)
{
    const __m512d sc = _mm512_set1_pd(32768.0), lo = _mm512_set1_pd(-32768.0), hi = _mm512_set1_pd(32767.0);
    long long i = 0;
    bool ok = true;
    if (d) while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 63) != 0) { ok = pcm16_sse2(s + i, d + i, 1) && ok; ++i; }
    unsigned bad = 0;
    for (; i + 32 <= n; i += 32) {
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch, _MM_HINT_T1);
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch + 64, _MM_HINT_T1);
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch + 128, _MM_HINT_T1);
        _mm_prefetch(reinterpret_cast<const char*>(s + i) + kPrefetch + 192, _MM_HINT_T1);
        __m256i k[4];
        for (int j = 0; j < 4; ++j) {
            const __m512d v = _mm512_mul_pd(_mm512_loadu_pd(s + i + 8 * j), sc);
            k[j] = _mm512_cvtpd_epi32(v);                                                  // round to nearest even (MXCSR default)
            const __mmask8 good = _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(k[j]), v, _CMP_EQ_OQ) & _mm512_cmp_pd_mask(v, lo, _CMP_GE_OQ) &
                                  _mm512_cmp_pd_mask(v, hi, _CMP_LE_OQ);
            bad |= (unsigned)(unsigned char)~good;
        }
        if (d) {
            const __m256i a = _mm512_cvtepi32_epi16(_mm512_inserti64x4(_mm512_castsi256_si512(k[0]), k[1], 1));
            const __m256i b = _mm512_cvtepi32_epi16(_mm512_inserti64x4(_mm512_castsi256_si512(k[2]), k[3], 1));
            _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), _mm512_inserti64x4(_mm512_castsi256_si512(a), b, 1));
        }
    }
    if (d) _mm_sfence();
    ok = ok && bad == 0;
    if (i < n) ok = pcm16_sse2(s + i, d ? d + i : nullptr, n - i) && ok;
    return ok;
}

// ---- dispatch -----------------------------------------------------------------------------------------------------------------
typedef void (*cvt_fn)(const double*, float*, long long);
typedef void (*zero_fn)(void*, long long);
typedef void (*copy_fn)(const void*, void*, long long);

static int pick_isa()
{
    // B200FE_HOST_ISA=sse2 forces the baseline (A/B runs, tools/pack_probe.py)
    const char* e = getenv("B200FE_HOST_ISA");
    if (e && strcmp(e, "sse2") == 0) return 0;
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? 1 : 0;
}

static const int g_isa = pick_isa();
static const cvt_fn g_cvt = g_isa ? cvt_f64_f32_avx512 : cvt_f64_f32_sse2;
static const zero_fn g_zero = g_isa ? zero_avx512 : zero_sse2;
static const copy_fn g_copy = g_isa ? copy_avx512 : copy_sse2;

typedef bool (*pcm_fn)(const double*, short*, long long);
static const pcm_fn g_pcm = g_isa ? pcm16_avx512 : pcm16_sse2;
bool cvt_f64_pcm16(const double* s, short* d, long long n) { return g_pcm(s, d, n); }
int host_isa() { return g_isa; }
void cvt_f64_f32(const double* s, float* d, long long n) { g_cvt(s, d, n); }
void zero_stream(void* dst, long long n) { g_zero(dst, n); }
void copy_stream(const void* src, void* dst, long long n) { g_copy(src, dst, n); }

}  // namespace b200fe_host

// C ABI of the B200 acoustic front end (see include/b200fe.h).  Host side: plan construction
// (window, FFT twiddles, mel tables in the layout the kernels want) and kernel launches.
#include "../../include/b200fe.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_pool.h"
#include "aux_kernels.cuh"
#include "stream_kernels.cuh"
#include "resample_kernels.cuh"
#include "fbank_kernel.cuh"
#include "fbank_instances.h"
#ifdef B200FE_WITH_WS          // the warp-specialised experiment (measured 30 % slower) is not part of the default build
#include "fbank_ws_kernel.cuh"
#endif

// the instantiations of the fused kernel live in fbank_inst.cu (one translation unit per group, built in parallel)
#define B200FE_X(g, n, s, p, d, i, m, l, a) extern template __global__ void b200fe::fbank_fused_kernel<n, s, p, d, i, m, l, a>(const __grid_constant__ b200fe::FbankArgs);
B200FE_FBANK_INSTANCES(B200FE_X)
#undef B200FE_X

#define B200FE_MEL_HOST_TABLES
#include "mel_static_default.inc"
#undef B200FE_MEL_HOST_TABLES

using namespace b200fe;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) return fail(B200FE_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

struct b200fe_plan {
    b200fe_opts o;
    int win, shift, nfft, nmel;
    int device;
    int num_sms;
    // device tables
    float* d_window_scaled = nullptr;   // window * 2^(bits-1)
    float* d_window_plain = nullptr;
    float2* d_twiddle = nullptr;
    float2* d_split_tw = nullptr;
    // host copies of the constant-bank tables
    short seg_start[kMaxMel + 3];
    short grp_begin[9];
    float2 w_updn[256];
    int tile_floats;
    int smem_bytes;
    int multi_tile_floats, multi_smem_bytes;   // multi-utterance tiles (streaming)
    int nload;          // 13 or 16
    int static_mel;     // 1: tables equal the baked-in LASR default -> straight-line phase B
    int ctas_per_sm;
    int use_ws;         // 1: the warp-specialised kernel (fbank_ws_kernel.cuh) serves this option set
    int ws_smem_bytes;
};

#ifdef B200FE_WITH_WS
static const void* ws_kernel(bool peak, bool i16)
{
    if (i16) return peak ? (const void*)fbank_ws_kernel<true, true> : (const void*)fbank_ws_kernel<false, true>;
    return peak ? (const void*)fbank_ws_kernel<true, false> : (const void*)fbank_ws_kernel<false, false>;
}
#endif
// Launch with programmatic stream serialisation (see pdl_wait in b200fe_common.cuh): the kernel may start while the preceding
// kernel of the stream drains; it synchronises with it through griddepcontrol.wait.
static cudaError_t launch_pdl(const void* fn, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t st)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    memset(&attr, 0, sizeof attr);
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

constexpr int kBuilderCtas = 16;     // CTAs of the work-list builder (each repeats the scan, writes 1/16 of the entries)
#ifdef B200FE_WITH_WS
static int plan_tile_frames(const b200fe_plan* p) { return p->use_ws ? kWsFT : kFT; }
#else
static int plan_tile_frames(const b200fe_plan*) { return kFT; }
#endif

static bool plan_has_multi(const b200fe_plan* p) { return p->nload == 13 && (p->static_mel || p->nfft == 256); }

static bool plan_has_lean(const b200fe_plan* p) { return p->nload == 13 && p->static_mel && p->nfft == 512; }

static const void* plan_kernel(const b200fe_plan* p, bool peak, bool i16 = false, bool multi = false, bool lean = false, bool apply = false)
{
    if (lean && apply) return (const void*)fbank_fused_kernel<13, true, false, false, false, false, true, true>;
    if (apply) return (const void*)fbank_fused_kernel<13, true, false, false, false, false, false, true>;       // completion tiles: SpecAugment mean fills
    if (lean) return (const void*)fbank_fused_kernel<13, true, false, false, false, false, true>;
    if (multi) {   // multi-utterance tiles (lock-step streaming): the two default option sets, float32, no peak normalisation
        if (p->nfft == 256) return (const void*)fbank_fused_kernel<13, false, false, true, false, true>;
        return (const void*)fbank_fused_kernel<13, true, false, false, false, true>;
    }
    if (i16) {   // int16 PCM input, 512-point family
        if (p->nload == 13) {
            if (p->static_mel) return peak ? (const void*)fbank_fused_kernel<13, true, true, false, true> : (const void*)fbank_fused_kernel<13, true, false, false, true>;
            return peak ? (const void*)fbank_fused_kernel<13, false, true, false, true> : (const void*)fbank_fused_kernel<13, false, false, false, true>;
        }
        return peak ? (const void*)fbank_fused_kernel<16, false, true, false, true> : (const void*)fbank_fused_kernel<16, false, false, false, true>;
    }
    if (p->nfft == 256) {   // 8 kHz family: two real frames per complex FFT
        if (p->nload == 13) return peak ? (const void*)fbank_fused_kernel<13, false, true, true> : (const void*)fbank_fused_kernel<13, false, false, true>;
        return peak ? (const void*)fbank_fused_kernel<16, false, true, true> : (const void*)fbank_fused_kernel<16, false, false, true>;
    }
    if (p->nload == 13) {
        if (p->static_mel) return peak ? (const void*)fbank_fused_kernel<13, true, true, false> : (const void*)fbank_fused_kernel<13, true, false, false>;
        return peak ? (const void*)fbank_fused_kernel<13, false, true, false> : (const void*)fbank_fused_kernel<13, false, false, false>;
    }
    return peak ? (const void*)fbank_fused_kernel<16, false, true, false> : (const void*)fbank_fused_kernel<16, false, false, false>;
}

extern "C" void b200fe_default_opts(b200fe_opts* o)
{
    memset(o, 0, sizeof *o);
    o->sample_frequency = 16000.f;
    o->frame_length_ms = 25.f;
    o->frame_shift_ms = 10.f;
    o->num_mel_bins = 80;
    o->low_freq = 20.f;
    o->high_freq = 0.f;
    o->preemphasis_coefficient = 0.97f;
    o->remove_dc_offset = 1;
    o->use_power = 1;
    o->use_log_fbank = 1;
    o->window_type = 0;
    o->blackman_coeff = 0.42f;
    o->audio_bit = 16;
    o->dither = 0.f;
}

extern "C" const char* b200fe_last_error(void) { return g_err.c_str(); }

static double mel_scale(double f) { return 1127.0 * std::log(1.0 + f / 700.0); }

extern "C" int b200fe_plan_create(const b200fe_opts* opts, b200fe_plan** out)
{
    if (!opts || !out) return fail(B200FE_EINVAL, "null argument");
    b200fe_plan* p = new b200fe_plan();
    p->o = *opts;
    const b200fe_opts& o = p->o;
    // TA:136-139
    p->shift = (int)(o.sample_frequency * o.frame_shift_ms * 0.001f);
    p->win = (int)(o.sample_frequency * o.frame_length_ms * 0.001f);
    p->nfft = next_pow2(p->win);
    p->nmel = o.num_mel_bins;
    auto bad = [&](const char* m) { delete p; return fail(B200FE_EINVAL, "%s", m); };
    if (p->win < 2 || p->shift <= 0) return bad("window size must be >= 2 and shift > 0 (TA:142-145)");
    if (p->nfft != 512 && p->nfft != 256) return bad("the padded window must be 512 samples (16 kHz family) or 256 samples (8 kHz family)");
    if (p->nmel <= 3 || p->nmel > kMaxMel) return bad("num_mel_bins must be in (3, 128] (TA:449)");
    if (o.preemphasis_coefficient < 0.f || o.preemphasis_coefficient > 1.f) return bad("preemphasis_coefficient must be in [0,1] (TA:149)");
    if (p->nfft == 512 && p->shift % 2 != 0) return bad("an odd frame shift (in samples) is not supported (8-byte aligned frame loads)");
    const int nbins = p->nfft / 2;
    const double nyq = 0.5 * o.sample_frequency;
    double high = o.high_freq;
    if (high <= 0.0) high += nyq;   // TA:457-458
    if (!(0.0 <= o.low_freq && o.low_freq < nyq && 0.0 < high && high <= nyq && o.low_freq < high))
        return bad("bad low_freq / high_freq versus Nyquist (TA:460-462)");

    // ---- window (TA:86-113), double precision unless the caller supplied its own table ----
    std::vector<float> window(512, 0.f), window_s(512, 0.f);   // zero-extended (the 256-point family uses the first 256)
    const double scale = std::ldexp(1.0, o.audio_bit - 1);
    for (int n = 0; n < p->win; ++n) {
        double w;
        if (o.window) w = o.window[n];
        else {
            const double a = 2.0 * M_PI / (p->win - 1);
            switch (o.window_type) {
                case 0: w = std::pow(0.5 - 0.5 * std::cos(a * n), 0.85); break;
                case 1: w = 0.5 - 0.5 * std::cos(a * n); break;
                case 2: w = 0.54 - 0.46 * std::cos(a * n); break;
                case 3: w = 1.0; break;
                case 4: w = o.blackman_coeff - 0.5 * std::cos(a * n) + (0.5 - o.blackman_coeff) * std::cos(2 * a * n); break;
                default: return bad("unknown window_type");
            }
        }
        window[n] = (float)w;
        window_s[n] = (float)w * (float)scale;   // exact: power of two
    }
    // ---- mel filterbank (TA:436-511, vtln_warp == 1) ----
    std::vector<float> W((size_t)p->nmel * nbins, 0.f);
    const double mel_lo = mel_scale(o.low_freq), mel_hi = mel_scale(high);
    const double delta = (mel_hi - mel_lo) / (p->nmel + 1);
    const double bin_w = (double)o.sample_frequency / p->nfft;
    if (o.mel_weights) memcpy(W.data(), o.mel_weights, W.size() * sizeof(float));
    else
        for (int j = 0; j < p->nmel; ++j) {
            const double left = mel_lo + j * delta, center = mel_lo + (j + 1) * delta, right = mel_lo + (j + 2) * delta;
            for (int k = 0; k < nbins; ++k) {
                const double m = mel_scale(bin_w * k);
                const double up = (m - left) / (center - left), down = (right - m) / (right - center);
                const double w = std::fmax(0.0, std::fmin(up, down));
                W[(size_t)j * nbins + k] = (float)w;
            }
        }
    // segment structure: seg(k) = number of centers strictly below mel(k); bin seg gets the
    // up-slope weight, bin seg-1 the down-slope weight; every other entry of column k must be 0.
    for (int k = 0; k < 256; ++k) p->w_updn[k] = make_float2(0.f, 0.f);
    std::vector<int> seg(nbins);
    for (int k = 0; k < nbins; ++k) {
        const double m = mel_scale(bin_w * k);
        int s = 0;
        if (m <= mel_lo) s = -1;                       // left of the first filter
        else { while (s <= p->nmel && mel_lo + (s + 1) * delta < m) ++s; }
        seg[k] = s;
    }
    for (int k = 0; k < nbins; ++k) {
        float up = 0.f, dn = 0.f;
        int s = seg[k];
        // tolerate a one-off between this double computation and a float32 table from the caller
        auto nz = [&](int j) { return j >= 0 && j < p->nmel && W[(size_t)j * nbins + k] != 0.f; };
        if (s >= 0 && !(nz(s) || nz(s - 1)) ) { if (nz(s + 1) || nz(s)) s = s + 1; else if (nz(s - 2)) s = s - 1; }
        if (s > p->nmel) s = p->nmel + 1;
        if (s >= 0 && s <= p->nmel) {
            if (s < p->nmel) up = W[(size_t)s * nbins + k];
            if (s >= 1) dn = W[(size_t)(s - 1) * nbins + k];
        }
        for (int j = 0; j < p->nmel; ++j)
            if (j != s && j != s - 1 && W[(size_t)j * nbins + k] != 0.f) return bad("mel weights are not a two-band triangular filterbank");
        seg[k] = s;
        p->w_updn[k] = make_float2(up * 0.25f, dn * 0.25f);   // |X|^2 arrives as 4|X|^2 (power) ...
        if (!o.use_power) p->w_updn[k] = make_float2(up * 0.5f, dn * 0.5f);   // ... or 2|X| (magnitude)
    }
    // monotone segments -> seg_start
    for (int k = 1; k < nbins; ++k) if (seg[k] < seg[k - 1] && seg[k] >= 0) return bad("mel segments are not monotone");
    {
        int k = 0;
        for (int s = 0; s <= p->nmel + 1; ++s) {
            while (k < nbins && seg[k] < s) ++k;
            p->seg_start[s] = (short)k;
        }
        p->seg_start[p->nmel + 2] = (short)nbins;
    }
    // warp groups: 8 consecutive runs of bins with balanced (#k read + per-bin epilogue) cost
    {
        std::vector<double> cost(p->nmel);
        double tot = 0;
        for (int j = 0; j < p->nmel; ++j) { cost[j] = (p->seg_start[j + 2] - p->seg_start[j]) * 0.5 + 6.0; tot += cost[j]; }
        int j = 0; double acc = 0;
        p->grp_begin[0] = 0;
        for (int w = 1; w < kMelGroups; ++w) {
            const double target = tot * w / (double)kMelGroups;
            while (j < p->nmel && acc + cost[j] * 0.5 < target) { acc += cost[j]; ++j; }
            p->grp_begin[w] = (short)j;
        }
        for (int w = kMelGroups; w <= 8; ++w) p->grp_begin[w] = (short)p->nmel;
    }
    // ---- FFT twiddles (double precision, rounded once) ----
    std::vector<float2> twd(256), stw(256);
    for (int n1 = 0; n1 < 16; ++n1)
        for (int k = 0; k < 16; ++k) {
            const double ang = -2.0 * M_PI * (double)((n1 * k) % 256) / 256.0;
            twd[n1 * 16 + k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
    for (int k = 0; k < 256; ++k) {   // -j * exp(-2 pi j k / 512) = -sin(t) - j cos(t), t = 2 pi k / 512
        const double t = 2.0 * M_PI * k / 512.0;
        stw[k] = make_float2((float)(-std::sin(t)), (float)(-std::cos(t)));
    }
    const int per_reg = p->nfft == 512 ? 32 : 16;     // samples covered by one register slot of the 16 lanes
    p->nload = (p->win > 12 * per_reg && p->win <= 13 * per_reg) ? 13 : 16;
    // a frame reads up to per_reg*nload samples from its start (zero-weighted past the window)
    p->tile_floats = ((kFT - 1) * p->shift + per_reg * p->nload + 3) & ~3;
    // multi-utterance tiles (lock-step streaming) get their own, larger tile buffer: room for kFT/4 utterances of 4 frames
    // (a 40 ms push); only those launches pay the extra shared memory
    p->multi_tile_floats = std::max(p->tile_floats, (kFT / 4) * ((3 * p->shift + per_reg * p->nload + 3) & ~3));
    p->multi_smem_bytes = make_layout(p->multi_tile_floats, p->nmel).total;
    p->smem_bytes = make_layout(p->tile_floats, p->nmel).total;
    // experiments only: B200FE_EXTRA_SMEM=<bytes> inflates the request so that fewer CTAs are resident (occupancy scaling probes)
    if (const char* xs = getenv("B200FE_EXTRA_SMEM")) { p->smem_bytes += atoi(xs); p->multi_smem_bytes += atoi(xs); }

    cudaError_t e = cudaGetDevice(&p->device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, p->device);
    if (e != cudaSuccess) { delete p; return fail(B200FE_ECUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
    p->num_sms = prop.multiProcessorCount;
    if ((size_t)p->smem_bytes > prop.sharedMemPerBlockOptin) { delete p; return fail(B200FE_EINVAL, "tile does not fit in shared memory (%d B)", p->smem_bytes); }
    auto up = [&](void** d, const void* h, size_t n) -> cudaError_t {
        cudaError_t r = cudaMalloc(d, n);
        if (r != cudaSuccess) return r;
        return cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice);
    };
    if ((e = up((void**)&p->d_window_scaled, window_s.data(), 512 * 4)) != cudaSuccess ||
        (e = up((void**)&p->d_window_plain, window.data(), 512 * 4)) != cudaSuccess ||
        (e = up((void**)&p->d_twiddle, twd.data(), 256 * 8)) != cudaSuccess ||
        (e = up((void**)&p->d_split_tw, stw.data(), 256 * 8)) != cudaSuccess) {
        b200fe_plan_destroy(p);
        return fail(B200FE_ECUDA, "plan upload: %s", cudaGetErrorString(e));
    }
    p->static_mel = 0;
    if (p->nfft == 512 && p->nload == 13 && p->nmel == B200FE_STATIC_NMEL && o.use_power && o.dither == 0.f) {
        // (the straight-line code carries its own bin groups, kStaticGrpBegin: min-max balanced by the generator; p->grp_begin
        // only serves the generic phase B)
        bool same = memcmp(p->seg_start, kStaticSegStart, sizeof(short) * (p->nmel + 3)) == 0;
        (void)kStaticGrpBegin;
        for (int k = 0; same && k < 256; ++k) same = (p->w_updn[k].x == kStaticUp[k] && p->w_updn[k].y == kStaticDn[k]);
        p->static_mel = same ? 1 : 0;
    }
    // Several plans (option sets) can share one kernel instantiation with different shared-memory sizes:
    // raise the limit to the device maximum once instead of to this plan's size.
    const void* kfn = plan_kernel(p, false);
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel(p, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    if (e == cudaSuccess && plan_has_lean(p)) {
        e = cudaFuncSetAttribute(plan_kernel(p, false, false, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel(p, false, false, false, true, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel(p, false, false, false, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    }
    if (e == cudaSuccess && plan_has_multi(p))
        e = cudaFuncSetAttribute(plan_kernel(p, false, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    if (p->nfft == 512) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel(p, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel(p, true, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    }
    if (e != cudaSuccess) { b200fe_plan_destroy(p); return fail(B200FE_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, kThreads, p->smem_bytes);
    if (e != cudaSuccess || occ < 1) { b200fe_plan_destroy(p); return fail(B200FE_ECUDA, "fbank kernel cannot be resident (occ=%d): %s", occ, cudaGetErrorString(e)); }
    p->ctas_per_sm = occ;
    // Experimental: B200FE_WS=1 routes LASR's default option set to the warp-specialised kernel (fbank_ws_kernel.cuh).
    // Bit-identical output, but measured 25-30 % SLOWER than the phase-ordered kernel on B200, so it is
    // opt-in and kept for A/B runs only.
    p->use_ws = 0;
    p->ws_smem_bytes = 0;
#ifdef B200FE_WITH_WS
    p->ws_smem_bytes = ws_layout(p->shift, p->win).total;
    const char* want_ws = getenv("B200FE_WS");
    if (p->static_mel && want_ws && want_ws[0] == '1' && (size_t)p->ws_smem_bytes <= prop.sharedMemPerBlockOptin) {
        for (int v = 0; v < 4 && e == cudaSuccess; ++v)
            e = cudaFuncSetAttribute(ws_kernel(v & 1, v & 2), cudaFuncAttributeMaxDynamicSharedMemorySize, p->ws_smem_bytes);
        if (e != cudaSuccess) { b200fe_plan_destroy(p); return fail(B200FE_ECUDA, "cudaFuncSetAttribute (ws): %s", cudaGetErrorString(e)); }
        p->use_ws = 1;
    }
#endif
    *out = p;
    return B200FE_OK;
}

extern "C" void b200fe_plan_destroy(b200fe_plan* p)
{
    if (!p) return;
    cudaFree(p->d_window_scaled);
    cudaFree(p->d_window_plain);
    cudaFree(p->d_twiddle);
    cudaFree(p->d_split_tw);
    delete p;
}

extern "C" int b200fe_window_size(const b200fe_plan* p) { return p->win; }
extern "C" int b200fe_window_shift(const b200fe_plan* p) { return p->shift; }
extern "C" int b200fe_padded_window_size(const b200fe_plan* p) { return p->nfft; }
extern "C" int b200fe_plan_info(const b200fe_plan* p, int what)
{
    switch (what) {
        case 0: return p->static_mel;
        case 1: return p->nload;
        case 2: return p->smem_bytes;
        case 3: return p->ctas_per_sm;
        case 4: return p->num_sms;
        case 5: return plan_tile_frames(p);
        case 6: return p->use_ws;
        case 7: return (plan_has_lean(p) && !p->use_ws) ? 1 : 0;
        case 8: return kApplyRows;
        case 9: return kWarps;
        default: return -1;
    }
}
extern "C" long long b200fe_num_frames(const b200fe_plan* p, long long n)
{
    return n < p->win ? 0 : 1 + (n - p->win) / p->shift;   // TA:63-67
}

extern "C" int b200fe_build_tile_table_padded(const b200fe_plan* p, const long long* nsamp, int batch, int max_frames, int* table, int capacity)
{
    if (!p || !nsamp || batch < 0 || max_frames <= 0) return fail(B200FE_EINVAL, "build_tile_table_padded: bad argument");
    if (p->use_ws) return fail(B200FE_EINVAL, "build_tile_table_padded: not available with the experimental warp-specialised kernel");
    long long n = 0;
    for (int u = 0; u < batch; ++u) {
        const long long T = std::min<long long>(b200fe_num_frames(p, nsamp[u]), max_frames);
        for (long long f0 = 0; f0 < T; f0 += plan_tile_frames(p), ++n)
            if (table && n < capacity) { table[2 * n] = u; table[2 * n + 1] = (int)f0; }
        for (long long r0 = T; r0 < max_frames; r0 += kPadTileRows, ++n)      // padding tiles follow the utterance's frames
            if (table && n < capacity) { table[2 * n] = u; table[2 * n + 1] = (int)(-r0 - 1); }
    }
    if (n > 0x7fffffffLL) return fail(B200FE_EINVAL, "build_tile_table_padded: too many tiles");
    return (int)n;
}

extern "C" int b200fe_tile_table_capacity(const b200fe_plan* p, int batch, int max_frames, int with_pads)
{
    if (!p || batch < 0 || max_frames <= 0) return fail(B200FE_EINVAL, "tile_table_capacity: bad argument");
    const long long per = (max_frames + plan_tile_frames(p) - 1) / plan_tile_frames(p) + (with_pads ? (max_frames + kPadTileRows - 1) / kPadTileRows : 0);
    const long long n = per * batch;
    if (n > 0x7fffffffLL) return fail(B200FE_EINVAL, "tile_table_capacity: too many tiles");
    return (int)std::max<long long>(n, 1);
}

static cudaError_t launch_builder(const b200fe_plan* p, const long long* d_nsamp, int batch, int max_frames, int with_pads, int* d_table, int capacity,
                                  int* d_n_tiles, int* d_work_counter, int apply_lag, int* d_utt_done, void* d_zero, long long zero_bytes, void* stream)
{
    int win = p->win, shift = p->shift, ft = plan_tile_frames(p), pads = with_pads ? 1 : 0, pad_rows = kPadTileRows, apply_bit = kApplyBit, apply_rows = kApplyRows;
    if (apply_lag < 0) { apply_lag = -apply_lag; apply_rows = 0x20000000; }      // ONE completion tile per utterance that has frames
    int2* table = reinterpret_cast<int2*>(d_table);
    uint4* zero16 = reinterpret_cast<uint4*>(d_zero);
    long long n_zero16 = zero_bytes / 16;
    void* args[] = {(void*)&d_nsamp, (void*)&batch, (void*)&win, (void*)&shift, (void*)&ft, (void*)&max_frames, (void*)&pads, (void*)&pad_rows, (void*)&table,
                    (void*)&capacity, (void*)&d_n_tiles, (void*)&d_work_counter, (void*)&apply_lag, (void*)&apply_bit, (void*)&apply_rows, (void*)&d_utt_done,
                    (void*)&zero16, (void*)&n_zero16};
    return launch_pdl((const void*)build_tile_table_kernel, dim3(kBuilderCtas), dim3(1024), args, 0, (cudaStream_t)stream);
}

extern "C" int b200fe_build_tile_table_device(const b200fe_plan* p, const long long* d_nsamp, int batch, int max_frames, int with_pads,
                                              int* d_table, int capacity, int* d_n_tiles, int* d_work_counter, void* stream)
{
    if (!p || !d_nsamp || !d_table || !d_n_tiles || batch <= 0 || max_frames <= 0 || capacity <= 0)
        return fail(B200FE_EINVAL, "build_tile_table_device: bad argument");
    if (p->use_ws && with_pads) return fail(B200FE_EINVAL, "build_tile_table_device: padding tiles are not available with the experimental kernel");
    CUDA_TRY(launch_builder(p, d_nsamp, batch, max_frames, with_pads, d_table, capacity, d_n_tiles, d_work_counter, 0, nullptr, nullptr, 0, stream));
    return B200FE_OK;
}

extern "C" int b200fe_build_work_list_device(const b200fe_plan* p, const long long* d_nsamp, int batch, int max_frames, int with_pads, int apply_lag,
                                             int* d_table, int capacity, int* d_n_tiles, int* d_work_counter, int* d_utt_done,
                                             void* d_zero, long long zero_bytes, void* stream)
{
    if (!p || !d_nsamp || !d_table || !d_n_tiles || !d_work_counter || batch <= 0 || max_frames <= 0 || capacity <= 0 || apply_lag < -(1 << 20))
        return fail(B200FE_EINVAL, "build_work_list_device: bad argument");
    if (p->use_ws) return fail(B200FE_EINVAL, "build_work_list_device: not available with the experimental kernel");
    if (apply_lag != 0 && !d_utt_done) return fail(B200FE_EINVAL, "build_work_list_device: apply tiles need d_utt_done");
    if (apply_lag != 0 && !plan_has_lean(p)) return fail(B200FE_EINVAL, "build_work_list_device: apply tiles need the default option set (plan_info 7)");
    if (zero_bytes < 0 || (zero_bytes > 0 && (!d_zero || (zero_bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(d_zero) & 15) != 0)))
        return fail(B200FE_EINVAL, "build_work_list_device: d_zero must be 16-byte aligned and zero_bytes a multiple of 16");
    CUDA_TRY(launch_builder(p, d_nsamp, batch, max_frames, with_pads, d_table, capacity, d_n_tiles, d_work_counter, apply_lag,
                            apply_lag != 0 ? d_utt_done : nullptr, d_zero, zero_bytes, stream));
    return B200FE_OK;
}

extern "C" int b200fe_build_tile_table(const b200fe_plan* p, const long long* nsamp, int batch, int* table, int capacity)
{
    if (!p || !nsamp || batch < 0) return fail(B200FE_EINVAL, "build_tile_table: bad argument");
    long long n = 0;
    for (int u = 0; u < batch; ++u) {
        const long long T = b200fe_num_frames(p, nsamp[u]);
        for (long long f0 = 0; f0 < T; f0 += plan_tile_frames(p), ++n)
            if (table && n < capacity) { table[2 * n] = u; table[2 * n + 1] = (int)f0; }
    }
    if (n > 0x7fffffffLL) return fail(B200FE_EINVAL, "build_tile_table: too many tiles");
    return (int)n;
}

extern "C" int b200fe_peak_absmax_i16(const b200fe_plan* plan, const short* d_wav, long long wav_stride, const long long* d_wav_offsets,
                                      const long long* d_nsamp, int batch, float* d_peak, void* stream)
{
    if (!plan || !d_wav || !d_nsamp || !d_peak || batch <= 0 || wav_stride <= 0) return fail(B200FE_EINVAL, "peak_absmax_i16: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(d_peak, 0, sizeof(float) * batch, st));
    const long long chunk = 256LL * 4 * 8;
    dim3 grid((unsigned)((wav_stride + chunk - 1) / chunk), (unsigned)batch);
    absmax_i16_kernel<<<grid, 256, 0, st>>>(d_wav, wav_stride, d_wav_offsets, d_nsamp, d_peak);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_peak_absmax(const b200fe_plan* plan, const float* d_wav, long long wav_stride, const long long* d_wav_offsets,
                                  const long long* d_nsamp, int batch, float* d_peak, void* stream)
{
    if (!plan || !d_wav || !d_nsamp || !d_peak || batch <= 0 || wav_stride <= 0) return fail(B200FE_EINVAL, "peak_absmax: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(d_peak, 0, sizeof(float) * batch, st));
    const long long chunk = 256LL * 4 * 8;
    dim3 grid((unsigned)((wav_stride + chunk - 1) / chunk), (unsigned)batch);
    absmax_kernel<<<grid, 256, 0, st>>>(d_wav, wav_stride, d_wav_offsets, d_nsamp, d_peak);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_h2d_ragged(const void* h_wav, long long h_stride, const long long* h_nsamp, const long long* h_offsets,
                                 int batch, void* d_packed, int elem_bytes, void* stream)
{
    if (!h_wav || !h_nsamp || !h_offsets || !d_packed || batch < 0 || (elem_bytes != 2 && elem_bytes != 4)) return fail(B200FE_EINVAL, "h2d_ragged: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int u = 0; u < batch; ++u) {
        if (h_nsamp[u] <= 0) continue;
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(d_packed) + h_offsets[u] * elem_bytes,
                                 static_cast<const char*>(h_wav) + (long long)u * h_stride * elem_bytes,
                                 (size_t)elem_bytes * (size_t)h_nsamp[u], cudaMemcpyHostToDevice, st));
    }
    return B200FE_OK;
}

// ---- host staging (include/b200fe.h): thread pool that packs / converts utterance lists and zero-fills padding rows ----
extern "C" int b200fe_host_pool_create(int n_threads, b200fe_host_pool** pool)
{
    if (!pool) return fail(B200FE_EINVAL, "host_pool_create: null argument");
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    n_threads = std::max(1, std::min(n_threads, 256));
    try { *pool = new b200fe_host_pool(n_threads); }
    catch (const std::exception& e) { return fail(B200FE_EINVAL, "host_pool_create: %s", e.what()); }
    return B200FE_OK;
}
extern "C" void b200fe_host_pool_destroy(b200fe_host_pool* pool) { delete pool; }
extern "C" int b200fe_host_pool_threads(const b200fe_host_pool* pool) { return pool ? (int)pool->threads.size() : 0; }
extern "C" int b200fe_host_isa(void) { return b200fe_host::host_isa(); }
extern "C" int b200fe_host_ndarray_data(const long long* py_objects, int n, void** data_out, int data_offset)
{
    if (!py_objects || !data_out || n < 0 || data_offset < 0 || (data_offset & 7) != 0) return fail(B200FE_EINVAL, "host_ndarray_data: bad argument");
    for (int i = 0; i < n; ++i) {
        if (py_objects[i] == 0) return fail(B200FE_EINVAL, "host_ndarray_data: null object %d", i);
        data_out[i] = *reinterpret_cast<void* const*>(static_cast<uintptr_t>(py_objects[i]) + (uintptr_t)data_offset);
    }
    return B200FE_OK;
}

static long long host_pack_submit(b200fe_host_pool* pool, const void* const* h_src, const long long* nsamp, int batch, int src_dtype,
                                  void* h_dst, const long long* dst_offsets, long long dst_capacity, std::function<void()> on_done)
{
    if (!pool || !h_src || !nsamp || !h_dst || !dst_offsets || batch < 0) return fail(B200FE_EINVAL, "host_pack: bad argument");
    if (src_dtype < 0 || src_dtype > 3) return fail(B200FE_EINVAL, "host_pack: src_dtype must be 0 (float32), 1 (int16), 2 (float64) or 3 (float64 holding PCM16 values -> int16)");
    if ((reinterpret_cast<uintptr_t>(h_dst) & 15) != 0) return fail(B200FE_EINVAL, "host_pack: the staging buffer must be 16-byte aligned");
    const long long dsz = (src_dtype == 1 || src_dtype == 3) ? 2 : 4, ssz = src_dtype == 1 ? 2 : src_dtype >= 2 ? 8 : 4, al = 16 / dsz;
    const long long chunk = (256LL << 10) / dsz;                  // elements per task: 256 kB of destination
    std::vector<b200fe_host::Task> tasks;
    for (int u = 0; u < batch; ++u) {
        const long long n = nsamp[u], o = dst_offsets[u];
        if (n < 0 || o < 0 || (o % al) != 0 || !h_src[u]) return fail(B200FE_EINVAL, "host_pack: utterance %d: bad length, offset or pointer", u);
        const long long padded = (n + al - 1) / al * al;
        if (o + padded > dst_capacity) return fail(B200FE_EINVAL, "host_pack: utterance %d does not fit the staging buffer", u);
        for (long long c = 0; c < n || c == 0; c += chunk) {
            const long long m = std::min(chunk, n - c);
            b200fe_host::Task t;
            t.kind = src_dtype == 2 ? 1 : src_dtype == 3 ? 3 : 0;
            t.src = static_cast<const char*>(h_src[u]) + c * ssz;
            t.dst = static_cast<char*>(h_dst) + (o + c) * dsz;
            t.n = src_dtype >= 2 ? m : m * dsz;
            t.tail_zero = (c + m >= n) ? (padded - n) * dsz : 0;
            tasks.push_back(t);
            if (n == 0) break;
        }
    }
    if (tasks.empty()) { b200fe_host::Task t; t.kind = 2; t.src = nullptr; t.dst = h_dst; t.n = 0; t.tail_zero = 0; tasks.push_back(t); }
    return pool->submit(tasks, std::move(on_done));
}

extern "C" long long b200fe_host_pack_begin(b200fe_host_pool* pool, const void* const* h_src, const long long* nsamp, int batch, int src_dtype,
                                            void* h_dst, const long long* dst_offsets, long long dst_capacity)
{
    return host_pack_submit(pool, h_src, nsamp, batch, src_dtype, h_dst, dst_offsets, dst_capacity, nullptr);
}

extern "C" long long b200fe_host_pack_copy_begin(b200fe_host_pool* pool, const void* const* h_src, const long long* nsamp, int batch, int src_dtype,
                                                 void* h_dst, const long long* dst_offsets, long long dst_capacity,
                                                 void* d_dst, long long copy_elems, int device, void* copy_stream, void* copy_event)
{
    if (!pool || !d_dst || copy_elems < 0 || batch <= 0 || !dst_offsets) return fail(B200FE_EINVAL, "host_pack_copy: bad argument");
    if (pool->device.load() != device) {
        // the workers attach to the device's primary context before their next task (once per thread), not inside an upload
        pool->bind_device = [](int dev) { cudaSetDevice(dev); cudaFree(nullptr); };
        pool->device.store(device);
    }
    const long long dsz = (src_dtype == 1 || src_dtype == 3) ? 2 : 4;
    const char* hs = static_cast<const char*>(h_dst) + dst_offsets[0] * dsz;
    char* dd = static_cast<char*>(d_dst) + dst_offsets[0] * dsz;
    const size_t bytes = (size_t)(copy_elems * dsz);
    auto upload = [=]() {
        // runs on a pool thread: the runtime API needs the stream's device current on THIS thread
        cudaSetDevice(device);
        if (bytes > 0) cudaMemcpyAsync(dd, hs, bytes, cudaMemcpyHostToDevice, (cudaStream_t)copy_stream);
        if (copy_event) cudaEventRecord((cudaEvent_t)copy_event, (cudaStream_t)copy_stream);
    };
    return host_pack_submit(pool, h_src, nsamp, batch, src_dtype, h_dst, dst_offsets, dst_capacity, upload);
}

extern "C" long long b200fe_host_zero_rows_begin(b200fe_host_pool* pool, float* h_feats, int batch, long long utt_rows, long long row_elems,
                                                 const long long* valid_rows, int elem_bytes)
{
    if (!pool || !h_feats || !valid_rows || batch < 0 || utt_rows < 0 || row_elems <= 0 || (elem_bytes != 2 && elem_bytes != 4))
        return fail(B200FE_EINVAL, "host_zero_rows: bad argument");
    std::vector<b200fe_host::Task> tasks;
    const long long row_bytes = row_elems * elem_bytes, chunk = 1LL << 20;
    for (int u = 0; u < batch; ++u) {
        const long long v = std::max(0LL, std::min(valid_rows[u], utt_rows));
        char* base = reinterpret_cast<char*>(h_feats) + ((long long)u * utt_rows + v) * row_bytes;
        const long long nb = (utt_rows - v) * row_bytes;
        for (long long c = 0; c < nb; c += chunk) {
            b200fe_host::Task t; t.kind = 2; t.src = nullptr; t.dst = base + c; t.n = std::min(chunk, nb - c); t.tail_zero = 0;
            tasks.push_back(t);
        }
    }
    if (tasks.empty()) { b200fe_host::Task t; t.kind = 2; t.src = nullptr; t.dst = h_feats; t.n = 0; t.tail_zero = 0; tasks.push_back(t); }
    return pool->submit(tasks);
}

extern "C" long long b200fe_host_zero_ranges_begin(b200fe_host_pool* pool, void* h_base, const long long* offsets, const long long* nbytes, int n)
{
    if (!pool || !h_base || n < 0 || (n > 0 && (!offsets || !nbytes))) return fail(B200FE_EINVAL, "host_zero_ranges: bad argument");
    std::vector<b200fe_host::Task> tasks;
    const long long chunk = 1LL << 20;
    for (int i = 0; i < n; ++i) {
        if (offsets[i] < 0 || nbytes[i] < 0) return fail(B200FE_EINVAL, "host_zero_ranges: range %d is negative", i);
        for (long long c = 0; c < nbytes[i]; c += chunk) {
            b200fe_host::Task t; t.kind = 2; t.src = nullptr; t.dst = static_cast<char*>(h_base) + offsets[i] + c; t.n = std::min(chunk, nbytes[i] - c); t.tail_zero = 0;
            tasks.push_back(t);
        }
    }
    if (tasks.empty()) { b200fe_host::Task t; t.kind = 2; t.src = nullptr; t.dst = h_base; t.n = 0; t.tail_zero = 0; tasks.push_back(t); }
    return pool->submit(tasks);
}

extern "C" int b200fe_host_wait_flag(b200fe_host_pool* pool, long long ticket, int* flag)
{
    if (!pool || ticket <= 0 || !flag) return fail(B200FE_EINVAL, "host_wait_flag: bad argument");
    if (pool->wait(ticket, flag) != 0) return fail(B200FE_EINVAL, "host_wait_flag: unknown ticket %lld", ticket);
    return B200FE_OK;
}

extern "C" int b200fe_host_pcm16_probe(const void* const* h_src, const long long* nsamp, int batch, int windows, int window_len)
{
    if (!h_src || !nsamp || batch < 0 || windows <= 0 || window_len <= 0) return fail(B200FE_EINVAL, "host_pcm16_probe: bad argument");
    for (int u = 0; u < batch; ++u) {
        const long long n = nsamp[u];
        if (n < 0 || (n > 0 && !h_src[u])) return fail(B200FE_EINVAL, "host_pcm16_probe: utterance %d: bad length or pointer", u);
        const double* s = static_cast<const double*>(h_src[u]);
        for (int w = 0; w < windows; ++w) {
            // windows spread over the utterance: leading digital silence holds PCM16 values whatever follows
            const long long len = std::min<long long>(window_len, n);
            const long long start = windows > 1 ? (n - len) * w / (windows - 1) : 0;
            if (len > 0 && !b200fe_host::cvt_f64_pcm16(s + start, nullptr, len)) return 0;
        }
    }
    return 1;
}

extern "C" int b200fe_host_wait(b200fe_host_pool* pool, long long ticket)
{
    if (!pool || ticket <= 0) return fail(B200FE_EINVAL, "host_wait: bad argument");
    if (pool->wait(ticket) != 0) return fail(B200FE_EINVAL, "host_wait: unknown ticket %lld", ticket);
    return B200FE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Host-only SpecAugment planner: replays the reference's draws on caller-held MT19937 states.
//   * CPython `random.randrange(a, a + n)` (lasr/utils/specaugment.py:23-24,64,95): _randbelow_with_getrandbits,
//     k = n.bit_length(), r = genrand_uint32() >> (32 - k) until r < n;
//   * numpy legacy `numpy.random.randint(0, high, size)` (specaugment.py:61,90): masked rejection on 32-bit draws,
//     val = next_uint32() & mask until val <= high - 1 (numpy/random/src/distributions: buffered_bounded_masked_uint32).
// Both generators are MT19937 with the standard tempering; their states are what random.getstate() / numpy.random.get_state()
// return (624 key words + position) and are advanced in place.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Mt19937 {
    unsigned int* key;
    int pos;
    unsigned int next32()
    {
        if (pos >= 624) {
            for (int kk = 0; kk < 624; ++kk) {
                const unsigned int y = (key[kk] & 0x80000000u) | (key[(kk + 1) % 624] & 0x7fffffffu);
                key[kk] = key[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        unsigned int y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    int py_randbelow(int n)                  // random.randrange(0, n), n > 0
    {
        int k = 0;
        while ((n >> k) != 0) ++k;
        unsigned int r;
        do { r = next32() >> (32 - k); } while (r >= (unsigned int)n);
        return (int)r;
    }
    int np_randint(int high)                 // numpy.random.randint(0, high), high > 0
    {
        const unsigned int rng = (unsigned int)high - 1u;
        if (rng == 0) return 0;
        unsigned int mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        unsigned int v;
        while ((v = (next32() & mask)) > rng) {}
        return (int)v;
    }
};
}  // namespace

extern "C" int b200fe_specaug_plan(unsigned int* py_key, int* py_pos, unsigned int* np_key, int* np_pos, const long long* frames, int batch,
                                   int num_mel, int max_freq_width, int n_freq_mask, int max_time_width, int n_time_mask,
                                   int draw_time_warp, int max_time_warp, int* masks, int* row_bounds, int* warps)
{
    if (!py_key || !py_pos || !np_key || !np_pos || (!frames && batch > 0) || batch < 0 || !masks || !row_bounds)
        return fail(B200FE_EINVAL, "specaug_plan: bad argument");
    if (*py_pos < 0 || *py_pos > 624 || *np_pos < 0 || *np_pos > 624) return fail(B200FE_EINVAL, "specaug_plan: bad generator position");
    if (n_freq_mask < 0 || n_freq_mask > B200FE_MAX_FREQ_MASKS || n_time_mask < 0 || n_time_mask > B200FE_MAX_TIME_MASKS)
        return fail(B200FE_EINVAL, "specaug_plan: too many masks");
    if (max_freq_width <= 0 || max_time_width <= 0 || num_mel < max_freq_width || max_time_warp < 0)
        return fail(B200FE_EINVAL, "specaug_plan: bad mask widths");
    Mt19937 py{py_key, *py_pos}, np_{np_key, *np_pos};
    const int nm = n_freq_mask + n_time_mask, W = max_time_warp;
    for (int u = 0; u < batch; ++u) {
        const long long T = frames[u];
        int* mk = masks + (long long)u * nm * 2;
        int* rb = row_bounds + (long long)u * 2 * n_time_mask;
        // time warp first (specaugment.py:20-24): no draw when T - W <= W
        int center = -1, warped = -1;
        if (draw_time_warp && T - W > W) {
            center = W + py.py_randbelow((int)(T - 2 * W));
            warped = center - W + py.py_randbelow(2 * W) + 1;
        }
        if (warps) { warps[2 * u] = center; warps[2 * u + 1] = warped; }
        // frequency masks (specaugment.py:61-74): one numpy call for the (n, 2) table, then a start per row
        int fs[2 * B200FE_MAX_FREQ_MASKS];
        for (int i = 0; i < 2 * n_freq_mask; ++i) fs[i] = np_.np_randint(max_freq_width);
        for (int i = 0; i < n_freq_mask; ++i) {
            const int f = fs[2 * i], w = fs[2 * i + 1];
            const int f0 = py.py_randbelow(num_mel - f);                // drawn before the skip test (:64)
            int lo = 0, hi = 0;
            if (f != 0) { lo = std::min(f0, num_mel); hi = std::min(f0 + w, num_mel); if (hi <= lo) lo = hi = 0; }
            mk[2 * i] = lo; mk[2 * i + 1] = hi;
        }
        // time masks (specaugment.py:90-105)
        int ts[2 * B200FE_MAX_TIME_MASKS];
        for (int i = 0; i < 2 * n_time_mask; ++i) ts[i] = np_.np_randint(max_time_width);
        for (int i = 0; i < n_time_mask; ++i) {
            const int t = ts[2 * i], w = ts[2 * i + 1];
            int lo = 0, hi = 0;
            if (T - t > 0) {                                            // else: skipped without drawing (:93-94)
                const long long t0 = py.py_randbelow((int)(T - t));
                if (t != 0) {
                    lo = (int)std::min<long long>(t0, T); hi = (int)std::min<long long>(t0 + w, T);
                    if (hi <= lo) lo = hi = 0;
                }
            }
            mk[2 * (n_freq_mask + i)] = lo; mk[2 * (n_freq_mask + i) + 1] = hi;
            rb[2 * i] = lo; rb[2 * i + 1] = hi;
        }
        std::sort(rb, rb + 2 * n_time_mask);
    }
    *py_pos = py.pos; *np_pos = np_.pos;
    return B200FE_OK;
}

extern "C" int b200fe_src_mask(const b200fe_plan* plan, const long long* d_len, int len_is_samples, int batch, int max_frames,
                               int subsample, unsigned char* d_mask, long long* d_out_len, void* stream)
{
    if (!d_len || batch <= 0 || max_frames < 0) return fail(B200FE_EINVAL, "src_mask: bad argument");
    if (len_is_samples && !plan) return fail(B200FE_EINVAL, "src_mask: a plan is needed to turn sample counts into frame counts");
    if (subsample != 1 && subsample != 4) return fail(B200FE_EINVAL, "src_mask: subsample must be 1 (encoder input) or 4 (after Conv2dSubsampling)");
    const int Tout = subsample == 1 ? max_frames : (max_frames >= 7 ? ((max_frames - 1) / 2 - 1) / 2 : 0);
    if (Tout > 0 && !d_mask) return fail(B200FE_EINVAL, "src_mask: mask buffer missing");
    if (Tout == 0 && !d_out_len) return B200FE_OK;
    dim3 grid((unsigned)std::max(1, std::min(64, (Tout + 255) / 256)), (unsigned)batch);
    src_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_len, len_is_samples ? plan->win : 0, len_is_samples ? plan->shift : 1, Tout, subsample,
                                                            d_mask, d_out_len);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_copy_ragged(const void* src, const long long* d_src_off, void* dst, const long long* d_dst_off,
                                  const long long* d_nbytes, int batch, long long max_bytes, void* stream)
{
    if (!src || !dst || !d_src_off || !d_dst_off || !d_nbytes || batch < 0 || max_bytes < 0) return fail(B200FE_EINVAL, "copy_ragged: bad argument");
    if (batch == 0 || max_bytes == 0) return B200FE_OK;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long cpr = (max_bytes + kCopyChunk - 1) / kCopyChunk;
    if (cpr > 0x7fffffffLL) return fail(B200FE_EINVAL, "copy_ragged: row too long");
    const long long total = cpr * batch;
    // Few CTAs on purpose: PCIe needs ~100 kB in flight (16 CTAs x 32 kB), and SMs that wait on system-memory reads slow
    // co-resident compute kernels down (measured), so the copy is confined to a handful of SMs.
    static const int env_ctas = []() { const char* e = getenv("B200FE_COPY_CTAS"); return e ? atoi(e) : 0; }();
    const int want = env_ctas > 0 ? env_ctas : 16;
    const int grid = (int)std::max<long long>(1, std::min<long long>(total, (long long)want));
    (void)sms;
    ragged_copy_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(static_cast<const char*>(src), d_src_off, static_cast<char*>(dst), d_dst_off,
                                                                        d_nbytes, batch, (int)cpr);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_resample_poly(const float* d_in, const long long* d_in_off, const long long* d_n_in, int batch, float* d_out,
                                    const long long* d_out_off, const long long* d_n_out, long long max_n_out, const float* d_filter, int filter_len,
                                    int up, int down, int pre_remove, float scale, void* stream)
{
    if (!d_in || !d_in_off || !d_n_in || !d_out || !d_out_off || !d_n_out || !d_filter || batch <= 0 || batch > 65535 || filter_len <= 0 || up <= 0 || down <= 0 ||
        pre_remove < 0 || max_n_out < 0)
        return fail(B200FE_EINVAL, "resample_poly: bad argument");
    if (max_n_out == 0) return B200FE_OK;
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((max_n_out + 255) / 256, 4096)), (unsigned)batch);
    resample_poly_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, d_in_off, d_n_in, d_out, d_out_off, d_n_out, d_filter, filter_len, up, down, pre_remove, scale);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_avg_channels(const float* d_in, float* d_out, long long n, int channels, void* stream)
{
    if (!d_in || !d_out || n < 0 || channels <= 0) return fail(B200FE_EINVAL, "avg_channels: bad argument");
    if (n == 0) return B200FE_OK;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16));
    avg_channels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, channels);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_cast_bf16(const float* d_in, void* d_out, long long n, void* stream)
{
    if (!d_in || !d_out || n < 0) return fail(B200FE_EINVAL, "cast_bf16: bad argument");
    if (((reinterpret_cast<uintptr_t>(d_in) & 15) | (reinterpret_cast<uintptr_t>(d_out) & 7)) != 0) return fail(B200FE_EINVAL, "cast_bf16: buffers must be 16 / 8-byte aligned");
    if (n == 0) return B200FE_OK;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, 148 * 16));
    cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, static_cast<__nv_bfloat16*>(d_out), n);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_copy_ragged_bf16(const float* src, const long long* d_src_off, void* dst, const long long* d_dst_off,
                                       const long long* d_nbytes, int batch, long long max_bytes, void* stream)
{
    if (!src || !dst || !d_src_off || !d_dst_off || !d_nbytes || batch < 0 || max_bytes < 0) return fail(B200FE_EINVAL, "copy_ragged_bf16: bad argument");
    if (batch == 0 || max_bytes == 0) return B200FE_OK;
    const long long cpr = (max_bytes + kCopyChunk - 1) / kCopyChunk;
    if (cpr > 0x7fffffffLL) return fail(B200FE_EINVAL, "copy_ragged_bf16: row too long");
    const int grid = (int)std::max<long long>(1, std::min<long long>(cpr * batch, 16LL));
    ragged_cast_bf16_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const char*>(src), d_src_off, static_cast<char*>(dst), d_dst_off,
                                                                             d_nbytes, batch, (int)cpr);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_d2h_ragged(const float* d_feats, long long row_elems, long long utt_rows, const long long* h_rows, int batch,
                                 float* h_feats, void* stream)
{
    if (!d_feats || !h_rows || !h_feats || batch < 0 || row_elems <= 0 || utt_rows <= 0) return fail(B200FE_EINVAL, "d2h_ragged: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int u = 0; u < batch; ++u) {
        const long long r = h_rows[u] < utt_rows ? h_rows[u] : utt_rows;
        if (r <= 0) continue;
        const long long off = (long long)u * utt_rows * row_elems;
        CUDA_TRY(cudaMemcpyAsync(h_feats + off, d_feats + off, sizeof(float) * (size_t)(r * row_elems), cudaMemcpyDeviceToHost, st));
    }
    return B200FE_OK;
}

extern "C" int b200fe_fbank_fused(const b200fe_plan* p, const b200fe_fbank_args* g, void* stream)
{
    if (!p || !g) return fail(B200FE_EINVAL, "fbank_fused: null argument");
    if (g->struct_size != sizeof(b200fe_fbank_args))
        return fail(B200FE_EINVAL, "fbank_fused: struct_size %u != sizeof(b200fe_fbank_args) %zu (stale binding of include/b200fe.h?)", g->struct_size, sizeof(b200fe_fbank_args));
    if (!g->d_wav || !g->d_nsamp || g->batch <= 0 || g->wav_stride <= 0) return fail(B200FE_EINVAL, "fbank_fused: bad waveform arguments");
    if (!g->d_out && !g->d_stats) return fail(B200FE_EINVAL, "fbank_fused: neither an output nor a statistics buffer");
    if (g->batch > 65535) return fail(B200FE_EINVAL, "fbank_fused: at most 65535 utterances per call (split the batch)");
    if (g->max_frames <= 0) return fail(B200FE_EINVAL, "fbank_fused: max_frames must be positive");
    if ((g->d_cmvn_mean == nullptr) != (g->d_cmvn_istd == nullptr)) return fail(B200FE_EINVAL, "fbank_fused: cmvn mean and istd go together");
    if (g->d_cmvn_mean && g->cmvn_stride != 0 && g->cmvn_stride != p->nmel) return fail(B200FE_EINVAL, "fbank_fused: cmvn_stride must be 0 or num_mel_bins");
    if (g->n_freq_masks < 0 || g->n_freq_masks > B200FE_MAX_FREQ_MASKS || g->n_time_masks < 0 || g->n_time_masks > B200FE_MAX_TIME_MASKS)
        return fail(B200FE_EINVAL, "fbank_fused: too many masks");
    int n_cls = g->d_stats ? (g->n_row_classes > 0 ? g->n_row_classes : 1) : 1;
    if (n_cls > kMaxRowClasses) return fail(B200FE_EINVAL, "fbank_fused: too many row classes");

    FbankArgs a;
    memset(&a, 0, sizeof a);
    a.wav = g->d_wav; a.wav_stride = g->wav_stride; a.wav_offsets = g->d_wav_offsets; a.nsamp = g->d_nsamp; a.peak = g->d_peak; a.B = g->batch;
    a.out = g->d_out; a.out_len = g->d_out_len; a.Tmax = g->max_frames; a.nmel = p->nmel;
    a.out_offsets = g->d_out ? g->d_out_offsets : nullptr;
    a.win = p->win; a.shift = p->shift;
    a.remove_dc = p->o.remove_dc_offset; a.use_power = p->o.use_power; a.use_log = p->o.use_log_fbank;
    a.preemph = p->o.preemphasis_coefficient;
    a.log_floor = 1.1920928955078125e-07f;   // TA:21-22
    a.in_scale = (float)std::ldexp(1.0, p->o.audio_bit - 1);
    a.dither = p->o.dither; a.dither_seed = g->dither_seed; a.dither_noise = g->d_dither_noise;
    const bool i16 = g->wav_dtype == 1;
    if (g->wav_dtype != 0 && g->wav_dtype != 1) return fail(B200FE_EINVAL, "fbank_fused: wav_dtype must be 0 (float32) or 1 (int16)");
    if (i16 && p->nfft != 512) return fail(B200FE_EINVAL, "fbank_fused: int16 input is implemented for the 512-point family only");
    if (i16 && p->o.dither != 0.f) return fail(B200FE_EINVAL, "fbank_fused: dither with int16 input is not implemented");
    // int16 samples already carry the 2^(bits-1) scale
    a.window = (g->d_peak || i16) ? p->d_window_plain : p->d_window_scaled;
    a.twiddle = p->d_twiddle; a.split_tw = p->d_split_tw;
    a.cm_mean = g->d_cmvn_mean; a.cm_istd = g->d_cmvn_istd; a.cm_stride = g->cmvn_stride;
    a.masks = (g->n_freq_masks + g->n_time_masks) > 0 ? g->d_masks : nullptr;
    a.n_fmask = g->n_freq_masks; a.n_tmask = g->n_time_masks; a.mask_zero = g->mask_zero;
    a.stats = g->d_stats; a.stats_stride = g->stats_stride; a.row_bounds = g->d_row_bounds; a.n_cls = n_cls;
    const bool ws = p->use_ws != 0 && !g->d_out_offsets;        // the experimental kernel writes the padded layout only
    const int tile_ft = plan_tile_frames(p);
    a.tiles_per_utt = (g->max_frames + tile_ft - 1) / tile_ft;
    const bool compact = g->d_tile_table != nullptr;
    if (compact && (!g->d_work_counter || g->n_tiles < 0)) return fail(B200FE_EINVAL, "fbank_fused: a tile table needs n_tiles and d_work_counter");
    if (g->d_n_tiles && (!compact || ws)) return fail(B200FE_EINVAL, "fbank_fused: d_n_tiles needs a tile table (and the default kernel)");
    const long long ntiles = compact ? (long long)g->n_tiles : (long long)a.tiles_per_utt * g->batch;
    if (ntiles > 0x7fffffffLL) return fail(B200FE_EINVAL, "fbank_fused: too many tiles");
    a.ntiles = (int)ntiles;
    // packed input: the caller promises 4-sample aligned offsets through offsets_aligned
    a.use_tma = ((reinterpret_cast<uintptr_t>(g->d_wav) & 15) == 0 &&
                 (g->d_wav_offsets ? g->offsets_aligned != 0 : (g->wav_stride % (i16 ? 8 : 4)) == 0)) ? 1 : 0;
    a.tile_floats = p->tile_floats;
    memcpy(a.seg_start, p->seg_start, sizeof a.seg_start);
    memcpy(a.grp_begin, p->grp_begin, sizeof a.grp_begin);
    memcpy(a.w_updn, p->w_updn, sizeof a.w_updn);

    a.tile_table = reinterpret_cast<const int2*>(g->d_tile_table);
    a.work_counter = g->d_work_counter;
    a.ntiles_ptr = g->d_n_tiles;          // table built on the device: n_tiles is only the capacity bound for the grid size
    if (g->apply_cmvn_mode != 0) {
        // utterance CMVN inside the launch: apply tiles of b200fe_build_work_list_device, lean instantiation only
        if (g->apply_cmvn_mode == 3) {
            // SpecAugment mean fills inside the launch: completion tiles (apply_lag < 0), full-epilogue instantiation
            if (!(plan_has_lean(p) && !p->use_ws) || g->d_peak || i16 || g->uniform_frames || g->d_out_offsets || !g->d_out || (g->d_cmvn_mean && g->cmvn_stride != 0))
                return fail(B200FE_EINVAL, "fbank_fused: in-launch mean fills need the default option set, float32 input, the padded output layout and global or no CMVN");
            if (!a.masks || g->mask_zero || !g->d_utt_done || !g->d_n_tiles || !g->d_stats || g->stats_stride < (long long)(n_cls + 1) * p->nmel)
                return fail(B200FE_EINVAL, "fbank_fused: in-launch mean fills need d_masks (mask_zero = 0), d_utt_done, a device-built work list and per-utterance statistics");
            if ((reinterpret_cast<uintptr_t>(g->d_out) & 15) != 0) return fail(B200FE_EINVAL, "fbank_fused: in-launch mean fills need a 16-byte aligned output");
            a.apply_mode = 3; a.utt_done = g->d_utt_done; a.fills = g->d_fills;
        } else {
        if (g->apply_cmvn_mode != 1 && g->apply_cmvn_mode != 2) return fail(B200FE_EINVAL, "fbank_fused: apply_cmvn_mode must be 0, 1, 2 or 3");
        if (!(plan_has_lean(p) && !p->use_ws) || g->d_peak || i16 || g->uniform_frames || g->d_cmvn_mean || a.masks || g->d_out_offsets || !g->d_out)
            return fail(B200FE_EINVAL, "fbank_fused: in-launch utterance CMVN needs the default option set, float32 input and the padded output layout");
        if (!g->d_utt_done || !g->d_n_tiles || !g->d_stats || g->stats_stride < 2LL * p->nmel || n_cls != 1)
            return fail(B200FE_EINVAL, "fbank_fused: in-launch utterance CMVN needs d_utt_done, a device-built work list and per-utterance statistics");
        if ((g->d_utt_mean == nullptr) != (g->d_utt_istd == nullptr)) return fail(B200FE_EINVAL, "fbank_fused: d_utt_mean and d_utt_istd go together");
        if ((reinterpret_cast<uintptr_t>(g->d_out) & 15) != 0) return fail(B200FE_EINVAL, "fbank_fused: in-launch utterance CMVN needs a 16-byte aligned output");
        a.apply_mode = g->apply_cmvn_mode; a.utt_done = g->d_utt_done; a.utt_mean = g->d_utt_mean; a.utt_istd = g->d_utt_istd;
        }
    }
    // Lock-step streaming (every utterance yields exactly max_frames frames): tiles take several utterances, so a 4-frame
    // push fills a 32-frame tile with 8 streams instead of occupying one tile per stream.
    if (g->uniform_frames && plan_has_multi(p) && !i16 && !compact && !ws && !g->d_out_offsets && g->max_frames <= kFT / 2 && !g->d_peak && !g->d_stats && !a.masks &&
        !(g->d_cmvn_mean && g->cmvn_stride != 0) && p->o.dither == 0.f && (p->nfft != 256 || g->max_frames % 2 == 0)) {
        const int per_reg = p->nfft == 512 ? 32 : 16;
        const int span = ((g->max_frames - 1) * p->shift + per_reg * p->nload + 3) & ~3;
        const int upt = std::min(kFT / g->max_frames, p->multi_tile_floats / span);
        if (upt >= 2) {
            a.multi_fpu = g->max_frames; a.multi_upt = upt; a.multi_span = span;
            a.multi_ragged = g->uniform_frames == 2 ? 1 : 0;
            a.ntiles = (g->batch + upt - 1) / upt;
            a.tile_floats = p->multi_tile_floats;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (compact && !g->d_n_tiles) CUDA_TRY(cudaMemsetAsync(g->d_work_counter, 0, sizeof(int), st));   // the device builder resets it itself
    if (compact && g->tile_table_pads && ws) return fail(B200FE_EINVAL, "fbank_fused: padding tiles are not available with the experimental kernel");
    if (compact || ws) {
        if (g->d_out && !g->d_out_offsets && !(compact && g->tile_table_pads)) {   // packed output has no padding rows; padding tiles zero them in the fused launch
            const long long per_utt = (long long)g->max_frames * p->nmel;
            dim3 zg((unsigned)((per_utt + 8191) / 8192), (unsigned)g->batch);
            zero_pad_kernel<<<zg, 256, 0, st>>>(g->d_out, g->d_nsamp, g->max_frames, p->nmel, p->win, p->shift);
            CUDA_TRY(cudaGetLastError());
        }
        if (ntiles == 0) return B200FE_OK;
    }
    void* kargs[] = {(void*)&a};
#ifdef B200FE_WITH_WS
    if (ws) {
        // persistent, one CTA per SM: FFT warps / epilogue warps / producer warp hand tiles over through mbarrier rings
        const int wgrid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)p->num_sms));
        CUDA_TRY(cudaLaunchKernel(ws_kernel(g->d_peak != nullptr, i16), dim3(wgrid), dim3(kWsThreads), kargs, (size_t)p->ws_smem_bytes, st));
        return B200FE_OK;
    }
#endif
    const int grid = (int)std::max<long long>(1, std::min<long long>(a.ntiles, (long long)p->num_sms * p->ctas_per_sm));
    // the lean instantiation serves the default option set whenever the launch applies no CMVN, no masks and writes the padded layout
    const bool lean = plan_has_lean(p) && !g->d_peak && !i16 && a.multi_fpu == 0 && !a.cm_mean && !a.masks && !a.out_offsets && a.apply_mode != 3;
    CUDA_TRY(launch_pdl(plan_kernel(p, g->d_peak != nullptr, i16, a.multi_fpu > 0, lean, a.apply_mode != 0), dim3(grid), dim3(kThreads), kargs,
                        (size_t)(a.multi_fpu > 0 ? p->multi_smem_bytes : p->smem_bytes), st));
    return B200FE_OK;
}

// ---- streaming front end: per-stream carry on the device, push = append -> fused launch on the state rows -> advance ----
struct b200fe_stream {
    const b200fe_plan* plan;
    int n_streams, max_chunk, cap, device;
    float* d_state = nullptr;        // [n_streams][cap]
    int* d_fill = nullptr;           // [n_streams] samples held per stream
    long long* d_offsets = nullptr;  // [n_streams] scratch of a push: row offsets / sample counts handed to the fused launch
    long long* d_nsamp = nullptr;
    int* d_flags = nullptr;
};

extern "C" int b200fe_stream_create(const b200fe_plan* plan, int n_streams, int max_chunk, b200fe_stream** out)
{
    if (!plan || !out || n_streams <= 0 || n_streams > 65535 || max_chunk <= 0 || max_chunk > (1 << 20))
        return fail(B200FE_EINVAL, "stream_create: bad argument (1..65535 streams, 1..2^20 samples per push)");
    b200fe_stream* st = new b200fe_stream();
    st->plan = plan; st->n_streams = n_streams; st->max_chunk = max_chunk;
    st->cap = (plan->win - 1 + max_chunk + 7) & ~7;            // 16-byte aligned rows (TMA loader), also for int16-sized offsets
    cudaError_t e = cudaGetDevice(&st->device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_state, sizeof(float) * (size_t)n_streams * st->cap + 64);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_fill, sizeof(int) * n_streams);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_offsets, sizeof(long long) * n_streams);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_nsamp, sizeof(long long) * n_streams);
    if (e == cudaSuccess) e = cudaMalloc((void**)&st->d_flags, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(st->d_state, 0, sizeof(float) * (size_t)n_streams * st->cap + 64);
    if (e == cudaSuccess) e = cudaMemset(st->d_fill, 0, sizeof(int) * n_streams);
    if (e == cudaSuccess) e = cudaMemset(st->d_flags, 0, sizeof(int));
    if (e != cudaSuccess) { b200fe_stream_destroy(st); return fail(B200FE_ECUDA, "stream_create: %s", cudaGetErrorString(e)); }
    *out = st;
    return B200FE_OK;
}

extern "C" void b200fe_stream_destroy(b200fe_stream* st)
{
    if (!st) return;
    cudaFree(st->d_state); cudaFree(st->d_fill); cudaFree(st->d_offsets); cudaFree(st->d_nsamp); cudaFree(st->d_flags);
    delete st;
}

extern "C" int b200fe_stream_max_frames(const b200fe_stream* st)
{
    if (!st) return fail(B200FE_EINVAL, "stream_max_frames: null handle");
    const int n = st->plan->win - 1 + st->max_chunk;
    return n >= st->plan->win ? 1 + (n - st->plan->win) / st->plan->shift : 0;
}

extern "C" int b200fe_stream_reset(b200fe_stream* st, const int* d_ids, int n, void* stream)
{
    if (!st || n < 0 || n > st->n_streams) return fail(B200FE_EINVAL, "stream_reset: bad argument");
    if (n == 0) return B200FE_OK;
    stream_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(st->d_fill, st->n_streams, d_ids, n);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_stream_push(b200fe_stream* st, const int* d_ids, int n, const float* d_chunks, long long chunk_stride, const int* d_chunk_len,
                                  const float* d_cmvn_mean, const float* d_cmvn_istd, float* d_out, int max_out_frames, long long* d_out_frames, void* stream)
{
    if (!st || n <= 0 || n > st->n_streams || !d_chunks || !d_out || max_out_frames <= 0 || chunk_stride < 0)
        return fail(B200FE_EINVAL, "stream_push: bad argument");
    if ((d_cmvn_mean == nullptr) != (d_cmvn_istd == nullptr)) return fail(B200FE_EINVAL, "stream_push: cmvn mean and istd go together");
    const b200fe_plan* p = st->plan;
    cudaStream_t cs = (cudaStream_t)stream;
    stream_append_kernel<<<n, 256, 0, cs>>>(st->d_state, st->d_fill, st->cap, st->n_streams, d_ids, d_chunks, chunk_stride, d_chunk_len, st->max_chunk,
                                            p->win, p->shift, max_out_frames, st->d_offsets, st->d_nsamp, d_out_frames, st->d_flags);
    CUDA_TRY(cudaGetLastError());
    b200fe_fbank_args a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a;
    a.d_wav = st->d_state; a.wav_stride = st->cap; a.d_wav_offsets = st->d_offsets; a.offsets_aligned = 1;
    a.d_nsamp = st->d_nsamp; a.batch = n; a.d_out = d_out; a.max_frames = max_out_frames;
    a.d_cmvn_mean = d_cmvn_mean; a.d_cmvn_istd = d_cmvn_istd; a.cmvn_stride = 0;
    a.uniform_frames = 2;          // every stream yields AT MOST max_out_frames frames: several streams share one 32-frame tile
    const int rc = b200fe_fbank_fused(p, &a, stream);
    if (rc != B200FE_OK) return rc;
    stream_advance_kernel<<<n, 128, 0, cs>>>(st->d_state, st->d_fill, st->cap, st->n_streams, d_ids, st->d_nsamp, p->win, p->shift);
    CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

extern "C" int b200fe_stream_flags(b200fe_stream* st, int* h_flags, void* stream)
{
    if (!st || !h_flags) return fail(B200FE_EINVAL, "stream_flags: bad argument");
    CUDA_TRY(cudaMemcpyAsync(h_flags, st->d_flags, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return B200FE_OK;
}

extern "C" int b200fe_postpass(const b200fe_plan* p, const b200fe_post_args* g, void* stream)
{
    if (!p || !g) return fail(B200FE_EINVAL, "postpass: null argument");
    if (g->struct_size != sizeof(b200fe_post_args))
        return fail(B200FE_EINVAL, "postpass: struct_size %u != sizeof(b200fe_post_args) %zu (stale binding of include/b200fe.h?)", g->struct_size, sizeof(b200fe_post_args));
    if (!g->d_feats || !g->d_nsamp || g->batch <= 0 || g->max_frames <= 0) return fail(B200FE_EINVAL, "postpass: bad argument");
    const int nm = g->n_freq_masks + g->n_time_masks;
    const bool masks = g->d_masks != nullptr && nm > 0;
    if (!masks && g->cmvn_mode == 0) return B200FE_OK;
    if (g->batch > 65535) return fail(B200FE_EINVAL, "postpass: at most 65535 utterances per call (split the batch)");
    if (!g->d_stats || g->stats_stride <= 0) return fail(B200FE_EINVAL, "postpass: per-utterance statistics are required");
    if (g->cmvn_mode != 0 && (!g->d_cmvn_mean || !g->d_cmvn_istd)) return fail(B200FE_EINVAL, "postpass: cmvn workspace missing");
    if (masks && !g->d_fills) return fail(B200FE_EINVAL, "postpass: fill buffer missing");
    if (g->n_freq_masks > B200FE_MAX_FREQ_MASKS || g->n_time_masks > B200FE_MAX_TIME_MASKS) return fail(B200FE_EINVAL, "postpass: too many masks");
    PostArgs a;
    memset(&a, 0, sizeof a);
    a.feats = g->d_feats; a.nsamp = g->d_nsamp; a.B = g->batch; a.Tmax = g->max_frames; a.nmel = p->nmel;
    a.win = p->win; a.shift = p->shift;
    a.stats = g->d_stats; a.stats_stride = g->stats_stride; a.row_bounds = g->d_row_bounds;
    a.n_cls = g->n_row_classes > 0 ? g->n_row_classes : 1;
    if (a.n_cls > kMaxRowClasses) return fail(B200FE_EINVAL, "postpass: too many row classes");
    a.cmvn_mode = g->cmvn_mode; a.cm_mean = g->d_cmvn_mean; a.cm_istd = g->d_cmvn_istd;
    a.masks = masks ? g->d_masks : nullptr; a.n_fmask = masks ? g->n_freq_masks : 0; a.n_tmask = masks ? g->n_time_masks : 0;
    a.fills = g->d_fills;
    a.fill_zero = g->fill_zero;
    a.feat_offsets = g->d_feat_offsets;
    a.rows_per_cta = 64;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (p->nmel % 4 == 0) && p->nmel <= 256 * 4 && ((reinterpret_cast<uintptr_t>(g->d_feats) & 15) == 0) &&
                     (a.cmvn_mode == 0 || (((reinterpret_cast<uintptr_t>(g->d_cmvn_mean) | reinterpret_cast<uintptr_t>(g->d_cmvn_istd)) & 15) == 0));
    // utterance CMVN without SpecAugment fills: the post pass derives the vectors itself, one launch instead of two
    a.inline_finalize = (vec && !masks && a.cmvn_mode != 0 && a.n_cls == 1) ? 1 : 0;
    if (!a.inline_finalize) {
        finalize_kernel<<<g->batch, 128, 0, st>>>(a);          // (launching it with programmatic serialisation measured no gain)
        CUDA_TRY(cudaGetLastError());
    }
    a.rows_per_cta = vec ? 96 : 64;
    dim3 grid((unsigned)((g->max_frames + a.rows_per_cta - 1) / a.rows_per_cta), (unsigned)g->batch);
    if (vec) {
        void* pargs[] = {(void*)&a};
        // plain utterance CMVN has its own lean kernel (more resident CTAs: the pass is latency bound)
        const bool lean_cmvn = a.inline_finalize && p->nmel <= kMaxMel;
        CUDA_TRY(launch_pdl(lean_cmvn ? (const void*)postpass_cmvn_kernel : (const void*)postpass_vec_kernel, grid, dim3(256), pargs, 0, st));
    } else {
        postpass_kernel<<<grid, 256, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    return B200FE_OK;
}

extern "C" int b200fe_time_warp(const b200fe_plan* p, const b200fe_warp_args* g, void* stream)
{
    if (!p || !g) return fail(B200FE_EINVAL, "time_warp: null argument");
    if (g->struct_size != sizeof(b200fe_warp_args))
        return fail(B200FE_EINVAL, "time_warp: struct_size %u != sizeof(b200fe_warp_args) %zu (stale binding of include/b200fe.h?)", g->struct_size, sizeof(b200fe_warp_args));
    if (!g->d_in || !g->d_out || !g->d_nsamp || !g->d_warp || g->batch <= 0 || g->max_frames <= 0)
        return fail(B200FE_EINVAL, "time_warp: bad argument");
    if (g->d_in == g->d_out) return fail(B200FE_EINVAL, "time_warp: in-place operation is not possible");
    WarpArgs a;
    memset(&a, 0, sizeof a);
    a.in = g->d_in; a.out = g->d_out; a.nsamp = g->d_nsamp; a.B = g->batch; a.Tmax = g->max_frames; a.nmel = p->nmel;
    a.win = p->win; a.shift = p->shift; a.warp = g->d_warp;
    a.stats = g->d_stats; a.stats_stride = g->stats_stride; a.row_bounds = g->d_row_bounds;
    a.n_cls = g->d_stats ? (g->n_row_classes > 0 ? g->n_row_classes : 1) : 1;
    if (a.n_cls > kMaxRowClasses) return fail(B200FE_EINVAL, "time_warp: too many row classes");
    if (g->d_masks) {
        if (!g->d_stats || !g->d_fills || !g->d_utt_done) return fail(B200FE_EINVAL, "time_warp: masks need d_stats, d_fills and d_utt_done");
        if (g->n_freq_masks < 0 || g->n_freq_masks > kMaxFreqMasks || g->n_time_masks < 0 || g->n_time_masks > kMaxTimeMasks)
            return fail(B200FE_EINVAL, "time_warp: at most %d frequency and %d time masks", kMaxFreqMasks, kMaxTimeMasks);
        a.masks = g->d_masks; a.n_fmask = g->n_freq_masks; a.n_tmask = g->n_time_masks; a.fills = g->d_fills; a.fill_zero = g->fill_zero; a.done = g->d_utt_done;
    }
    dim3 grid((unsigned)((g->max_frames + kWarpRows - 1) / kWarpRows), (unsigned)g->batch);
    // dynamic shared memory (static tables + this stay below 48 kB, no opt-in): the CTA's source rows when staging is compiled in
    // and enough of them fit, else just the statistics partials; with neither, taps come from global memory and the statistics
    // take the generic path
    const size_t budget = 48 * 1024 - sizeof(double) * kWarpRows * (kWarpTaps + 3) - sizeof(int) * kWarpRows * 4 - 2 * kWarpRows * kWarpTaps - 2048;
    size_t smem = 0;
    if ((p->nmel & 3) == 0) {
        const size_t rowb = (size_t)p->nmel * sizeof(float);
        const size_t part = (size_t)2 * std::min(1024 / p->nmel, kWarpRows) * rowb;      // (row slots that own a row) x columns x 2 moments
        const int fit = (int)std::min<size_t>(budget / rowb, (size_t)(kWarpRows + kWarpWinExtra));
        if (kWarpStage && fit >= kWarpRows + 12 && (size_t)fit * rowb >= part) { a.win_rows = fit; a.part_ok = 1; smem = (size_t)fit * rowb; }
        else if (part <= budget) { a.part_ok = 1; smem = part; }
    }
    void* kargs[] = {(void*)&a};
    static const bool use_pdl = []() { const char* v = getenv("B200FE_WARP_PDL"); return v ? atoi(v) != 0 : true; }();
    cudaError_t e;
    if (use_pdl) e = launch_pdl((const void*)time_warp_kernel, grid, dim3(256), kargs, smem, (cudaStream_t)stream);
    else { time_warp_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a); e = cudaGetLastError(); }
    if (e != cudaSuccess)
        return fail(B200FE_ECUDA, "time_warp: launch failed: %s (grid %u x %u, %zu B dynamic shared memory, %d staged rows)", cudaGetErrorString(e), grid.x, grid.y, smem, a.win_rows);
    return B200FE_OK;
}

extern "C" int b200fe_cmvn_from_stats(const double* stats, int nmel, int norm_vars, float* mean, float* istd)
{
    if (!stats || !mean || !istd || nmel <= 0) return fail(B200FE_EINVAL, "cmvn_from_stats: bad argument");
    const double n = stats[nmel];
    if (!(n > 0)) return fail(B200FE_EINVAL, "cmvn_from_stats: zero frame count");
    for (int d = 0; d < nmel; ++d) {
        const double m = stats[d] / n;
        double var = stats[(nmel + 1) + d] / n - m * m;
        if (var < 1e-20) var = 1e-20;
        mean[d] = (float)m;
        istd[d] = norm_vars ? (float)(1.0 / std::sqrt(var)) : 1.0f;
    }
    return B200FE_OK;
}

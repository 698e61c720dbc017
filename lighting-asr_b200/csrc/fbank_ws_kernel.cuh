// Warp-specialised fused Kaldi-fbank kernel for sm_100a (B200): the LASR default option set
// (16 kHz, 25 ms / 10 ms, 512-point FFT, 80 mel bins; lasr/data/datatrans.py:45-70).
//
// Same arithmetic as fbank_fused_kernel (fbank_kernel.cuh; TA:154-217, TA:616-633) -- the phase-A
// building blocks are shared -- but the three phases no longer take turns behind CTA-wide barriers:
//   * 12 FFT warps (24 half-warps, one frame each per tile) run phase A back to back: wait for the
//     waveform tile, load + window + FFT + real split + |X|^2, publish the power spectra of the
//     tile's 24 frames in a ring of PT buffers;
//   * 4 epilogue warps (one per SM sub-partition) consume PT buffers: mel projection (warp = group
//     of mel bins, lane = frame), log / CMVN, zero masks, coalesced copy-out, CMVN statistics;
//   * lane 0 of the first epilogue warp is also the producer: it resolves tile descriptors one tile ahead
//     (static round-robin of the compact tile list over the CTAs) and issues the 1-D TMA bulk copies of the
//     waveform tiles into a ring of tile buffers.
// All hand-offs are mbarrier full/empty pairs (TMA transaction counts for the waveform ring), so
// an FFT warp never waits for the other FFT warps and the FMA pipe stays busy while the epilogue
// of earlier tiles runs on the fourth warp of every scheduler.
// One persistent CTA per SM, 512 threads (128 registers each), ~215 kB of shared memory.
#pragma once
#include "fbank_kernel.cuh"

namespace b200fe {

#ifndef B200FE_WS_EPI
#define B200FE_WS_EPI 4             // epilogue warps (mel groups): 4 or 3
#endif
#ifndef B200FE_WS_STAGGER_NS
#define B200FE_WS_STAGGER_NS 0      // initial delay between the FFT warps that share a scheduler (de-phases FMA and shared-memory bursts)
#endif
constexpr int kWsFftWarps = 12;
constexpr int kWsEpiWarps = B200FE_WS_EPI;
constexpr int kWsThreads = 32 * (kWsFftWarps + kWsEpiWarps);
constexpr int kWsEpiThreads = 32 * kWsEpiWarps;
constexpr int kWsFT = 2 * kWsFftWarps;        // frames per tile
constexpr int kWsPTStride = kWsFT + 1;        // float4 groups per row of PT4[k/4][frame]; 4*(kWsFT+1) = 4 mod 32 words
constexpr int kWsNW = 4;                      // waveform tile ring
constexpr int kWsNP = 3;                      // PT ring
constexpr int kWsND = 16;                     // descriptor ring (>= kWsNW + kWsNP + 2)
constexpr int kWsNmel = 80;
constexpr int kWsOStride = kWsNmel + 1;

struct WsLayout {
    int wave_off, wave_bytes, xbuf_off, pt_off, pt_bytes, outs_off, outs_bytes, misc_off, bar_off, desc_off, total;
};

__host__ __device__ inline WsLayout ws_layout(int shift, int win)
{
    WsLayout L;
    int o = 0;
    // + 32 floats: the last lanes of a frame read (masked) samples past the window
    L.wave_bytes = (((kWsFT - 1) * shift + win + 32) * 4 + 127) & ~127;
    L.wave_off = o; o += kWsNW * L.wave_bytes;
    L.xbuf_off = o; o += 2 * kWsFftWarps * 16 * kXRow * 8;
    L.pt_bytes = 64 * kWsPTStride * 16;
    L.pt_off = o; o += kWsNP * L.pt_bytes;
    L.outs_bytes = (kWsFT * kWsOStride * 4 + 15) & ~15;
    L.outs_off = o; o += 2 * L.outs_bytes;
    // mean[80] | istd[80] | cmask[2][8] | split twiddles (128 float2) | window pairs (256 float2)
    L.misc_off = o; o += (2 * kWsNmel + 16) * 4 + (128 + 256) * 8;
    L.bar_off = o; o += 8 * (2 * kWsNW + 2 * kWsNP);
    o = (o + 15) & ~15;
    L.desc_off = o; o += 16 * kWsND;
    L.total = o;
    return L;
}

struct WsEmit {
    float* orow;
    const float* s_mean;
    const float* s_istd;
    float lf;
    bool lg, affine;
};

// The mel energies of a group stay in registers until the whole group is done: no shared-memory store sits between
// the PT loads, so the compiler is free to hoist them (the log / CMVN / store burst follows with full ILP).
#define B200FE_MEL_WS_CODE
#define MGROUP_BEGIN(w, jb, je) __device__ __forceinline__ void mel_ws_group##w(const float4* __restrict__ pcol, const WsEmit& em) { \
        constexpr int kJb = (jb), kJe = (je); float vals[kJe - kJb]; \
        float au = 0.f, ad = 0.f, au1 = 0.f, ad1 = 0.f, up_prev = 0.f; float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f); int cur = -1; (void)p4; (void)cur;
#define MK(k, wu, wd) { if (((k) >> 2) != cur) { cur = (k) >> 2; p4 = pcol[cur * kWsPTStride]; } \
        const float p = ((k) & 3) == 0 ? p4.x : ((k) & 3) == 1 ? p4.y : ((k) & 3) == 2 ? p4.z : p4.w; \
        if ((k) & 1) { if ((wu) != 0.f) au1 = fmaf((wu), p, au1); if ((wd) != 0.f) ad1 = fmaf((wd), p, ad1); } \
        else         { if ((wu) != 0.f) au = fmaf((wu), p, au);   if ((wd) != 0.f) ad = fmaf((wd), p, ad); } }
#define MEND0() { up_prev = au + au1; au = 0.f; ad = 0.f; au1 = 0.f; ad1 = 0.f; }
#define MEND(j) { vals[(j) - kJb] = up_prev + (ad + ad1); up_prev = au + au1; au = 0.f; ad = 0.f; au1 = 0.f; ad1 = 0.f; }
#define MGROUP_END(w) \
        _Pragma("unroll") for (int i = 0; i < kJe - kJb; ++i) { \
            float x = vals[i]; \
            if (em.lg) x = fast_log(fmaxf(x, em.lf)); \
            if (em.affine) x = (x - em.s_mean[kJb + i]) * em.s_istd[kJb + i]; \
            em.orow[kJb + i] = x; } }
#include "mel_static_default.inc"
#undef MGROUP_BEGIN
#undef MK
#undef MEND0
#undef MEND
#undef MGROUP_END
#undef B200FE_MEL_WS_CODE

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void ws_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kWsEpiThreads) : "memory"); }

template <bool kPeak, bool kI16>
__global__ void __launch_bounds__(kWsThreads, 1) fbank_ws_kernel(const __grid_constant__ FbankArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const WsLayout L = ws_layout(a.shift, a.win);
    float* s_mean = reinterpret_cast<float*>(smem + L.misc_off);
    float* s_istd = s_mean + kWsNmel;
    unsigned* s_cmask = reinterpret_cast<unsigned*>(s_istd + kWsNmel);     // [2][8]: [0..3] column bits, [4] row bits
    float2* s_stw = reinterpret_cast<float2*>(s_cmask + 16);
    float2* s_win = s_stw + 128;
    uint64_t* wave_full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* wave_empty = wave_full + kWsNW;
    uint64_t* pt_full = wave_empty + kWsNW;
    uint64_t* pt_empty = pt_full + kWsNP;
    volatile int4* s_desc = reinterpret_cast<volatile int4*>(smem + L.desc_off);   // (utt | -1 = end, f0, nvalid, T)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    for (int k = tid; k < 256; k += kWsThreads) s_win[k] = make_float2(__ldg(a.window + 2 * k), __ldg(a.window + 2 * k + 1));
    for (int k = tid; k < 128; k += kWsThreads) s_stw[k] = __ldg(a.split_tw + k);
    const bool cm_per_utt = a.cm_mean != nullptr && a.cm_stride != 0;
    if (tid < kWsNmel) {
        const bool on = a.cm_mean != nullptr && !cm_per_utt;
        s_mean[tid] = on ? __ldg(a.cm_mean + tid) : 0.f;
        s_istd[tid] = on ? __ldg(a.cm_istd + tid) : 1.f;
    }
    if (tid < 16) s_cmask[tid] = 0u;
    if (tid == 0) {
        for (int s = 0; s < kWsNW; ++s) { mbar_init(&wave_full[s], a.use_tma ? 1 : 32); mbar_init(&wave_empty[s], kWsFftWarps); }
        for (int s = 0; s < kWsNP; ++s) { mbar_init(&pt_full[s], kWsFftWarps); mbar_init(&pt_empty[s], kWsEpiWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp < kWsFftWarps) {
        // =========================== FFT warps: phase A ===========================
        const int h2 = lane >> 4, l = lane & 15;
        float2* xbuf = reinterpret_cast<float2*>(smem + L.xbuf_off) + (warp * 2 + h2) * 16 * kXRow;
        const float2* wl = s_win + l;
        const float2* stw = s_stw + l;
        float2 tw[16];
#pragma unroll
        for (int k = 1; k < 16; ++k) tw[k] = __ldg(a.twiddle + l * 16 + k);
        const int fl = (warp & 3) + 8 * (warp >> 2) + 4 * h2;     // half-warps of a warp sit 4 frames apart (disjoint PT banks)
        FrameCtx fc;
        fc.c_pre = a.preemph; fc.inv_win = 1.0f / (float)a.win;
        fc.dc_coef = a.remove_dc ? (float)(1.0 - (double)a.preemph) : 0.0f;
        fc.win = a.win; fc.pmax = 1.0f; fc.prcp = 0.0f; fc.pscale = 1.0f;
        fc.dither = 0.f; fc.seed = 0; fc.utt = 0; fc.ta = fc.tb = 0; fc.noise_a = fc.noise_b = nullptr;
        int ws = 0, ps = 0, dslot = 0;
        uint32_t wph = 0, pph = 0;
        if (B200FE_WS_STAGGER_NS > 0 && (warp >> 2) > 0) __nanosleep((warp >> 2) * B200FE_WS_STAGGER_NS);
#pragma unroll 1
        for (;;) {
            mbar_wait(&wave_full[ws], wph);
            const int4 d = make_int4(s_desc[dslot].x, s_desc[dslot].y, s_desc[dslot].z, s_desc[dslot].w);
            const int utt = d.x, nvalid = d.x < 0 ? 0 : d.z;
            const bool any = fl - 4 * h2 < nvalid;        // warp-uniform
            const bool fvalid = fl < nvalid;
            float2 v[16];
            if (any) {
                if (kPeak) {
                    // reference: x / (max + 1e-9) in fp64, rounded to fp32, times 2^(bits-1) (datatrans.py:24-25,73-74)
                    fc.pmax = __ldg(a.peak + utt);
                    fc.prcp = (float)(1.0 / ((double)fc.pmax + 1e-9));
                    fc.pscale = a.in_scale;
                }
                const unsigned char* xs = smem + L.wave_off + ws * L.wave_bytes;
                if (kI16) load_frame_single_i16<13, kPeak>(v, reinterpret_cast<const short*>(xs) + fl * a.shift + 2 * l, wl, fc, l);
                else load_frame_single<13, kPeak, false>(v, reinterpret_cast<const float*>(xs) + fl * a.shift + 2 * l, wl, fc, l);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&wave_empty[ws]);     // the tile's samples are in registers
            float2 rc[8];
            if (any) {
                fft256_halfwarp(v, tw, xbuf, l);
                pair_exchange(v, rc, l, h2);
            }
            mbar_wait(&pt_empty[ps], pph ^ 1u);
            if (any) {
                float* pt = reinterpret_cast<float*>(smem + L.pt_off + ps * L.pt_bytes);
                float* pa = pt + ((l >> 2) * kWsPTStride + fl) * 4 + (l & 3);                          // k = l + 16 r
                float* pb = pt + (((256 - l) >> 2) * kWsPTStride + fl) * 4 + ((256 - l) & 3);          // k = 256 - l - 16 r
                // ---- real-FFT split + power: 2X[k] = S + T, 2 conj X[256-k] = S - T ----
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 bcj = make_float2(rc[r].x, -rc[r].y);
                    const float2 S = add2(v[r], bcj), D = sub2(v[r], bcj);
                    const float2 sw = stw[16 * r];
                    const float2 T = c_mul(D, sw.x, sw.y);
                    float2 xa = add2(S, T), xb = sub2(S, T);
                    xa = mul2(xa, xa); xb = mul2(xb, xb);
                    const float pwa = xa.x + xa.y, pwb = xb.x + xb.y;
                    if (fvalid) pa[r * 4 * kWsPTStride * 4] = pwa;
                    if (fvalid && (r != 0 || l != 0)) pb[-r * 4 * kWsPTStride * 4] = pwb;
                }
                if (l == 0 && fvalid) {   // bin 128 is its own partner: X[128] = conj Z[128]
                    const float2 z = v[8];
                    pt[(32 * kWsPTStride + fl) * 4] = 4.0f * (z.x * z.x + z.y * z.y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pt_full[ps]);
            if (d.x < 0) break;
            if (++ws == kWsNW) { ws = 0; wph ^= 1u; }
            if (++ps == kWsNP) { ps = 0; pph ^= 1u; }
            dslot = (dslot + 1) & (kWsND - 1);
        }
    } else {
        // =========================== epilogue warps: phases B and C ===========================
        const int ew = warp - kWsFftWarps;
        const int et = tid - 32 * kWsFftWarps;
        constexpr int grp[5] = B200FE_WS_GRP_BEGIN;
        const int jb = ew == 0 ? grp[0] : ew == 1 ? grp[1] : ew == 2 ? grp[2] : grp[3];
        const int je = ew == 0 ? grp[1] : ew == 1 ? grp[2] : ew == 2 ? grp[3] : grp[4];
        static_assert(kWsNmel <= kWsEpiThreads, "one epilogue thread per mel bin");
        const bool zmask = a.mask_zero && a.masks != nullptr;
        const int nmask = a.n_fmask + a.n_tmask;
        const bool lg = a.use_log != 0;
        const bool affine = a.cm_mean != nullptr;
        const float lf = a.log_floor;
        const bool want_stats = a.stats != nullptr;
        int ps = 0, dslot = 0, sbuf = 0;
        uint32_t pph = 0;
        // ---- producer state (lane 0 of epilogue warp 0): tiles of this CTA are id = blockIdx.x + k * gridDim.x ----
        const bool producer = (ew == 0);       // warp-collective; lane 0 owns the barrier operations
        const bool tab = a.tile_table != nullptr;
        const int esz = kI16 ? 2 : 4;
        int p_k = 0;                 // next candidate (k-th tile of this CTA)
        int p_ws = 0, p_dslot = 0;   // next waveform slot / descriptor slot
        uint32_t p_wph = 0;
        bool p_done = false;
        int2 p_e = make_int2(-1, 0);         // table entry of candidate p_k (prefetched)
        long long p_n = 0;                   // its sample count (prefetched)
        int2 p_e2 = make_int2(-1, 0);        // entry of candidate p_k + 1 (requested one step ahead)
        auto fetch_entry = [&](int k) -> int2 {
            const long long id = (long long)blockIdx.x + (long long)k * gridDim.x;
            if (id >= a.ntiles) return make_int2(-1, 0);
            if (tab) return __ldg(a.tile_table + id);
            const int u = (int)(id / a.tiles_per_utt);
            return make_int2(u, (int)(id - (long long)u * a.tiles_per_utt) * kWsFT);
        };
        if (producer) {
            p_e = fetch_entry(0);
            if (p_e.x >= 0) p_n = __ldg(a.nsamp + p_e.x);
            p_e2 = fetch_entry(1);
        }
        auto produce = [&]() {
            // issue as many tiles as the waveform ring takes right now (never blocks)
            while (!p_done) {
                int ready = 0;
                if (lane == 0) ready = mbar_test(&wave_empty[p_ws], p_wph ^ 1u) ? 1 : 0;
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (!ready) break;
                const int utt = p_e.x, f0 = p_e.y;
                int T = 0, nvalid = 0;
                if (utt >= 0) {
                    const unsigned n = (unsigned)p_n;                  // < 2^31 samples per utterance
                    T = n >= (unsigned)a.win ? (int)(1u + (n - (unsigned)a.win) / (unsigned)a.shift) : 0;
                    nvalid = min(max(T - f0, 0), kWsFT);
                }
                // advance the prefetch pipeline: entry k+1 becomes current (its sample count is requested now),
                // entry k+2 is requested
                ++p_k;
                p_e = p_e2;
                if (p_e.x >= 0) p_n = __ldg(a.nsamp + p_e.x);
                p_e2 = fetch_entry(p_k + 1);
                if (utt >= 0 && nvalid <= 0) continue;       // padded grid: nothing to compute (rows are zeroed by zero_pad_kernel)
                if (lane == 0) { s_desc[p_dslot].x = utt; s_desc[p_dslot].y = f0; s_desc[p_dslot].z = nvalid; s_desc[p_dslot].w = T; }
                __syncwarp();
                if (utt < 0) {                                   // end marker
                    if (!a.use_tma || lane == 0) mbar_arrive(&wave_full[p_ws]);
                    p_done = true;
                    break;
                }
                const int nsmp = (nvalid - 1) * a.shift + a.win;
                const long long eoff = (a.wav_offsets ? __ldg(a.wav_offsets + utt) : (long long)utt * a.wav_stride) + (long long)f0 * a.shift;
                unsigned char* dst = smem + L.wave_off + p_ws * L.wave_bytes;
                if (a.use_tma) {
                    if (lane == 0) {
                        const uint32_t bytes = (uint32_t)((nsmp * esz + 15) & ~15);
                        mbar_expect_tx(&wave_full[p_ws], bytes);
                        tma_load_1d(dst, reinterpret_cast<const char*>(a.wav) + eoff * esz, bytes, &wave_full[p_ws]);
                    }
                } else if (kI16) {
                    // unaligned int16 input: plain 2-byte copies by this warp (slow path)
                    const short* src = reinterpret_cast<const short*>(a.wav) + eoff;
                    short* xd = reinterpret_cast<short*>(dst);
                    for (int i = lane; i < nsmp; i += 32) xd[i] = __ldg(src + i);
                    mbar_arrive(&wave_full[p_ws]);
                } else {
                    // unaligned float input: 4-byte asynchronous copies; every lane's arrival fires when its copies have landed
                    const float* src = a.wav + eoff;
                    const uint32_t d0 = smem_u32(dst);
                    for (int i = lane; i < nsmp; i += 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * (uint32_t)i), "l"(src + i) : "memory");
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&wave_full[p_ws])) : "memory");
                }
                if (++p_ws == kWsNW) { p_ws = 0; p_wph ^= 1u; }
                p_dslot = (p_dslot + 1) & (kWsND - 1);
            }
        };
#pragma unroll 1
        for (;;) {
            if (producer) produce();
            __syncwarp();
            mbar_wait(&pt_full[ps], pph);
            const int4 d = make_int4(s_desc[dslot].x, s_desc[dslot].y, s_desc[dslot].z, s_desc[dslot].w);
            if (d.x < 0) break;
            const int utt = d.x, f0 = d.y, nvalid = d.z;
            float* outs = reinterpret_cast<float*>(smem + L.outs_off + sbuf * L.outs_bytes);
            unsigned* cmask = s_cmask + 8 * sbuf;
            // per-utterance CMVN vectors: every warp stages the bins of its own group (read back by this warp only)
            if (cm_per_utt) {
                __syncwarp();
                for (int j = jb + lane; j < je; j += 32) {
                    s_mean[j] = __ldg(a.cm_mean + (long long)utt * a.cm_stride + j);
                    s_istd[j] = __ldg(a.cm_istd + (long long)utt * a.cm_stride + j);
                }
                __syncwarp();
            }
            if (zmask) {
                const int* mk = a.masks + (long long)utt * nmask * 2;
                const int col = et;
                bool m = false;
#pragma unroll 1
                for (int i = 0; i < a.n_fmask; ++i) m |= (col >= __ldg(mk + 2 * i) && col < __ldg(mk + 2 * i + 1));
                const unsigned bal = __ballot_sync(0xffffffffu, m);
                if (lane == 0) cmask[ew] = bal;
                if (ew == 0) {
                    const int* mt = mk + 2 * a.n_fmask;
                    bool mr = false;
#pragma unroll 1
                    for (int i = 0; i < a.n_tmask; ++i) mr |= (f0 + lane >= __ldg(mt + 2 * i) && f0 + lane < __ldg(mt + 2 * i + 1));
                    const unsigned balr = __ballot_sync(0xffffffffu, mr);
                    if (lane == 0) cmask[4] = balr;
                }
            }
            // ---- phase B: warp = mel-bin group, lane = frame ----
            if (lane < kWsFT) {
                const float4* pcol = reinterpret_cast<const float4*>(smem + L.pt_off + ps * L.pt_bytes) + lane;
                WsEmit em;
                em.orow = outs + lane * kWsOStride; em.s_mean = s_mean; em.s_istd = s_istd; em.lf = lf; em.lg = lg; em.affine = affine;
                switch (ew) {
                    case 0: mel_ws_group0(pcol, em); break;
                    case 1: mel_ws_group1(pcol, em); break;
#if B200FE_WS_EPI == 4
                    case 2: mel_ws_group2(pcol, em); break;
                    default: mel_ws_group3(pcol, em); break;
#else
                    default: mel_ws_group2(pcol, em); break;
#endif
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pt_empty[ps]);
            ws_epi_bar();       // the staging tile is complete
            // ---- phase C: masks, coalesced copy-out, statistics ----
            float* obase = a.out != nullptr ? a.out + ((long long)utt * a.Tmax + f0) * kWsNmel : nullptr;
            const int nv = nvalid * kWsNmel;
            if (!zmask) {
                if (obase != nullptr) {
                    if (nvalid == kWsFT) {
                        constexpr int kPer = kWsFT * kWsNmel / kWsEpiThreads;
                        float x[kPer];
#pragma unroll
                        for (int i = 0; i < kPer; ++i) { const int e = et + i * kWsEpiThreads; x[i] = outs[e + e / kWsNmel]; }
#pragma unroll
                        for (int i = 0; i < kPer; ++i) obase[et + i * kWsEpiThreads] = x[i];
                    } else {
#pragma unroll 2
                        for (int e = et; e < nv; e += kWsEpiThreads) obase[e] = outs[e + e / kWsNmel];
                    }
                }
            } else {
                const unsigned rmask = cmask[4];
#pragma unroll 1
                for (int e = et; e < nv; e += kWsEpiThreads) {
                    const int row = e / kWsNmel, col = e - row * kWsNmel;
                    float* sp = outs + e + row;
                    float x = *sp;
                    if (((rmask >> row) & 1u) || ((cmask[col >> 5] >> (col & 31)) & 1u)) x = 0.f;
                    if (obase) obase[e] = x;
                    if (want_stats) *sp = x;
                }
                if (want_stats) ws_epi_bar();
            }
            if (a.out_len != nullptr && f0 == 0 && et == 0) a.out_len[utt] = d.w;
            if (want_stats && et < kWsNmel) {
                // Column statistics of the staged tile: thread = column; fp32 sums about a pivot over <= 24 rows,
                // flushed with fp64 atomics (sum = s1 + n p, sumsq = s2 + 2 p s1 + n p^2).
                const int j = et;
                const int nb = a.n_cls - 1;
                const int* bounds = (a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
                double* sb = a.stats + (long long)utt * a.stats_stride;
                if (bounds == nullptr) {
                    const float pivot = outs[j];
                    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
                    int fr = 0;
                    for (; fr + 1 < nvalid; fr += 2) {
                        const float da = outs[fr * kWsOStride + j] - pivot, db = outs[(fr + 1) * kWsOStride + j] - pivot;
                        s1a += da; s2a = fmaf(da, da, s2a);
                        s1b += db; s2b = fmaf(db, db, s2b);
                    }
                    if (fr < nvalid) { const float da = outs[fr * kWsOStride + j] - pivot; s1a += da; s2a = fmaf(da, da, s2a); }
                    const double dp = (double)pivot, d1 = (double)s1a + (double)s1b, d2 = (double)s2a + (double)s2b;
                    atomicAdd(sb + j, d1 + nvalid * dp);
                    atomicAdd(sb + (long long)a.n_cls * kWsNmel + j, d2 + 2.0 * dp * d1 + nvalid * dp * dp);
                } else {
                    int cls = row_class(bounds, nb, f0);
                    float s1 = 0.f, s2 = 0.f, pivot = 0.f;
                    int cnt = 0;
                    double q2 = 0.0;
                    for (int fr = 0; fr < nvalid; ++fr) {
                        const int cc = row_class(bounds, nb, f0 + fr);
                        if (cc != cls) {
                            const double dp = (double)pivot, d1 = (double)s1;
                            atomicAdd(sb + (long long)cls * kWsNmel + j, d1 + cnt * dp);
                            q2 += (double)s2 + 2.0 * dp * d1 + cnt * dp * dp;
                            s1 = 0.f; s2 = 0.f; cnt = 0; cls = cc;
                        }
                        const float x = outs[fr * kWsOStride + j];
                        if (fr == 0) pivot = x;
                        const float dd = x - pivot;
                        s1 += dd;
                        s2 = fmaf(dd, dd, s2);
                        ++cnt;
                    }
                    const double dp = (double)pivot, d1 = (double)s1;
                    atomicAdd(sb + (long long)cls * kWsNmel + j, d1 + cnt * dp);
                    atomicAdd(sb + (long long)a.n_cls * kWsNmel + j, q2 + (double)s2 + 2.0 * dp * d1 + cnt * dp * dp);
                }
            }
            sbuf ^= 1;
            if (++ps == kWsNP) { ps = 0; pph ^= 1u; }
            dslot = (dslot + 1) & (kWsND - 1);
        }
    }
}

}  // namespace b200fe

// Fused Kaldi-fbank kernel for sm_100a (B200).
//
// Replaces, for a whole padded batch in one launch, the per-utterance CPU chain
//   lasr/data/datatrans.py:42-104 (WavToKaldiFbank) -> torchaudio/compliance/kaldi.py:514-645
//   (framing TA:44-83, DC removal / pre-emphasis / window / zero-pad TA:154-217, rfft+power
//   TA:616-618, mel projection TA:621-630, floored log TA:631-633) and the zero-padded collate
//   lasr/data/dataset.py:8-22, plus the CMVN / SpecAugment epilogue of SURVEY.md section 8.
//
// Work decomposition:
//   * grid  : persistent CTAs striding over tiles; a tile = FT(32) consecutive frames of one
//             utterance.  256 threads = 16 half-warps.
//   * load  : the waveform span of a tile ((FT-1)*shift + win samples, each sample fetched once
//             per tile) is brought to shared memory with one 1-D TMA bulk copy (UBLKCP) per
//             tile into ONE tile buffer (kStages = 1: the next tile's copy is issued right after the
//             phase-A barrier), completion on an mbarrier.
//   * phase A (half-warp per frame, 2 frames each): 16 lanes x 16 complex registers hold the
//             frame packed as a 256-point complex sequence z[n] = y[2n] + j y[2n+1].
//             DC removal, pre-emphasis and the window are applied while loading; DFT-16 in
//             registers (packed FADD2/FFMA2), twiddle, one transposition through shared memory,
//             second DFT-16, conjugate-pair exchange with warp shuffles, real-FFT split (split twiddles = the
//             lane's r = 0 value x W_32^r immediates) and |X|^2 -> the frame's power row INSIDE the half-warp's
//             transposition region (see kRegionHW below).
//   * phase B (warp = group of mel bins, lane = frame): sparse triangular mel accumulate.  For the
//             LASR default option set the projection is straight-line code with the weights as
//             immediates (mel_static_default.inc); otherwise the (up, down) weights are
//             warp-uniform loads from the constant bank (kernel parameters).  log / CMVN /
//             zero-masks are applied and the tile's features land in a staging buffer.
//   * phase C : coalesced copy-out (+ zero fill of padded rows) and, in statistics mode,
//             per-column sums / sums of squares per SpecAugment row class, flushed with one
//             fp64 atomic per column, class and tile.
#pragma once
#include "b200fe_common.cuh"

namespace b200fe {

#ifndef B200FE_WARPS
#define B200FE_WARPS 8              // warps per CTA: 8 (2 CTAs/SM, 32-frame tiles), 6 (3 CTAs/SM, 24-frame tiles) or 4 (4 CTAs/SM, 16-frame tiles)
#endif
// Two scheduling variants kept behind macros because they were measured SLOWER on B200 (A/B on one box, C2 plain launch:
// 0.3285 ms baseline): issuing the next tile's TMA as soon as every warp holds its last frames in registers instead of
// after the phase-A barrier (0.3422 ms: the transfer then competes with phase A for the shared-memory pipe), and routing
// the plain copy-out through the (row part, column) mapping of the statistics path (0.3400 ms: 240 of 256 threads, 11 rounds).
#ifndef B200FE_DESC_PIPE
#define B200FE_DESC_PIPE 1          // 0 (A/B runs): thread 0 resolves a tile descriptor in one blocking chain at the top of every tile, as in round 1
#endif
#ifndef B200FE_EARLY_TMA
#define B200FE_EARLY_TMA 0
#endif
#ifndef B200FE_XP_SHFL
#define B200FE_XP_SHFL 0
#endif

#ifndef B200FE_ROWPART_PLAIN
#define B200FE_ROWPART_PLAIN 0
#endif
constexpr bool kEarlyTma = B200FE_EARLY_TMA != 0;
constexpr int kWarps = B200FE_WARPS;
constexpr int kFT = 4 * kWarps;             // frames per tile: every half-warp transforms two frames
constexpr int kThreads = 32 * kWarps;
constexpr int kHalfWarps = 2 * kWarps;
#ifndef B200FE_CTAS
#define B200FE_CTAS (B200FE_WARPS == 8 ? 2 : B200FE_WARPS == 4 ? 4 : 3)
#endif
constexpr int kCtasPerSm = B200FE_CTAS;
constexpr bool kAliasStaging = false;      // (the 3 CTAs/SM variant used to alias the staging tile onto the transposition buffers)
constexpr int kMelGroups = kWarps == 4 ? 8 : kWarps;   // bin groups of phase B (16-frame tiles: one per half-warp)   // 3 CTAs/SM only fit with the staging tile aliased onto the transposition buffers
constexpr int kXRow = 17;           // padded row length (float2) of the transposition buffer
// Power spectra live INSIDE the transposition buffers (a frame's 256 power values replace the 2 kB the FFT no longer needs):
// every half-warp owns a region [transposition: 16 rows x 17 float2 = 544 words | PT row of its first frame: 256 words | pad],
// the PT row of its second frame overwrites the transposition area once the frame's FFT has left it.  Strides are chosen so
// that (a) the two half-warps of a warp, which store in the same instruction, hit disjoint banks (region stride = 16 mod 32
// words) and (b) the eight frames a quarter-warp reads with one LDS.128 in phase B (four consecutive warps x two half-warps)
// start 4 words apart modulo 32 (warp stride = 4 mod 32): both the writers and the readers are conflict free.
constexpr int kXWords = 16 * 17 * 2;                 // transposition area of a half-warp (float2 rows padded to 17)
constexpr int kRegionHW = kXWords + 256 + 16;        // 816 words = 16 mod 32
constexpr int kRegionWarp = 2 * kRegionHW + 4;       // 1636 words = 4 mod 32
static_assert(kRegionHW % 32 == 16 && kRegionWarp % 32 == 4, "bank layout of the power-spectrum rows");
#ifndef B200FE_TABLES_L1
#define B200FE_TABLES_L1 0          // 1: window / split twiddles are read through L1 (ld.global.nc) instead of shared-memory copies
#endif
constexpr bool kTablesL1 = B200FE_TABLES_L1 != 0;
#ifndef B200FE_STW_ROT
#define B200FE_STW_ROT 1            // 1 (default, measured -3.2 % on the C2 launch): split twiddles -j W_512^(l + 16 r) = (lane's r = 0 value, two registers) x W_32^r (immediates): no table loads
#endif
constexpr bool kStwRot = B200FE_STW_ROT != 0;
constexpr int kMaxMel = 128;
constexpr int kMaxTimeMasks = 4;
constexpr int kMaxFreqMasks = 4;
constexpr int kMaxRowClasses = 2 * kMaxTimeMasks + 1;
constexpr int kPadTileRows = 256;   // rows zeroed by one padding tile of the compact work list (table entry with first frame < 0)
constexpr int kStages = 1;         // one tile buffer: the next tile's TMA is issued right after the phase-A barrier
constexpr int kApplyBit = 0x40000000;   // work-list entry (utt | kApplyBit, row0): CMVN-apply tile, rows [row0, row0 + kApplyRows)
constexpr int kApplyRows = kWarps == 8 ? 192 : 96;         // rows of one apply tile: 192 * 80 * 4 B = 61.4 kB fit the transposition/PT regions + the staging tile
constexpr int kSigBatch = 8;            // completions a CTA collects before one fence publishes them
constexpr int kReadyBit = 0x20000000;   // descriptor only: the utterance was already complete when the tile was claimed (no wait)

struct FbankArgs {
    // input
    const float* wav;            // [B][wav_stride] fp32 waveform, or int16 PCM when the kernel is instantiated with kI16
    long long wav_stride;
    const long long* wav_offsets; // optional [B]: utterance u starts at wav + wav_offsets[u] (packed / ragged input)
    const long long* nsamp;      // [B] valid samples
    const float* peak;           // [B] per-utterance abs-max (peak normalisation) or nullptr
    int B;
    // output
    float* out;                  // [B][Tmax][nmel] (nullptr in statistics-only mode)
    const long long* out_offsets; // optional [B] first output row of every utterance: packed [sum T][nmel] output, no padding rows
    long long* out_len;          // [B] frames per utterance (optional)
    int Tmax;
    int nmel;
    // framing / fbank options
    int win, shift;              // samples
    int remove_dc, use_power, use_log;
    float preemph;
    float log_floor;             // FLT_EPSILON
    float in_scale;              // 2^(audio_bit-1), applied on load when peak != nullptr (else folded in window)
    // dither (TA:179-181): x[f][j] += dither * n[f][j], independent noise per (frame, sample).  n comes from
    // dither_noise [B][Tmax][win] when given (parity with a host generator), else from Philox4x32-10.
    float dither;
    unsigned long long dither_seed;
    const float* dither_noise;
    // plan tables
    const float* window;         // [512] window * (in_scale or 1), zero-extended
    const float2* twiddle;       // [16][16] W_256^(n1*klo)
    const float2* split_tw;      // [256]  -j * W_512^k
    // CMVN applied in the epilogue: (x - mean) * istd ; nullptr = off ; stride 0 = global, nmel = per utterance
    const float* cm_mean;
    const float* cm_istd;
    long long cm_stride;
    // SpecAugment rectangles, per utterance [n_fmask + n_tmask][2] = (start, stop) ; freq first
    const int* masks;
    int n_fmask, n_tmask;
    int mask_zero;               // 1: zero the masked cells here (replace_with_zero); 0: leave them for the fill kernel
    // statistics: stats[u*stats_stride + c*nmel + d], c < n_cls row-class sums, c == n_cls sums of squares
    double* stats;
    long long stats_stride;      // 0 = one global accumulator
    const int* row_bounds;       // per utterance [n_cls-1] sorted class boundaries (nullptr -> one class)
    int n_cls;
    // scheduling
    int tiles_per_utt;
    int ntiles;
    // optional compact work list of valid tiles (utt, first frame) consumed through an atomic counter
    // (dynamic scheduling; padded rows are then zeroed by zero_pad_kernel)
    const int2* tile_table;
    int* work_counter;
    const int* ntiles_ptr;        // optional: the number of tiles lives in device memory (work list built on the device)
    int use_tma;
    int tile_floats;             // floats reserved per tile stage
    // multi-utterance tiles (lock-step streaming: every utterance yields exactly Tmax = multi_fpu frames): a tile takes
    // multi_upt consecutive utterances, multi_fpu frames each; utterance j's samples sit at j * multi_span in the tile buffer
    int multi_fpu, multi_upt, multi_span;
    int multi_ragged;               // 1: an utterance may yield FEWER than multi_fpu frames (independent streams): per-slot validity mask
    // utterance CMVN inside the launch (apply tiles of the work list; lean instantiation only)
    int apply_mode;                 // 0 off, 1 mean, 2 mean + variance, 3 SpecAugment mean fills (one completion tile per utterance)
    float* fills;                   // apply_mode 3: [B][n_fmask + n_tmask] fills (output, optional)
    int* utt_done;                  // [B] frame tiles of the utterance whose features and statistics are globally visible; [B] = error flag
    float* utt_mean;                // optional [B][nmel]: the vectors the apply tiles used
    float* utt_istd;
    // mel tables (warp-uniform, read through the constant bank) -- generic (non-static) phase B
    short seg_start[kMaxMel + 3];   // k where segment s begins, s = 0..nmel+1  (segment s feeds bin s (up) and s-1 (down))
    short grp_begin[9];             // mel bins handled by warp w: [grp_begin[w], grp_begin[w+1])
    float2 w_updn[256];             // (up weight -> bin seg(k), down weight -> bin seg(k)-1) * 0.25
};

struct SmemLayout {
    int tile_off[kStages];
    int xbuf_off, pt_off, outs_off, misc_off, bar_off, total;
};

// misc area: mean[128] | istd[128] | colmask[4]
__host__ __device__ inline SmemLayout make_layout(int tile_floats, int nmel)
{
    SmemLayout L;
    int o = 0;
    for (int s = 0; s < kStages; ++s) { L.tile_off[s] = o; o += tile_floats * 4; }
    L.xbuf_off = o;
    const int xbytes = kWarps * kRegionWarp * 4, obytes = (kFT * (nmel + 1) * 4 + 15) & ~15;
    L.pt_off = o;                                        // power-spectrum rows live inside the half-warp regions
    o += xbytes; L.outs_off = o; o += obytes;            // phase C of tile i overlaps phase A of tile i+1
    L.misc_off = o; o += (2 * ((nmel + 3) & ~3) + 8) * 4 + (kTablesL1 ? 0 : (128 + 256) * 8);   // mean | istd | masks | split twiddles (k < 128) | window pairs
    L.bar_off = o; o += 32 + 32 + 48;       // 4 mbarrier slots | 2 tile descriptors (int4) | pending completion signals (count + kSigBatch utterances)
    L.total = o;
    return L;
}

// fp32 evaluation of round_fp32(x / (m + 1e-9)): reciprocal estimate plus one exact-residual
// correction (the 1e-9 is applied to the residual because m + 1e-9 is not an fp32 number).
__device__ __forceinline__ float peak_div(float x, float m, float rcp)
{
    float q = x * rcp;
    float e = fmaf(-q, m, x);
    e = fmaf(-q, 1e-9f, e);
    return fmaf(e, rcp, q);
}

__device__ __forceinline__ int row_class(const int* __restrict__ bounds, int nb, int t)
{
    int c = 0;
    for (int i = 0; i < nb; ++i) c += (bounds[i] <= t);
    return c;
}

#ifdef B200FE_TIMELINE
// Debug build only (tools/timeline.py): per CTA, warp and tile the SM clock at the phase boundaries of the tile loop.
constexpr int kTlTiles = 48, kTlStamps = 8;
static __device__ long long g_timeline[296][8][kTlTiles][kTlStamps];   // one copy per translation unit; tools read group 0's
#define TL_STAMP(k) do { if (lane == 0 && it < kTlTiles && blockIdx.x < 296) g_timeline[blockIdx.x][warp][it][k] = clock64(); } while (0)
#else
#define TL_STAMP(k) do { } while (0)
#endif

struct TileGeom {
    unsigned mask;                  // multi-utterance tiles: bit fl = frame slot fl holds a real frame
    int utt, f0, nvalid, nrows, T;
    bool apply, ready;
};

// Phase B only stores the raw mel energy of (frame, bin) into the staging tile; log / CMVN / masks are
// applied by the compact, warp-uniform phase C so that the straight-line mel code stays small.
__device__ __forceinline__ void emit_bin(float* orow, int j, float e) { orow[j] = e; }


// Word offset (inside the transposition / PT area) of the power-spectrum row of frame slot fl.  Frame slots are dealt to
// (warp, half-warp, pass) as  fl = (slot & 3) + 8 (slot >> 2) + 4 h2,  slot = warp + kWarps * pass  (dual-256 mode: the half-warp
// of frames fl, fl + 1 with  fl = 8 (warp >> 1) + 2 (warp & 1) + 4 h2);  the first frame of a half-warp sits behind its
// transposition area, the second one on top of it.
template <bool kDual>
__device__ __forceinline__ int pt_row(int fl)
{
    const int h2 = (fl >> 2) & 1;
    int warp, second;
    if (kDual) { const int fa = fl & ~1; warp = 2 * (fa >> 3) + ((fa >> 1) & 1); second = fl & 1; }
    else { const int slot = (fl & 3) + 4 * (fl >> 3); warp = slot % kWarps; second = slot / kWarps; }
    return warp * kRegionWarp + h2 * kRegionHW + (second ? 0 : kXWords);
}

// ---------------------------------------------------------------------------------------------
// Phase A building blocks (all inlined; 16 lanes cooperate on one 256-point complex FFT)
// ---------------------------------------------------------------------------------------------
struct FrameCtx {
    float c_pre, inv_win, dc_coef;   // pre-emphasis, 1 / window size, (1 - preemph) or 0
    float pmax, prcp, pscale;        // peak normalisation
    int win;
    // dither (generic kernels only)
    float dither;                    // in input units (already divided by the folded 2^15 scale)
    unsigned long long seed;
    const float* noise_a;            // external noise row of frame a (nullptr -> Philox)
    const float* noise_b;            // ... of frame b (dual mode)
    unsigned utt, ta, tb;            // counters for the generator
};

// Philox4x32-10 counter-based generator -> two standard normals (Box-Muller) for the sample pair
// (2*pair, 2*pair+1) of frame t of utterance utt.
__device__ __forceinline__ float2 philox_normal_pair(unsigned long long seed, unsigned utt, unsigned t, unsigned pair)
{
    unsigned c0 = pair, c1 = t, c2 = utt, c3 = 0x5eed5eedu;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float u1 = ((float)c0 + 0.5f) * 2.3283064365386963e-10f;   // (0, 1)
    const float u2 = ((float)c1 + 0.5f) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// noise for sample j of one frame
__device__ __forceinline__ float dither_at(const FrameCtx& c, const float* __restrict__ ext, unsigned t, int j)
{
    j = min(j, c.win - 1);                       // registers past the window are masked afterwards
    if (ext != nullptr) return __ldg(ext + j);
    const float2 p = philox_normal_pair(c.seed, c.utt, t, (unsigned)(j >> 1));
    return (j & 1) ? p.y : p.x;
}

// 512-point family: the frame's 400 (<= 32*NLOAD) samples are packed as z[n] = y[2n] + j y[2n+1].
// Load, (peak-normalise,) pre-emphasise, remove DC, window:  y[j] = ((x[j]-m) - c (x[j-1]-m)) w[j]
//                                                                 = (x[j] - c x[j-1] - (1-c) m) w[j]   (TA:183-204)
template <int NLOAD, bool kPeak, bool kDither>
__device__ __forceinline__ void load_frame_single(float2 (&v)[16], const float* __restrict__ xf, const float2* __restrict__ wl,
                                                  const FrameCtx& c, int l)
{
    float2 acc2 = make_float2(0.f, 0.f);
#if B200FE_XP_SHFL
    // x[j-1] is the previous lane's second sample (lane 0: lane 15's second sample of the previous register): one shuffle
    // instead of a 4-byte load that collides with the other half-warp's (frames start 0 mod 32 banks apart)
    float2 xall[NLOAD];
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) xall[n2] = *reinterpret_cast<const float2*>(xf + 32 * n2);
#endif
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) {
        const int j = 2 * (l + 16 * n2);
#if B200FE_XP_SHFL
        float2 xr = xall[n2];
        // lane l (> 0) takes lane l-1's y of this register; lane 0 takes lane 15's y of the previous register
        const float snd = (n2 > 0 && l == 15) ? xall[n2 - 1].y : xall[n2].y;
        const float got = __shfl_sync(0xffffffffu, snd, (l + 15) & 15, 16);
        float xp = (n2 == 0 && l == 0) ? xall[0].x : got;
#else
        float2 xr = *reinterpret_cast<const float2*>(xf + 32 * n2);
        float xp = (n2 == 0) ? xf[l == 0 ? 0 : -1] : xf[32 * n2 - 1];   // replicate pad at the frame start (TA:195)
#endif
        if (kPeak) {
            xr.x = peak_div(xr.x, c.pmax, c.prcp) * c.pscale;
            xr.y = peak_div(xr.y, c.pmax, c.prcp) * c.pscale;
            xp = peak_div(xp, c.pmax, c.prcp) * c.pscale;
        }
        if (kDither && c.dither != 0.f) {
            // noise is drawn per (frame, sample) AFTER framing (TA:174-181): x[j-1] carries this frame's n[j-1]
            xr.x = fmaf(c.dither, dither_at(c, c.noise_a, c.ta, j), xr.x);
            xr.y = fmaf(c.dither, dither_at(c, c.noise_a, c.ta, j + 1), xr.y);
            xp = fmaf(c.dither, dither_at(c, c.noise_a, c.ta, j > 0 ? j - 1 : 0), xp);
        }
        if (NLOAD == 16 || n2 == NLOAD - 1) {
            // samples past the window must not enter the mean (their window weight is 0, but whatever
            // sits in shared memory there may be NaN)
            if (j >= c.win) { xr.x = 0.f; xp = 0.f; }
            if (j + 1 >= c.win) xr.y = 0.f;
        }
        acc2 = add2(acc2, xr);
        v[n2].x = fmaf(-c.c_pre, xp, xr.x);
        v[n2].y = fmaf(-c.c_pre, xr.x, xr.y);
    }
    float sum = acc2.x + acc2.y;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float ncdc = -(sum * c.inv_win * c.dc_coef);   // -(1 - preemph) * frame mean
    // (p - c) w = p w - c w: the products p w do not wait for the shuffle reduction above
    float2 wv[NLOAD];
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) { wv[n2] = wl[16 * n2]; v[n2] = mul2(v[n2], wv[n2]); }
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) v[n2] = fma2(bc(ncdc), wv[n2], v[n2]);
#pragma unroll
    for (int n2 = NLOAD; n2 < 16; ++n2) v[n2] = make_float2(0.f, 0.f);
}

// int16 PCM variant of load_frame_single: the tile holds int16 samples; (float)s16 equals the float path's
// x * 2^15 exactly, so the UNSCALED window table is used and the results are bit-identical.
template <int NLOAD, bool kPeak>
__device__ __forceinline__ void load_frame_single_i16(float2 (&v)[16], const short* __restrict__ xf, const float2* __restrict__ wl,
                                                      const FrameCtx& c, int l)
{
    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) {
        const int j = 2 * (l + 16 * n2);
        const int w = *reinterpret_cast<const int*>(xf + 32 * n2);              // samples j (low half) and j + 1 (high half)
        float2 xr = make_float2((float)(short)(w & 0xffff), (float)(w >> 16));
        float xp = (float)((n2 == 0) ? xf[l == 0 ? 0 : -1] : xf[32 * n2 - 1]);   // replicate pad at the frame start (TA:195)
        if (kPeak) {
            // reference: soundfile's s16 / 2^15, then x / (max + 1e-9), rounded to fp32, times 2^15
            xr.x = peak_div(xr.x * 3.0517578125e-05f, c.pmax, c.prcp) * c.pscale;
            xr.y = peak_div(xr.y * 3.0517578125e-05f, c.pmax, c.prcp) * c.pscale;
            xp = peak_div(xp * 3.0517578125e-05f, c.pmax, c.prcp) * c.pscale;
        }
        if (NLOAD == 16 || n2 == NLOAD - 1) {
            if (j >= c.win) { xr.x = 0.f; xp = 0.f; }
            if (j + 1 >= c.win) xr.y = 0.f;
        }
        acc2 = add2(acc2, xr);
        v[n2].x = fmaf(-c.c_pre, xp, xr.x);
        v[n2].y = fmaf(-c.c_pre, xr.x, xr.y);
    }
    float sum = acc2.x + acc2.y;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float ncdc = -(sum * c.inv_win * c.dc_coef);
    float2 wv[NLOAD];
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) { wv[n2] = wl[16 * n2]; v[n2] = mul2(v[n2], wv[n2]); }
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) v[n2] = fma2(bc(ncdc), wv[n2], v[n2]);
#pragma unroll
    for (int n2 = NLOAD; n2 < 16; ++n2) v[n2] = make_float2(0.f, 0.f);
}

// 256-point family: two consecutive real frames a (at xa) and b (at xb) -> z[n] = ya[n] + j yb[n].
// b_ok = false: frame b lies past the utterance's last frame; it is replaced by silence so that whatever the tile buffer holds
// there (stale samples, possibly NaN bit patterns) cannot leak into frame a through the shared FFT's rounding.
template <int NLOAD, bool kPeak, bool kDither>
__device__ __forceinline__ void load_frame_dual(float2 (&v)[16], const float* __restrict__ xa, const float* __restrict__ xb,
                                                const float* __restrict__ wls, const FrameCtx& c, int l, bool b_ok)
{
    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) {
        const int j = l + 16 * n2;
        const int po = (n2 == 0 && l == 0) ? 0 : 16 * n2 - 1;            // replicate pad at the frame start
        float2 xr = make_float2(xa[16 * n2], b_ok ? xb[16 * n2] : 0.f);
        float2 xp = make_float2(xa[po], b_ok ? xb[po] : 0.f);
        if (kPeak) {
            xr.x = peak_div(xr.x, c.pmax, c.prcp) * c.pscale; xr.y = peak_div(xr.y, c.pmax, c.prcp) * c.pscale;
            xp.x = peak_div(xp.x, c.pmax, c.prcp) * c.pscale; xp.y = peak_div(xp.y, c.pmax, c.prcp) * c.pscale;
        }
        if (kDither && c.dither != 0.f) {
            const int jp = j > 0 ? j - 1 : 0;
            xr.x = fmaf(c.dither, dither_at(c, c.noise_a, c.ta, j), xr.x);
            xr.y = fmaf(c.dither, dither_at(c, c.noise_b, c.tb, j), xr.y);
            xp.x = fmaf(c.dither, dither_at(c, c.noise_a, c.ta, jp), xp.x);
            xp.y = fmaf(c.dither, dither_at(c, c.noise_b, c.tb, jp), xp.y);
        }
        if (NLOAD == 16 || n2 == NLOAD - 1) {
            if (j >= c.win) { xr = make_float2(0.f, 0.f); xp = make_float2(0.f, 0.f); }
        }
        acc2 = add2(acc2, xr);
        v[n2] = fma2(bc(-c.c_pre), xp, xr);
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        acc2.x += __shfl_xor_sync(0xffffffffu, acc2.x, o);
        acc2.y += __shfl_xor_sync(0xffffffffu, acc2.y, o);
    }
    const float2 ncdc2 = make_float2(-(acc2.x * c.inv_win * c.dc_coef), -(acc2.y * c.inv_win * c.dc_coef));
#pragma unroll
    for (int n2 = 0; n2 < NLOAD; ++n2) v[n2] = mul2(add2(v[n2], ncdc2), bc(wls[16 * n2]));
#pragma unroll
    for (int n2 = NLOAD; n2 < 16; ++n2) v[n2] = make_float2(0.f, 0.f);
}

// 256-point complex FFT across the half-warp: DFT-16 over n2, twiddle W_256^(n1*klo), transposition
// through shared memory (row n1 = l, column klo), DFT-16 over n1 -> Z[l + 16 r] in v[r].
__device__ __forceinline__ void fft256_halfwarp(float2 (&v)[16], const float2 (&tw)[16], float2* __restrict__ xbuf, int l)
{
    dft16(v);
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = c_mul(v[k], tw[k].x, tw[k].y);
#pragma unroll
    for (int k = 0; k < 16; ++k) xbuf[l * kXRow + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = xbuf[n1 * kXRow + l];
    __syncwarp();
    dft16(v);
}

// Conjugate-pair exchange: the lane's own r = 0..7 (bins l + 16 r) pair with lane (16-l)&15, register
// 15-r (bins 256 - l - 16 r).  Lane 0 pairs 16 r with 16 (16 - r): shift by one register, bin 0 with itself.
__device__ __forceinline__ void pair_exchange(const float2 (&v)[16], float2 (&rc)[8], int l, int h2)
{
    const int partner = (16 * h2) + ((16 - l) & 15);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        rc[r].x = __shfl_sync(0xffffffffu, v[15 - r].x, partner);
        rc[r].y = __shfl_sync(0xffffffffu, v[15 - r].y, partner);
    }
    if (l == 0) {
#pragma unroll
        for (int r = 7; r >= 1; --r) rc[r] = rc[r - 1];
        rc[0] = v[0];
    }
}

#define B200FE_MEL_DEVICE_CODE
#define MGROUP_BEGIN(w) __device__ __forceinline__ void mel_static_group##w(const float4* __restrict__ pcol, float* __restrict__ orow) { \
        float au = 0.f, ad = 0.f, au1 = 0.f, ad1 = 0.f, up_prev = 0.f; float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f); int cur = -1; (void)p4; (void)cur;
#define MK(k, wu, wd) { if (((k) >> 2) != cur) { cur = (k) >> 2; p4 = pcol[cur]; } \
        const float p = ((k) & 3) == 0 ? p4.x : ((k) & 3) == 1 ? p4.y : ((k) & 3) == 2 ? p4.z : p4.w; \
        if ((k) & 1) { if ((wu) != 0.f) au1 = fmaf((wu), p, au1); if ((wd) != 0.f) ad1 = fmaf((wd), p, ad1); } \
        else         { if ((wu) != 0.f) au = fmaf((wu), p, au);   if ((wd) != 0.f) ad = fmaf((wd), p, ad); } }
#define MEND0() { up_prev = au + au1; au = 0.f; ad = 0.f; au1 = 0.f; ad1 = 0.f; }
#define MEND(j) { emit_bin(orow, (j), up_prev + (ad + ad1)); up_prev = au + au1; au = 0.f; ad = 0.f; au1 = 0.f; ad1 = 0.f; }
#define MGROUP_END(w) }
#include "mel_static_default.inc"
#undef MGROUP_BEGIN
#undef MK
#undef MEND0
#undef MEND
#undef MGROUP_END
#undef B200FE_MEL_DEVICE_CODE

// CMVN-apply tile (utterance CMVN inside the fused launch): rows [row0, row0 + kApplyRows) of utterance utt are normalised in
// place, (x - mean) * istd with the same fp64 -> fp32 vectors finalize_kernel derives from the column sums.  The tile may start
// once every frame tile of the utterance has been signalled (their stores and atomics are then visible at L2): normally thread 0
// saw that when it claimed the tile (ready), otherwise it waits here.  The wait is bounded: the tiles it depends on precede this
// one in the work list, so they are owned by resident CTAs that never wait on anything later in the list.  The rows make one round
// trip: ONE bulk copy (TMA) brings them -- still L2-resident -- into the transposition / staging / PT buffers, which are idle
// between two frame tiles, the column vectors are derived from the fp64 sums (ld.global.cg) while they fly, the rows are
// normalised in shared memory and ONE bulk copy writes them back.  No registers are held across the latency.
// Completion counters of the apply tiles.  Producer: ONE fence.acq_rel.gpu (orders the CTA's earlier feature stores and
// statistics atomics -- cumulative over the preceding CTA barrier -- before the increments), then relaxed reductions; a fence
// costs 0.5 - 1 us on the critical path of warp 0, hence the batching.  Consumer: relaxed polling, then ONE fence.acq_rel
// before the dependent reads.
__device__ __forceinline__ int ld_relaxed_gpu(const int* p)
{
    int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add(int* p, int v)
{
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ bool apply_cmvn_tile(const FbankArgs& a, int utt, int row0, int Tu, bool ready, float* s_mean, float* s_istd,
                                                unsigned char* s_rows, uint64_t* bar, uint32_t parity, int tid, int nmel)
{
    const int T = min(Tu, a.Tmax);
    const int nr = max(min(T - row0, kApplyRows), 0);
    const uint32_t bytes = (uint32_t)nr * (uint32_t)nmel * 4u;   // nmel % 4 == 0: a multiple of 16
    float* grows = a.out + ((long long)utt * a.Tmax + row0) * nmel;
    if (tid == 0) {
        if (!ready) {
            const int need = (T + kFT - 1) / kFT;
            int spins = 0;
            while (ld_relaxed_gpu(a.utt_done + utt) < need) {
                __nanosleep(128);
                if (++spins > (1 << 22)) { atomicExch(a.utt_done + a.B, 1 + utt); break; }
            }
        }
        // acquire side of the completion counter (the count may have been observed two tiles ago, when the tile was claimed);
        // the rows were written through the generic proxy by other SMs, the bulk copy reads them through the async proxy
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (bytes > 0) { mbar_expect_tx(bar, bytes); tma_load_1d(s_rows, grows, bytes, bar); }
    }
    __syncthreads();          // orders thread 0's acquire before everybody's reads of the column sums
    // the column vectors are derived while the rows fly
    if (tid < nmel) {
        const double* sb = a.stats + (long long)utt * a.stats_stride;
        double mean = 0.0, istd = 1.0;
        if (Tu > 0) {
            mean = __ldcg(sb + tid) / Tu;
            if (a.apply_mode == 2) {
                const double var = __ldcg(sb + (long long)a.n_cls * nmel + tid) / Tu - mean * mean;
                istd = 1.0 / sqrt(var > 1e-20 ? var : 1e-20);
            }
        }
        s_mean[tid] = (float)mean; s_istd[tid] = (float)istd;
        if (row0 == 0 && a.utt_mean != nullptr) {
            a.utt_mean[(long long)utt * nmel + tid] = (float)mean;
            a.utt_istd[(long long)utt * nmel + tid] = (float)istd;
        }
    }
    __syncthreads();
    if (bytes == 0) return false;
    mbar_wait(bar, parity);
    const int nq = nmel >> 2, n4 = nr * nq;
    float4* rows4 = reinterpret_cast<float4*>(s_rows);
    const float4* m4 = reinterpret_cast<const float4*>(s_mean);
    const float4* i4 = reinterpret_cast<const float4*>(s_istd);
#pragma unroll 4
    for (int e = tid; e < n4; e += kThreads) {
        const int q = e % nq;
        const float4 m = m4[q], sc = i4[q], v = rows4[e];
        rows4[e] = make_float4((v.x - m.x) * sc.x, (v.y - m.y) * sc.y, (v.z - m.z) * sc.z, (v.w - m.w) * sc.w);
    }
    fence_proxy_async();      // the normalised rows (generic-proxy writes) become visible to the bulk store
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grows), "r"(smem_u32(s_rows)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the buffers belong to the next tile's phase A afterwards
    }
    return true;
}

// SpecAugment mean fills inside the fused launch (apply_mode 3; the reference's default masks, specaugment.py:47-106, on top of
// global or no CMVN): the work list carries ONE completion tile per utterance; when every frame tile of the utterance has been
// signalled, the CTA that took the tile derives the fills from the row-class column sums exactly as finalize_kernel does
// (x.mean() of the current array before every mask, later masks see the earlier fills) and overwrites the masked cells -- whole
// float4 stores where four columns are masked, scalar stores otherwise, nothing is read.  Replaces the finalize launch and the
// mask pass over the batch (8 + 70 us on BASELINE config 3).  `scratch`: >= 256 bytes of idle shared memory.
__device__ __forceinline__ void apply_mask_tile(const FbankArgs& a, int utt, int Tu, bool ready, unsigned char* scratch, int tid, int nmel)
{
    const int T = min(Tu, a.Tmax);
    if (tid == 0) {
        if (!ready) {
            const int need = (T + kFT - 1) / kFT;
            int spins = 0;
            while (ld_relaxed_gpu(a.utt_done + utt) < need) {
                __nanosleep(128);
                if (++spins > (1 << 22)) { atomicExch(a.utt_done + a.B, 1 + utt); break; }
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    double* red = reinterpret_cast<double*>(scratch);                        // [kWarps]
    int* cls_lo = reinterpret_cast<int*>(red + kWarps);                      // [kMaxRowClasses + 1]
    float* s_fill = reinterpret_cast<float*>(cls_lo + kMaxRowClasses + 3);   // [kMaxFreqMasks + kMaxTimeMasks]
    const int nb = a.n_cls - 1;
    const int* bounds = (a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
    if (tid == 0) {
        cls_lo[0] = 0;
        for (int c = 0; c < nb; ++c) cls_lo[c + 1] = min(max(bounds ? __ldg(bounds + c) : T, 0), T);
        cls_lo[a.n_cls] = T;
    }
    __syncthreads();          // orders thread 0's acquire before everybody's reads of the column sums
    const double* sb = a.stats + (long long)utt * a.stats_stride;
    const bool col = tid < nmel;
    double S[kMaxRowClasses];
    double part = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxRowClasses; ++c) {
        S[c] = (col && c < a.n_cls) ? __ldcg(sb + (long long)c * nmel + tid) : 0.0;
        part += S[c];
    }
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w];
        return t;
    };
    double total = block_sum(part);
    const int nm = a.n_fmask + a.n_tmask;
    const int* mk = a.masks + (long long)utt * nm * 2;
    const double cells = (double)T * (double)nmel;
    for (int i = 0; i < nm; ++i) {
        const double fill = cells > 0 ? (double)(float)(total / cells) : 0.0;     // numpy: x[...] = x.mean() of the current array
        int lo = __ldg(mk + 2 * i), hi = __ldg(mk + 2 * i + 1);
        double delta = 0.0;
        if (i < a.n_fmask) {
            lo = max(lo, 0); hi = min(hi, nmel);
            if (col && tid >= lo && tid < hi) {
#pragma unroll
                for (int c = 0; c < kMaxRowClasses; ++c)
                    if (c < a.n_cls) { const double nv = fill * (double)(cls_lo[c + 1] - cls_lo[c]); delta += nv - S[c]; S[c] = nv; }
            }
        } else {
            lo = max(lo, 0); hi = min(hi, T);
            if (col) {
#pragma unroll
                for (int c = 0; c < kMaxRowClasses; ++c)
                    if (c < a.n_cls && cls_lo[c] >= lo && cls_lo[c + 1] <= hi && cls_lo[c + 1] > cls_lo[c]) {
                        const double nv = fill * (double)(cls_lo[c + 1] - cls_lo[c]); delta += nv - S[c]; S[c] = nv;
                    }
            }
        }
        total += block_sum(delta);
        if (tid == 0) { s_fill[i] = (float)fill; if (a.fills != nullptr) a.fills[(long long)utt * nm + i] = (float)fill; }
    }
    __syncthreads();
    int tlo[kMaxTimeMasks], thi[kMaxTimeMasks];
    float tfill[kMaxTimeMasks];
#pragma unroll
    for (int i = 0; i < kMaxTimeMasks; ++i) {
        tlo[i] = 0; thi[i] = 0; tfill[i] = 0.f;
        if (i < a.n_tmask) { tlo[i] = __ldg(mk + 2 * (a.n_fmask + i)); thi[i] = __ldg(mk + 2 * (a.n_fmask + i) + 1); tfill[i] = s_fill[a.n_fmask + i]; }
    }
    const int nq = nmel >> 2;                     // nmel % 4 == 0 (the static option set)
    const int slots = kThreads / nq, slot = tid / nq, qd = tid - slot * nq;
    if (slot < slots) {
        int fhit[4] = {-1, -1, -1, -1};
        float ffill[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < a.n_fmask; ++i) {
            const int lo = __ldg(mk + 2 * i), hi = __ldg(mk + 2 * i + 1);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * qd + c >= lo && 4 * qd + c < hi) { fhit[c] = i; ffill[c] = s_fill[i]; }
        }
        const bool anyf = fhit[0] >= 0 || fhit[1] >= 0 || fhit[2] >= 0 || fhit[3] >= 0;
        const bool allf = fhit[0] >= 0 && fhit[1] >= 0 && fhit[2] >= 0 && fhit[3] >= 0;
        float* base = a.out + (long long)utt * a.Tmax * nmel + 4 * qd;
        for (int r = slot; r < T; r += slots) {
            int thit = -1; float tf = 0.f;
#pragma unroll
            for (int i = 0; i < kMaxTimeMasks; ++i) if (r >= tlo[i] && r < thi[i]) { thit = i; tf = tfill[i]; }    // later masks win
            float* ptr = base + (long long)r * nmel;
            if (thit >= 0) *reinterpret_cast<float4*>(ptr) = make_float4(tf, tf, tf, tf);             // time masks come after frequency masks
            else if (allf) *reinterpret_cast<float4*>(ptr) = make_float4(ffill[0], ffill[1], ffill[2], ffill[3]);
            else if (anyf) {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (fhit[c] >= 0) ptr[c] = ffill[c];
            }
        }
    }
}

// kDual: padded window of 256 samples (8 kHz family).  Two consecutive real frames a, b are packed as
// z[n] = ya[n] + j yb[n]; after the same 256-point complex FFT, 2 Xa[k] = Z[k] + conj Z[256-k] and
// 2j Xb[k] = Z[k] - conj Z[256-k], so the conjugate-pair exchange yields both power spectra without any
// split twiddle (k = 0..127; the Nyquist bin has zero mel weight, TA:627).
// kI16: the waveform is int16 PCM (2 bytes per sample over PCIe / HBM); 512-point family only.
// kMulti: multi-utterance tiles for lock-step streaming (instantiated for the two default option sets only).
// kApply: the lean kernel plus the CMVN-apply tiles of b200fe_build_work_list_device (utterance CMVN inside the launch).
// kLean: no CMVN / zero masks / packed output inside the launch -- the plain and the statistics (utterance CMVN by post pass)
// modes of the default option set; the same arithmetic with those runtime switches compiled out (-2 % time, A/B measured).
template <int NLOAD, bool kStaticMel, bool kPeak, bool kDual, bool kI16 = false, bool kMulti = false, bool kLean = false, bool kApply = false>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) fbank_fused_kernel(const __grid_constant__ FbankArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L = make_layout(a.tile_floats, a.nmel);
    float2* xbuf_all = reinterpret_cast<float2*>(smem + L.xbuf_off);
    float* pt = reinterpret_cast<float*>(smem + L.pt_off);
    float* outs = reinterpret_cast<float*>(smem + L.outs_off);
    float* s_mean = reinterpret_cast<float*>(smem + L.misc_off);
    float* s_istd = s_mean + ((a.nmel + 3) & ~3);
    unsigned* s_cmask = reinterpret_cast<unsigned*>(s_istd + ((a.nmel + 3) & ~3));   // [0..3] column bits, [4] row bits
    float2* s_stw = reinterpret_cast<float2*>(s_cmask + 8);
    float2* s_win = s_stw + 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    int4* s_desc = reinterpret_cast<int4*>(bars + 4);          // 16-byte aligned: four 8-byte slots are reserved for mbarriers
    int* s_sig = reinterpret_cast<int*>(bars + 8);             // [0] count, [1..kSigBatch] utterances whose tile completed (thread 0 only)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int h2 = lane >> 4, l = lane & 15;
    const int hw = warp * 2 + h2;
    float* region = reinterpret_cast<float*>(xbuf_all) + warp * kRegionWarp + h2 * kRegionHW;   // this half-warp's transposition area + first PT row
    float2* xbuf = reinterpret_cast<float2*>(region);
    (void)hw;
    // the straight-line mel path fixes num_mel_bins and the power spectrum at compile time
    const int nmel = kStaticMel ? B200FE_STATIC_NMEL : a.nmel;
    const bool use_power = kStaticMel ? true : (a.use_power != 0);
    const int ostride = nmel + 1;

    // ---- per-lane constants (live in registers across all tiles) ----
    // window (x 2^15) as float2 pairs in shared memory: lane l reads pair l + 16 n2 (conflict free)
    if (!kTablesL1) for (int k = tid; k < 256; k += kThreads) s_win[k] = make_float2(__ldg(a.window + 2 * k), __ldg(a.window + 2 * k + 1));
    const float2* wl = (kTablesL1 ? reinterpret_cast<const float2*>(a.window) : s_win) + l;
    const float* wls = (kTablesL1 ? a.window : reinterpret_cast<const float*>(s_win)) + l;     // dual-256 mode: scalars w[l + 16 n2]
    float2 tw[16];
#pragma unroll
    for (int k = 1; k < 16; ++k) tw[k] = __ldg(a.twiddle + l * 16 + k);
    // split twiddles -j W_512^k live in shared memory (lane l reads k = l + 16 r: conflict free)
    if (!kTablesL1) for (int k = tid; k < 128; k += kThreads) s_stw[k] = __ldg(a.split_tw + k);   // k = l + 16 r, r < 8
    const float2* stw = (kTablesL1 ? a.split_tw : s_stw) + l;
    const float2 sw0 = kStwRot ? __ldg(a.split_tw + l) : make_float2(0.f, 0.f);      // -j W_512^l

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        mbar_init(&bars[kStages], kWarps);        // "tile consumed": every warp has its last frames in registers
        mbar_init(&bars[2], 1);                   // apply tiles: bulk load of the rows
        s_sig[0] = 0;
        fence_mbar_init();
    }
    // everything above reads plan constants only; the work list, the counters and all caller data come after the preceding kernel
    pdl_launch_dependents();
    pdl_wait();
    // global CMVN vectors (or the identity) are staged once; per-utterance vectors per tile
    const bool cm_per_utt = kLean ? false : (a.cm_mean != nullptr && a.cm_stride != 0);
    if (tid < nmel) {
        const bool on = a.cm_mean != nullptr && !cm_per_utt;
        s_mean[tid] = on ? __ldg(a.cm_mean + tid) : 0.f;
        s_istd[tid] = on ? __ldg(a.cm_istd + tid) : 1.f;
    }
    if (tid < 8) s_cmask[tid] = 0u;
    __syncthreads();

    const float c_pre = a.preemph;
    const float inv_win = 1.0f / (float)a.win;
    const float dc_coef = a.remove_dc ? (float)(1.0 - (double)a.preemph) : 0.0f;
    const bool zmask = kLean ? false : (a.mask_zero && a.masks != nullptr);
    const int nmask = a.n_fmask + a.n_tmask;
    // statistics without SpecAugment row classes are reduced inside phase C (no staging write-back, no extra barrier)
    // (with SpecAugment row classes: for every tile that lies inside ONE class, which is all but a handful per utterance)
    const bool stats_cap = a.stats != nullptr && !zmask && nmel <= kThreads;
    const bool has_cls = !kLean && a.row_bounds != nullptr && a.n_cls > 1;
    // the same thread <-> (row part, column) mapping serves the plain copy-out: constant strides, no index arithmetic
    // ... and the CMVN / zero-mask epilogue of the training front end (BASELINE config 3): the thread's column fixes its CMVN pair and
    // its frequency-mask bit, its rows are rg + k * parts, so the per-element row / column arithmetic of the generic loop disappears
    const bool epi_rowpart = !kLean && a.stats == nullptr && a.out != nullptr && (a.cm_mean != nullptr || zmask) && nmel <= kThreads;
    const bool rowpart_nostats = (!zmask && nmel <= kThreads && B200FE_ROWPART_PLAIN != 0 && a.stats == nullptr) || epi_rowpart;

    // ---- tile scheduler ------------------------------------------------------------------------
    // Thread 0 resolves tile descriptors (id, utterance, first frame, frame count of the utterance) TWO tiles
    // ahead and publishes them through shared memory, so the dependent global loads (atomic counter ->
    // tile table -> sample count) never sit on any warp's critical path.
    struct Desc { int id, utt, f0, T; };
    const bool dyn = a.tile_table != nullptr;
    const int ntiles = a.ntiles_ptr != nullptr ? __ldg(a.ntiles_ptr) : a.ntiles;
    constexpr bool multi = kMulti;
    auto resolve = [&](int id) -> Desc {               // thread 0 only
        Desc d; d.id = id; d.utt = 0; d.f0 = 0; d.T = 0;
        if (id < ntiles) {
            if (multi) {                                    // the tile's utterances form one run of consecutive output rows
                d.utt = id * a.multi_upt;
                const int nu = min(a.multi_upt, a.B - d.utt);
                d.T = a.multi_fpu * nu;
                unsigned m = 0xffffffffu;               // f0 (always 0 for these tiles) carries the mask of valid frame slots
                if (a.multi_ragged) {
                    m = 0u;
                    for (int j = 0; j < nu; ++j) {
                        const unsigned n = (unsigned)__ldg(a.nsamp + d.utt + j);
                        const int Tj = n >= (unsigned)a.win ? min((int)(1u + (n - (unsigned)a.win) / (unsigned)a.shift), a.multi_fpu) : 0;
                        m |= ((1u << Tj) - 1u) << (j * a.multi_fpu);
                    }
                }
                d.f0 = (int)m;
                return d;
            }
            if (dyn) { const int2 e = __ldg(a.tile_table + id); d.utt = e.x; d.f0 = e.y; }
            else { d.utt = (int)((unsigned)id / (unsigned)a.tiles_per_utt); d.f0 = (id - d.utt * a.tiles_per_utt) * kFT; }
            if (d.f0 < 0) return d;                                     // padding tile: rows [-f0 - 1, +kPadTileRows) are zeroed
            const unsigned n = (unsigned)__ldg(a.nsamp + (kApply ? (d.utt & ~kApplyBit) : d.utt));       // < 2^31 samples per utterance
            d.T = n >= (unsigned)a.win ? (int)(1u + (n - (unsigned)a.win) / (unsigned)a.shift) : 0;
            if (kApply && (d.utt & kApplyBit)) {
                // claimed two tiles ahead: normally every frame tile of the utterance has been signalled by now, and the
                // apply tile will start without a round trip to the counter
                const int need = (min(d.T, a.Tmax) + kFT - 1) / kFT;
                if (ld_relaxed_gpu(a.utt_done + (d.utt & ~kApplyBit)) >= need) d.utt |= kReadyBit;
            }
        }
        return d;
    };
    auto geom = [&](const Desc& d) -> TileGeom {
        TileGeom g;
        g.utt = d.utt; g.f0 = d.f0; g.T = d.T; g.apply = false; g.ready = false; g.mask = 0xffffffffu;
        if (multi) {                 // slots of consecutive utterances; which of them hold frames says the mask
            g.f0 = 0; g.mask = (unsigned)d.f0;
            g.nvalid = min(max(d.T, 0), kFT);
            g.nrows = g.nvalid;
            return g;
        }
        if (kApply && (d.utt & kApplyBit)) {   // CMVN-apply tile: no frames, no padding rows; handled after the phase-C block
            g.utt = d.utt & ~(kApplyBit | kReadyBit); g.nvalid = 0; g.nrows = 0; g.apply = true; g.ready = (d.utt & kReadyBit) != 0;
            return g;
        }
        if (d.f0 < 0) {          // padding tile of the compact list: no frames, rows [row0, row0 + kPadTileRows) clipped at Tmax
            g.f0 = -d.f0 - 1;
            g.nvalid = 0;
            g.nrows = max(min(a.Tmax - g.f0, kPadTileRows), 0);
            return g;
        }
        g.nvalid = min(max(d.T - d.f0, 0), kFT);
        g.nrows = (dyn || a.out_offsets != nullptr) ? g.nvalid : min(a.Tmax - d.f0, kFT);
        return g;
    };
    auto issue_load = [&](const TileGeom& g, int stage) {      // called by thread 0 only
        if (g.nvalid <= 0) return;
        const int esz = kI16 ? 2 : 4;
        if (multi) {
            const int nu = g.nvalid / a.multi_fpu;
            uint32_t total = 0;
            for (int j = 0; j < nu; ++j) {
                const int Tj = __popc((g.mask >> (j * a.multi_fpu)) & ((1u << a.multi_fpu) - 1u));
                if (Tj > 0) total += (uint32_t)((((Tj - 1) * a.shift + a.win) * esz + 15) & ~15);
            }
            mbar_expect_tx(&bars[stage], total);          // (a tile whose streams all delivered nothing completes with 0 bytes)
            for (int j = 0; j < nu; ++j) {
                const int Tj = __popc((g.mask >> (j * a.multi_fpu)) & ((1u << a.multi_fpu) - 1u));
                if (Tj <= 0) continue;
                const uint32_t bytes = (uint32_t)((((Tj - 1) * a.shift + a.win) * esz + 15) & ~15);
                const long long eoff = a.wav_offsets ? __ldg(a.wav_offsets + g.utt + j) : (long long)(g.utt + j) * a.wav_stride;
                tma_load_1d(smem + L.tile_off[stage] + (size_t)j * a.multi_span * esz, reinterpret_cast<const char*>(a.wav) + eoff * esz, bytes, &bars[stage]);
            }
            return;
        }
        const int nsmp = (g.nvalid - 1) * a.shift + a.win;
        const uint32_t bytes = (uint32_t)((nsmp * esz + 15) & ~15);
        const long long eoff = (a.wav_offsets ? __ldg(a.wav_offsets + g.utt) : (long long)g.utt * a.wav_stride) + (long long)g.f0 * a.shift;
        const char* src = reinterpret_cast<const char*>(a.wav) + eoff * esz;
        mbar_expect_tx(&bars[stage], bytes);
        tma_load_1d(smem + L.tile_off[stage], src, bytes, &bars[stage]);
    };

    // ---- descriptor pipeline (frame / padding tiles of the batch kernels) ------------------------------------------------------
    // resolve() is a chain of three dependent global accesses (work counter -> tile table -> sample count, ~2 us): run in one go at
    // the top of a tile it stalls warp 0 for a sixth of the tile period, and every CTA barrier then waits for warp 0 (ncu source
    // page, second session of round 2: 11 % of warp 0's samples on that chain, 7.5 % of ALL warp time at the phase-A barrier).
    // The chain is therefore spread over three tiles: each tile issues one access per stage and consumes what the previous tile
    // requested, so thread 0 never waits for global memory.  In flight (registers of thread 0): the claimed id, the table entry
    // of the id claimed one tile earlier, the sample count of the entry loaded one tile earlier; the rest of the state and the
    // finished descriptor rest in shared memory (s_pipe, private to thread 0) so that nothing else stays live across phase A.
    constexpr bool kPipe = !kMulti && !kApply && B200FE_DESC_PIPE != 0;
    int* s_pipe = s_sig;        // [0] id, [1] utterance, [2] first frame of the entry whose sample count is in flight; [3] id whose entry is in flight;
                                // [4] next id of the static schedule; [5..8] descriptor finished at the top of this tile (published after the phase-A barrier)
    int p_id = 0;
    int2 p_ent = make_int2(0, 0);
    unsigned p_n = 0u;
    auto claim = [&]() -> int {                            // thread 0
        if (dyn) return atomicAdd(a.work_counter, 1);
        const int id = s_pipe[4];
        s_pipe[4] = id + (int)gridDim.x;
        return id;
    };
    auto entry_of = [&](int id) -> int2 {                  // (utterance, first frame); issues the load, does not wait for it
        if (id >= ntiles) return make_int2(0, 0);
        if (dyn) return __ldg(a.tile_table + id);
        const int u = (int)((unsigned)id / (unsigned)a.tiles_per_utt);
        return make_int2(u, (id - u * a.tiles_per_utt) * kFT);
    };
    auto nsamp_of = [&](int id, int2 e) -> unsigned {      // padding tiles (first frame < 0) need no sample count
        if (id >= ntiles || e.y < 0) return 0u;
        return (unsigned)__ldg(a.nsamp + e.x);
    };
    auto finish = [&](int id, int2 e, unsigned n) -> Desc {
        Desc d; d.id = id; d.utt = 0; d.f0 = 0; d.T = 0;
        if (id < ntiles) {
            d.utt = e.x; d.f0 = e.y;
            if (e.y >= 0) d.T = n >= (unsigned)a.win ? (int)(1u + (n - (unsigned)a.win) / (unsigned)a.shift) : 0;
        }
        return d;
    };

    int it = 0;
    uint32_t phase_bits = 0, consumed_phase = 0;
    // in-launch utterance CMVN: thread 0 collects the utterances of completed frame tiles (a tile is complete at the next CTA-wide
    // barrier, when every warp's feature stores and statistics atomics have been issued) and publishes kSigBatch of them behind
    // ONE fence -- a fence per tile sits on warp 0's critical path and was measured at 5.6 % of the step.
    int pending = -1;
    auto flush_signals = [&]() {          // thread 0
        const int n = s_sig[0];
        if (n > 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = 1; i <= n; ++i) red_relaxed_gpu_add(a.utt_done + s_sig[i], 1);
            s_sig[0] = 0;
        }
    };
    auto signal_pending = [&](bool force) {          // thread 0, after a CTA barrier that follows the tile's phase C
        if (pending >= 0) { const int n = s_sig[0] + 1; s_sig[n] = pending; s_sig[0] = n; }
        if (force || s_sig[0] >= kSigBatch) flush_signals();
    };

    if (tid == 0 && kPipe) {
        // five tiles claimed up front: two finished descriptors, three stages of the pipeline primed (independent loads: they overlap)
        if (!dyn) s_pipe[4] = (int)blockIdx.x;
        const int i0 = claim(), i1 = claim(), i2 = claim(), i3 = claim();
        p_id = claim();
        const int2 e0 = entry_of(i0), e1 = entry_of(i1), e2 = entry_of(i2);
        p_ent = entry_of(i3);
        const unsigned n0 = nsamp_of(i0, e0), n1 = nsamp_of(i1, e1);
        p_n = nsamp_of(i2, e2);
        const Desc d0 = finish(i0, e0, n0), d1 = finish(i1, e1, n1);
        s_desc[0] = make_int4(d0.id, d0.utt, d0.f0, d0.T);
        s_desc[1] = make_int4(d1.id, d1.utt, d1.f0, d1.T);
        s_pipe[0] = i2; s_pipe[1] = e2.x; s_pipe[2] = e2.y; s_pipe[3] = i3;
    } else if (tid == 0) {
        const int id0 = dyn ? atomicAdd(a.work_counter, 1) : (int)blockIdx.x;
        const int id1 = dyn ? atomicAdd(a.work_counter, 1) : (int)(blockIdx.x + gridDim.x);
        const Desc d0 = resolve(id0), d1 = resolve(id1);
        s_desc[0] = make_int4(d0.id, d0.utt, d0.f0, d0.T);
        s_desc[1] = make_int4(d1.id, d1.utt, d1.f0, d1.T);
    }
    __syncthreads();
    Desc cur, nxt;
    { const int4 v0 = s_desc[0], v1 = s_desc[1]; cur = Desc{v0.x, v0.y, v0.z, v0.w}; nxt = Desc{v1.x, v1.y, v1.z, v1.w}; }
    TileGeom g = geom(cur);
    if (a.use_tma && tid == 0 && cur.id < ntiles) issue_load(g, 0);

    for (; cur.id < ntiles; ++it) {
        const int utt = g.utt, f0 = g.f0, nvalid = g.nvalid, nrows = g.nrows;
        float* xs = reinterpret_cast<float*>(smem + L.tile_off[0]);
        const TileGeom gn = geom(nxt);
        Desc fut; fut.id = ntiles; fut.utt = 0; fut.f0 = 0; fut.T = 0;
        int4 nxt_desc = make_int4(ntiles, 0, 0, 0);
        if (tid == 0 && kPipe) {
            // descriptor of the tile after next from the sample count requested one tile ago; then every stage moves on by one
            const int pe_id = s_pipe[3];
            const Desc d = finish(s_pipe[0], make_int2(s_pipe[1], s_pipe[2]), p_n);
            s_pipe[5] = d.id; s_pipe[6] = d.utt; s_pipe[7] = d.f0; s_pipe[8] = d.T;
            s_pipe[0] = pe_id; s_pipe[1] = p_ent.x; s_pipe[2] = p_ent.y;
            p_n = nsamp_of(pe_id, p_ent);
            s_pipe[3] = p_id;
            p_ent = entry_of(p_id);
            p_id = claim();
        } else if (tid == 0) {
            // descriptor of the tile after next: its loads complete while this tile is being computed
            const int id2 = dyn ? atomicAdd(a.work_counter, 1) : nxt.id + (int)gridDim.x;
            fut = resolve(nxt.id < ntiles ? id2 : ntiles);
        }

        TL_STAMP(0);
        if (a.use_tma) {
            // the tile buffer has been free since the last phase-A barrier: when the current tile carries no
            // frames (padded grid) the next tile's load can go out right away, otherwise after this tile's phase A
            if (nvalid <= 0 && tid == 0 && nxt.id < ntiles) issue_load(gn, 0);
            // the mbarrier phase advances only for tiles that actually carried a load
            if (nvalid > 0) { mbar_wait(&bars[0], phase_bits & 1u); phase_bits ^= 1u; }
        } else if (nvalid > 0) {
            // generic path (unaligned base / stride): cooperative coalesced loads
            const int nu = multi ? nvalid / a.multi_fpu : 1;
            for (int j = 0; j < nu; ++j) {
                const int Tj = multi ? __popc((g.mask >> (j * a.multi_fpu)) & ((1u << a.multi_fpu) - 1u)) : nvalid;
                if (Tj <= 0) continue;
                const int nsmp = (Tj - 1) * a.shift + a.win;
                const long long eoff = (a.wav_offsets ? __ldg(a.wav_offsets + utt + j) : (long long)(utt + j) * a.wav_stride) + (long long)f0 * a.shift;
                if (kI16) {
                    const short* src = reinterpret_cast<const short*>(a.wav) + eoff;
                    short* xd = reinterpret_cast<short*>(xs) + j * a.multi_span;
                    for (int i = tid; i < nsmp; i += kThreads) xd[i] = __ldg(src + i);
                } else {
                    const float* src = a.wav + eoff;
                    float* xd = xs + j * a.multi_span;
                    for (int i = tid; i < nsmp; i += kThreads) xd[i] = __ldg(src + i);
                }
            }
            __syncthreads();
        }

        TL_STAMP(1);
        if (nvalid > 0) {
            // ================= phase A: half-warp per frame (pair of frames in dual-256 mode) =================
            FrameCtx fc;
            fc.c_pre = c_pre; fc.inv_win = inv_win; fc.dc_coef = dc_coef; fc.win = a.win;
            fc.pmax = 1.0f; fc.prcp = 0.0f; fc.pscale = 1.0f;
            // the window table carries the 2^15 input scale unless peak normalisation applies it on load
            fc.dither = kPeak ? a.dither : a.dither / a.in_scale;
            fc.seed = a.dither_seed; fc.utt = (unsigned)utt; fc.ta = fc.tb = 0; fc.noise_a = fc.noise_b = nullptr;
            if (kPeak) {
                // reference: x / (max + 1e-9) in fp64, rounded to fp32, times 2^(bits-1) (datatrans.py:24-25,73-74)
                fc.pmax = __ldg(a.peak + utt);
                fc.prcp = (float)(1.0 / ((double)fc.pmax + 1e-9));
                fc.pscale = a.in_scale;
            }
#pragma unroll 1
            for (int sub = 0; sub < (kDual ? 1 : 2); ++sub) {
                // Frame slot inside the tile: the two half-warps of a warp store PT columns 4 frames apart
                // (disjoint banks).  Both half-warps run in lock step (full-mask shuffles); a half-warp whose
                // frame is past the utterance end computes on stale shared memory and skips its stores.
                const int slot = warp + kWarps * sub;
                const int fl = kDual ? (8 * (warp >> 1) + 2 * (warp & 1) + 4 * h2)            // frames fl (a), fl + 1 (b)
                                     : ((slot & 3) + 8 * (slot >> 2) + 4 * h2);
                const bool fvalid = fl < nvalid && (!multi || ((g.mask >> fl) & 1u));
                const bool last_pass = kDual || sub == 1;
                // (ragged multi-utterance tiles: a warp whose two half-warps hold no real frame skips the pass as well)
                const unsigned wm = multi ? (g.mask >> (fl - 4 * h2)) : 0xffffffffu;
                if (fl - 4 * h2 >= nvalid || (multi && ((wm | (wm >> 4) | (kDual ? ((wm >> 1) | (wm >> 5)) : 0u)) & 1u) == 0u)) {
                    if (kEarlyTma && last_pass && a.use_tma && lane == 0) mbar_arrive(&bars[kStages]);
                } else {
                    float2 v[16];
                    if (!kStaticMel && a.dither != 0.f) {      // generic kernels only; the static (LASR default) path has dither 0
                        const long long row = ((long long)utt * a.Tmax + f0 + fl) * a.win;
                        fc.ta = (unsigned)(f0 + fl); fc.tb = fc.ta + 1;
                        fc.noise_a = a.dither_noise ? a.dither_noise + row : nullptr;
                        fc.noise_b = a.dither_noise ? a.dither_noise + row + a.win : nullptr;
                    }
                    // sample offset of the frame inside the tile buffer (multi-utterance tiles: slot -> (utterance, frame))
                    int foff = fl * a.shift;
                    if (multi) { const int uj = fl / a.multi_fpu; foff = uj * a.multi_span + (fl - uj * a.multi_fpu) * a.shift; }
                    if (kDual) load_frame_dual<NLOAD, kPeak, !kStaticMel>(v, xs + foff + l, xs + foff + a.shift + l, wls, fc, l,
                                                                          fl + 1 < nvalid && (!multi || ((g.mask >> (fl + 1)) & 1u)));
                    else if (kI16) load_frame_single_i16<NLOAD, kPeak>(v, reinterpret_cast<const short*>(xs) + foff + 2 * l, wl, fc, l);
                    else load_frame_single<NLOAD, kPeak, !kStaticMel>(v, xs + foff + 2 * l, wl, fc, l);
                    if (kEarlyTma && last_pass && a.use_tma) {
                        // this warp no longer needs the tile buffer: once all warps say so, the next tile's TMA goes out
                        // (about half a tile earlier than after the phase-A barrier)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars[kStages]);
                    }
                    fft256_halfwarp(v, tw, xbuf, l);
                    float2 rc[8];
                    pair_exchange(v, rc, l, h2);
                    // power-spectrum row of this frame: the first frame of the half-warp behind its transposition area, the second
                    // on top of it (the FFT has left the area: fft256_halfwarp ends with __syncwarp after its last read)
                    float* prow = region + ((kDual || sub == 0) ? kXWords : 0);
                    float* pa = prow + l;                                                              // k = l + 16 r
                    if (kDual) {
                        // ---- separate the two real spectra; |2 Xa|^2 and |2 Xb|^2 (0.25 folded in the mel weights)
                        const bool bvalid = fl + 1 < nvalid && (!multi || ((g.mask >> (fl + 1)) & 1u));
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float2 bcj = make_float2(rc[r].x, -rc[r].y);
                            float2 S = add2(v[r], bcj), D = sub2(v[r], bcj);
                            S = mul2(S, S); D = mul2(D, D);
                            float pwa = S.x + S.y, pwb = D.x + D.y;
                            if (!use_power) { pwa = sqrtf(pwa); pwb = sqrtf(pwb); }
                            if (fvalid) pa[16 * r] = pwa;
                            if (bvalid) pa[16 * r - kXWords] = pwb;          // frame b: on top of the transposition area
                        }
                    } else {
                        // ---- real-FFT split + power: 2X[k] = S + T, 2 conj X[256-k] = S - T ----
                        float* pb = prow + 256 - l;                                                    // k = 256 - l - 16 r
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float2 bcj = make_float2(rc[r].x, -rc[r].y);
                            const float2 S = add2(v[r], bcj), D = sub2(v[r], bcj);
                            // W_32^r = exp(-j pi r / 16)
                            constexpr float kC32[8] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                                       0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
                            constexpr float kS32[8] = {0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                                                       -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
                            const float2 sw = (kStwRot && r > 0) ? c_mul(sw0, kC32[r], kS32[r]) : (kStwRot ? sw0 : stw[16 * r]);
                            const float2 T = c_mul(D, sw.x, sw.y);
                            float2 xa = add2(S, T), xb = sub2(S, T);
                            xa = mul2(xa, xa); xb = mul2(xb, xb);
                            float pwa = xa.x + xa.y, pwb = xb.x + xb.y;
                            if (!use_power) { pwa = sqrtf(pwa); pwb = sqrtf(pwb); }   // 2|X| (0.5 folded in weights)
                            if (fvalid) pa[16 * r] = pwa;
                            if (fvalid && (r != 0 || l != 0)) pb[-16 * r] = pwb;
                        }
                        if (l == 0 && fvalid) {   // bin 128 is its own partner: X[128] = conj Z[128]
                            const float2 z = v[8];
                            float p = 4.0f * (z.x * z.x + z.y * z.y);
                            if (!use_power) p = sqrtf(p);
                            prow[128] = p;                          // bin 128
                        }
                    }
                }
            }
            if (kEarlyTma && a.use_tma) {
                if (tid == 0) {
                    mbar_wait(&bars[kStages], consumed_phase);
                    if (nxt.id < ntiles) issue_load(gn, 0);
                }
                consumed_phase ^= 1u;
            }
            TL_STAMP(2);
            __syncthreads();   // B1: PT complete; every warp has finished phase C of the previous tile
            TL_STAMP(3);
            if (tid == 0) {
                if (!kEarlyTma && a.use_tma && nxt.id < ntiles) issue_load(gn, 0);
                if (kPipe) s_desc[it & 1] = make_int4(s_pipe[5], s_pipe[6], s_pipe[7], s_pipe[8]);     // read by everyone after B2
                else s_desc[it & 1] = make_int4(fut.id, fut.utt, fut.f0, fut.T);
                if (kApply) signal_pending(false);        // after the TMA issue: the fence must not delay the next tile's load
            }
            if (kApply) pending = utt;
            // per-tile epilogue tables for phase C (written here: no warp is still reading the previous tile's)
            if (cm_per_utt && tid < nmel) {
                s_mean[tid] = __ldg(a.cm_mean + (long long)utt * a.cm_stride + tid);
                s_istd[tid] = __ldg(a.cm_istd + (long long)utt * a.cm_stride + tid);
            }
            if (zmask && tid < kMaxMel) {
                const int* mk = a.masks + (long long)utt * nmask * 2;
                bool m = false;
#pragma unroll 1
                for (int i = 0; i < a.n_fmask; ++i) m |= (tid >= __ldg(mk + 2 * i) && tid < __ldg(mk + 2 * i + 1));
                const unsigned bal = __ballot_sync(0xffffffffu, m);
                if (lane == 0) s_cmask[warp] = bal;
            }
            if (zmask && warp == (kWarps > 4 ? 4 : 0)) {       // a warp without column work when there is one (128 columns = 4 warps)
                const int* mk = a.masks + (long long)utt * nmask * 2 + 2 * a.n_fmask;
                bool m = false;
#pragma unroll 1
                for (int i = 0; i < a.n_tmask; ++i) m |= (f0 + lane >= __ldg(mk + 2 * i) && f0 + lane < __ldg(mk + 2 * i + 1));
                const unsigned bal = __ballot_sync(0xffffffffu, m);
                if (lane == 0) s_cmask[4] = bal;
            }


            // ================= phase B: warp = mel-bin group, lane = frame =================
            {
                // 16-frame tiles (4 warps): the half-warps of a warp take two different bin groups
                const int fl = kWarps == 4 ? (lane & 15) : lane;
                const int grp = kWarps == 4 ? 2 * warp + (lane >> 4) : warp;
                float* orow = outs + fl * ostride;
                const float4* pcol = reinterpret_cast<const float4*>(pt + pt_row<kDual>(fl < kFT ? fl : 0));
                if (fl >= kFT) {
                    // 24-frame tiles leave lanes 24..31 idle in this phase
                } else if (kStaticMel) {
                    switch (grp) {
                        case 0: mel_static_group0(pcol, orow); break;
                        case 1: mel_static_group1(pcol, orow); break;
                        case 2: mel_static_group2(pcol, orow); break;
                        case 3: mel_static_group3(pcol, orow); break;
                        case 4: mel_static_group4(pcol, orow); break;
#if B200FE_WARPS == 8 || B200FE_WARPS == 4
                        case 5: mel_static_group5(pcol, orow); break;
                        case 6: mel_static_group6(pcol, orow); break;
                        default: mel_static_group7(pcol, orow); break;
#else
                        default: mel_static_group5(pcol, orow); break;
#endif
                    }
                } else {
                    const int jb = a.grp_begin[grp], je = a.grp_begin[grp + 1];
                    if (jb < je) {
                        const float* pf = pt + pt_row<kDual>(fl);
                        float up_prev = 0.f;
#pragma unroll 1
                        for (int s = jb; s <= je; ++s) {
                            const int kb = a.seg_start[s], ke = a.seg_start[s + 1];
                            float au = 0.f, ad = 0.f;
#pragma unroll 2
                            for (int k = kb; k < ke; ++k) {
                                const float p = pf[k];
                                const float2 w = a.w_updn[k];
                                au = fmaf(w.x, p, au);
                                ad = fmaf(w.y, p, ad);
                            }
                            if (s > jb) emit_bin(orow, s - 1, up_prev + ad);
                            up_prev = au;
                        }
                    }
                }
            }
            TL_STAMP(4);
            __syncthreads();   // B2: staging complete; PT free for the next tile's phase A
            TL_STAMP(5);
            nxt_desc = s_desc[it & 1];
        }

        // ================= phase C: epilogue + copy-out, zero padding, statistics =================
        // element e = row * nmel + col of the tile (contiguous in global memory) <-> staging row*(nmel+1)+col
        bool stats_fused = stats_cap;          // this tile's statistics are reduced inside the copy-out
        int tcls = 0, tcls_hi = 0;             // ... into these row classes (first and last row of the tile)
        if (has_cls && stats_cap && nvalid > 0) {
            const int* bounds = a.row_bounds + (long long)utt * (a.n_cls - 1);
            tcls = row_class(bounds, a.n_cls - 1, f0);
            tcls_hi = row_class(bounds, a.n_cls - 1, f0 + nvalid - 1);
            // a FULL tile that straddles class boundaries (one in eight on BASELINE config 3) stays on the register path too: the
            // thread adds its rows class by class; the write-back + barrier + column-reducer path below cost such a tile several
            // microseconds (fused launch 417 vs 354 us with and without row classes)
            stats_fused = tcls == tcls_hi || (kStaticMel && nvalid == kFT);
        }
        const bool rowpart_c = stats_fused || rowpart_nostats;
        {
            const long long orow0 = (!kLean && a.out_offsets != nullptr) ? __ldg(a.out_offsets + utt) : (long long)utt * a.Tmax;
            float* obase = a.out != nullptr ? a.out + (orow0 + f0) * nmel : nullptr;
            const int nv = nvalid * nmel, nt = nrows * nmel;
            if (nvalid > 0 && rowpart_c) {
                // Statistics mode without row classes: thread = (row part, column), element e = tid + k P with
                // P = parts * nmel, so the copy-out stays coalesced AND every thread keeps one column: its sum and
                // sum of squares (about a pivot, fp32 over <= 11 rows) never leave registers; one fp64 atomic pair
                // per thread and tile.  No write-back to the staging tile, no extra barrier.
                const bool affine = kLean ? false : (a.cm_mean != nullptr);
                const bool lg = a.use_log != 0;
                const float lf = a.log_floor;
                const int parts = kThreads / nmel, P = parts * nmel;
                if (tid < P) {
                    const int rg = tid / nmel, col = tid - rg * nmel;
                    const float* sp = outs + tid + rg;                  // staging index of e = tid: e + row
                    const float cm = affine ? s_mean[col] : 0.f, ci = affine ? s_istd[col] : 1.f;
                    // zero masks (replace_with_zero): the column bit is fixed per thread, row k of the thread is rg + k * parts
                    const bool colz = !kLean && zmask && ((s_cmask[col >> 5] >> (col & 31)) & 1u);
                    // rows of a ragged multi-utterance tile that hold no frame are written as zeros (the padded rows of their utterance)
                    const unsigned rowz = (((!kLean && zmask) ? (colz ? 0xffffffffu : s_cmask[4]) : 0u) | (multi ? ~g.mask : 0u)) >> rg;
                    float pivot = 0.f, s1 = 0.f, s2 = 0.f;
                    int cnt = 0;
                    if (kStaticMel && nvalid == kFT) {
                        constexpr int kIt = (kFT * B200FE_STATIC_NMEL + (kThreads / B200FE_STATIC_NMEL) * B200FE_STATIC_NMEL - 1) /
                                            ((kThreads / B200FE_STATIC_NMEL) * B200FE_STATIC_NMEL);
                        float x[kIt];
#pragma unroll
                        for (int k = 0; k < kIt; ++k) x[k] = (tid + k * P < nv) ? sp[k * (P + parts)] : 1.0f;
#pragma unroll
                        for (int k = 0; k < kIt; ++k) {
                            if (lg) x[k] = fast_log(fmaxf(x[k], lf));
                            if (affine) x[k] = (x[k] - cm) * ci;
                            if (!kLean && ((rowz >> (k * parts)) & 1u)) x[k] = 0.f;
                        }
                        pivot = x[0];
                        if (stats_fused) {
#pragma unroll
                            for (int k = 0; k < kIt; ++k) {
                                if (tid + k * P < nv) {
                                    if (obase) obase[tid + k * P] = x[k];
                                    const float dd = x[k] - pivot;
                                    s1 += dd; s2 = fmaf(dd, dd, s2); ++cnt;
                                }
                            }
                            if (!kLean && tcls != tcls_hi) {
                                // several row classes in this tile: the sums of squares go to the one total as usual, the sums
                                // are split by class -- the thread's rows are rg + k * parts, a class is a row interval
                                double* sb = a.stats + (long long)utt * a.stats_stride;
                                const int* bounds = a.row_bounds + (long long)utt * (a.n_cls - 1);
                                const double dp = (double)pivot;
                                for (int c = tcls; c <= tcls_hi; ++c) {
                                    const int lo = (c == 0 ? 0 : __ldg(bounds + c - 1)) - f0, hi = (c == a.n_cls - 1 ? 0x3fffffff : __ldg(bounds + c)) - f0;
                                    float sc = 0.f; int nc = 0;
#pragma unroll
                                    for (int k = 0; k < kIt; ++k) {
                                        const int row = rg + k * parts;
                                        if (tid + k * P < nv && row >= lo && row < hi) { sc += x[k] - pivot; ++nc; }
                                    }
                                    if (nc > 0) atomicAdd(sb + (long long)c * nmel + col, (double)sc + nc * dp);
                                }
                                const double d1 = (double)s1;
                                atomicAdd(sb + (long long)a.n_cls * nmel + col, (double)s2 + 2.0 * dp * d1 + cnt * dp * dp);
                                cnt = 0;                         // flushed
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < kIt; ++k) if (tid + k * P < nv) obase[tid + k * P] = x[k];
                        }
                    } else {
#pragma unroll 2
                        for (int e = tid, k = 0; e < nv; e += P, ++k) {
                            float x = sp[k * (P + parts)];
                            if (lg) x = fast_log(fmaxf(x, lf));
                            if (affine) x = (x - cm) * ci;
                            if (!kLean && ((rowz >> (k * parts)) & 1u)) x = 0.f;
                            if (obase) obase[e] = x;
                            if (cnt == 0) pivot = x;
                            const float dd = x - pivot;
                            s1 += dd; s2 = fmaf(dd, dd, s2); ++cnt;
                        }
                    }
                    if (stats_fused && cnt > 0) {
                        double* sb = a.stats + (long long)utt * a.stats_stride;
                        const double dp = (double)pivot, d1 = (double)s1;
                        atomicAdd(sb + (long long)tcls * nmel + col, d1 + cnt * dp);
                        atomicAdd(sb + (long long)a.n_cls * nmel + col, (double)s2 + 2.0 * dp * d1 + cnt * dp * dp);
                    }
                }
            } else if (nvalid > 0) {
                const bool wb = a.stats != nullptr;       // statistics read the transformed values back
                const bool affine = kLean ? false : (a.cm_mean != nullptr);
                const float lf = a.log_floor;
                if (kStaticMel && a.use_log != 0 && !zmask && !affine && !wb && obase != nullptr && nvalid == kFT) {
                    // full tile, plain log-mel output: ten independent load -> log -> store chains per thread
                    constexpr int kPer = kFT * B200FE_STATIC_NMEL / kThreads;
                    float x[kPer];
#pragma unroll
                    for (int i = 0; i < kPer; ++i) { const int e = tid + i * kThreads; x[i] = outs[e + e / B200FE_STATIC_NMEL]; }
#pragma unroll
                    for (int i = 0; i < kPer; ++i) {
                        float y = fast_log(fmaxf(x[i], lf));
                        if (multi && !((g.mask >> ((tid + i * kThreads) / B200FE_STATIC_NMEL)) & 1u)) y = 0.f;
                        obase[tid + i * kThreads] = y;
                    }
                } else if (kStaticMel && a.use_log != 0 && !zmask && !affine && wb && nvalid == kFT) {
                    // full tile in statistics mode (utterance CMVN is applied by the post pass): same ten chains, the
                    // log-mel values are also written back for the column reducers
                    constexpr int kPer = kFT * B200FE_STATIC_NMEL / kThreads;
                    float x[kPer];
#pragma unroll
                    for (int i = 0; i < kPer; ++i) { const int e = tid + i * kThreads; x[i] = outs[e + e / B200FE_STATIC_NMEL]; }
#pragma unroll
                    for (int i = 0; i < kPer; ++i) {
                        const int e = tid + i * kThreads;
                        x[i] = fast_log(fmaxf(x[i], lf));
                        if (obase) obase[e] = x[i];
                        outs[e + e / B200FE_STATIC_NMEL] = x[i];
                    }
                } else if (a.use_log != 0 && !zmask) {
                    // fast path: element e = row * nmel + col <-> staging e + row
                    if (affine) {
#pragma unroll 2
                        for (int e = tid; e < nv; e += kThreads) {
                            const int row = e / nmel, col = e - row * nmel;
                            float* sp = outs + e + row;
                            float x = fast_log(fmaxf(*sp, lf));
                            x = (x - s_mean[col]) * s_istd[col];
                            if (multi && !((g.mask >> row) & 1u)) x = 0.f;
                            if (obase) obase[e] = x;
                            if (wb) *sp = x;
                        }
                    } else {
#pragma unroll 2
                        for (int e = tid; e < nv; e += kThreads) {
                            const int row = e / nmel;
                            float* sp = outs + e + row;
                            float x = fast_log(fmaxf(*sp, lf));
                            if (multi && !((g.mask >> row) & 1u)) x = 0.f;
                            if (obase) obase[e] = x;
                            if (wb) *sp = x;
                        }
                    }
                } else {
                    const unsigned rmask = zmask ? s_cmask[4] : 0u;
                    const bool lg = a.use_log != 0;
#pragma unroll 1
                    for (int e = tid; e < nv; e += kThreads) {
                        const int row = e / nmel, col = e - row * nmel;
                        float* sp = outs + e + row;
                        float x = *sp;
                        if (lg) x = fast_log(fmaxf(x, lf));
                        if (affine) x = (x - s_mean[col]) * s_istd[col];
                        if (zmask && (((rmask >> row) & 1u) || ((s_cmask[col >> 5] >> (col & 31)) & 1u))) x = 0.f;
                        if (multi && !((g.mask >> row) & 1u)) x = 0.f;
                        if (obase) obase[e] = x;
                        if (wb) *sp = x;
                    }
                }
            }
            // rows past the utterance end are zero (pad_audio = 0)
            if (obase && nt > nv) {
                if ((nmel & 3) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0) {
                    float4* z4 = reinterpret_cast<float4*>(obase);
                    for (int q = (nv >> 2) + tid; q < (nt >> 2); q += kThreads) z4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    for (int e = nv + tid; e < nt; e += kThreads) obase[e] = 0.f;
                }
            }
        }
        if (multi) {
            if (a.out_len != nullptr && tid * a.multi_fpu < nvalid)
                a.out_len[utt + tid] = __popc((g.mask >> (tid * a.multi_fpu)) & ((1u << a.multi_fpu) - 1u));
        } else if (a.out_len != nullptr && f0 == 0 && nrows > 0 && tid == 0) a.out_len[utt] = g.T;
        if (a.stats != nullptr && nvalid > 0 && !stats_fused) {
            // Column statistics of what phase C wrote back into the staging tile.  thread = (column j, row
            // part); fp32 partial sums over <= 11 rows, flushed with fp64 atomics.
            __syncthreads();
            const int nb = a.n_cls - 1;
            const int* bounds = (a.row_bounds != nullptr && nb > 0) ? a.row_bounds + (long long)utt * nb : nullptr;
            double* sb = a.stats + (long long)utt * a.stats_stride;
            const int parts = max(1, min(kThreads / nmel, 4));
            const int part = tid / nmel, j = tid - part * nmel;
            if (part < parts) {
                const int rb = (nvalid * part) / parts, re = (nvalid * (part + 1)) / parts;
                if (re > rb) {
                    int cls = bounds ? row_class(bounds, nb, f0 + rb) : 0;
                    // sums are taken about a pivot (the part's first value) so that fp32 keeps ~7 digits of the
                    // spread, not of the offset; converted back in fp64: sum = s1 + n p, sumsq = s2 + 2 p s1 + n p^2
                    float s1 = 0.f, s2 = 0.f, pivot = 0.f;
                    int cnt = 0;
                    double q2 = 0.0;
                    for (int fr = rb; fr < re; ++fr) {
                        if (bounds) {
                            const int cc = row_class(bounds, nb, f0 + fr);
                            if (cc != cls) {
                                const double dp = (double)pivot, d1 = (double)s1;
                                atomicAdd(sb + (long long)cls * nmel + j, d1 + cnt * dp);
                                q2 += (double)s2 + 2.0 * dp * d1 + cnt * dp * dp;
                                s1 = 0.f; s2 = 0.f; cnt = 0; cls = cc;
                            }
                        }
                        const float x = outs[fr * ostride + j];
                        if (fr == rb) pivot = x;
                        const float d = x - pivot;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                        ++cnt;
                    }
                    const double dp = (double)pivot, d1 = (double)s1;
                    atomicAdd(sb + (long long)cls * nmel + j, d1 + cnt * dp);
                    atomicAdd(sb + (long long)a.n_cls * nmel + j, q2 + (double)s2 + 2.0 * dp * d1 + cnt * dp * dp);
                }
            }
        }
        TL_STAMP(6);
#ifdef B200FE_TIMELINE
        if (lane == 0 && it < kTlTiles && blockIdx.x < 296) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); g_timeline[blockIdx.x][warp][it][7] = ((long long)smid << 32) | (unsigned)nvalid; }
#endif
        // the staging area aliases the transposition buffers that the next phase A overwrites
        // No closing barrier: a warp that finishes phase C goes straight to the next tile's phase A.  (A tile
        // without frames has no B1 / B2, so the descriptor exchange gets its own barrier.)
        if (nvalid <= 0) {
            if (tid == 0) {
                if (kPipe) s_desc[it & 1] = make_int4(s_pipe[5], s_pipe[6], s_pipe[7], s_pipe[8]);
                else s_desc[it & 1] = make_int4(fut.id, fut.utt, fut.f0, fut.T);
            }
            __syncthreads();
            nxt_desc = s_desc[it & 1];
            if (kApply) {
                // an apply tile may depend on completions this CTA still holds back: publish them before waiting
                if (tid == 0) signal_pending(g.apply);
                pending = -1;
                if (g.apply) {
                    static_assert(kWarps * kRegionWarp * 4 + ((kFT * (B200FE_STATIC_NMEL + 1) * 4 + 15) & ~15) >= kApplyRows * B200FE_STATIC_NMEL * 4,
                                  "an apply tile must fit the transposition / PT regions + the staging tile");
                    if (!kLean) apply_mask_tile(a, g.utt, g.T, g.ready, smem + L.xbuf_off, tid, nmel);      // apply_mode 3 (the lean kernel has no masks)
                    else if (apply_cmvn_tile(a, g.utt, g.f0, g.T, g.ready, s_mean, s_istd, smem + L.xbuf_off, &bars[2], (phase_bits >> 1) & 1u, tid, nmel))
                        phase_bits ^= 2u;
                }
            }
            __syncthreads();
        } else if (kAliasStaging) {
            __syncthreads();      // the staging tile aliases the transposition buffers of the next phase A
        }
        cur = nxt;
        g = gn;
        nxt = Desc{nxt_desc.x, nxt_desc.y, nxt_desc.z, nxt_desc.w};
    }
    if (kApply) {
        __syncthreads();
        if (tid == 0) signal_pending(true);
    }
}

}  // namespace b200fe

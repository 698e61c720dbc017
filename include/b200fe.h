/* b200fe -- C ABI of the B200 (sm_100a) acoustic front end: Kaldi-style 80-dim log-mel fbank
 * + CMVN + SpecAugment masks + zero-padded batch layout, for gaochangfeng/lighting-asr (LASR).
 *
 * Every entry point replaces a piece of the reference's per-utterance CPU path; citations use
 * R/ = the LASR checkout and TA: = torchaudio/compliance/kaldi.py (torchaudio 2.11.0), the
 * library R/lasr/data/datatrans.py:75-102 delegates to.
 *
 * Conventions: all functions return 0 on success or a negative status (B200FE_E*); the text of
 * the last failure on the calling thread is returned by b200fe_last_error().  Pointers named
 * d_* are DEVICE pointers owned by the caller (e.g. the torch allocator); the library never
 * allocates per call, never synchronises the stream and keeps no global state: a plan holds a
 * few kB of read-only device tables and may be shared by any number of streams/threads.
 * `stream` is a cudaStream_t passed as void*.
 */
#ifndef B200FE_H
#define B200FE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FE_OK 0
#define B200FE_EINVAL (-1)      /* bad argument / unsupported option combination */
#define B200FE_ECUDA (-2)       /* CUDA runtime error, see b200fe_last_error() */
#define B200FE_ESHORT (-3)      /* an utterance is shorter than one window (TA:142 asserts) */

#define B200FE_MAX_MEL 128
#define B200FE_MAX_FREQ_MASKS 4
#define B200FE_MAX_TIME_MASKS 4

typedef struct b200fe_plan b200fe_plan;

/* Options = the keyword arguments of WavToKaldiFbank (R/lasr/data/datatrans.py:43-71) that
 * reach torchaudio.compliance.kaldi.fbank (TA:514-541).  Not supported (rejected with EINVAL):
 * snip_edges=False, use_energy=True, vtln_warp != 1, htk_compat (no effect without energy). */
typedef struct b200fe_opts {
    float sample_frequency;        /* 16000 */
    float frame_length_ms;         /* 25 */
    float frame_shift_ms;          /* 10 */
    int num_mel_bins;              /* 80 */
    float low_freq;                /* 20 */
    float high_freq;               /* 0 (= Nyquist offset, TA:457-458) */
    float preemphasis_coefficient; /* 0.97 */
    int remove_dc_offset;          /* 1 */
    int use_power;                 /* 1 */
    int use_log_fbank;             /* 1 */
    int window_type;               /* 0 povey, 1 hanning, 2 hamming, 3 rectangular, 4 blackman (TA:86-113) */
    float blackman_coeff;          /* 0.42 */
    int audio_bit;                 /* 16: waveform is scaled by 2^(audio_bit-1) (datatrans.py:74) */
    float dither;                  /* 0 (LASR default, datatrans.py:47); else x[f][j] += dither * N(0,1) per (frame, sample), TA:179-181 */
    /* Optional host tables that override the built-in (double precision, rounded once)
     * constants so that they match the caller's torch build bit for bit.  May be NULL. */
    const float* window;           /* [window_size] */
    const float* mel_weights;      /* [num_mel_bins][padded_window_size/2] row-major (TA:436-511) */
} b200fe_opts;

void b200fe_default_opts(b200fe_opts* o);
int b200fe_plan_create(const b200fe_opts* opts, b200fe_plan** plan);
void b200fe_plan_destroy(b200fe_plan* plan);
const char* b200fe_last_error(void);

/* Host helpers: TA:125-151 window properties and the snip_edges frame count TA:63-67. */
int b200fe_window_size(const b200fe_plan* plan);
int b200fe_window_shift(const b200fe_plan* plan);
int b200fe_padded_window_size(const b200fe_plan* plan);
long long b200fe_num_frames(const b200fe_plan* plan, long long num_samples);
/* Introspection for tests / benchmarks: what = 0 straight-line mel path in use, 1 registers
 * loaded per lane (13|16), 2 dynamic shared memory per CTA, 3 resident CTAs per SM, 4 SM count, 5 frames
 * per tile (the granularity of b200fe_build_tile_table), 6 experimental warp-specialised kernel in use, 7 in-launch utterance CMVN
 * (apply tiles) available, 8 rows per CMVN-apply tile, 9 warps per CTA. */
int b200fe_plan_info(const b200fe_plan* plan, int what);

/* Host helper: fills table[2*i] = utterance, table[2*i+1] = first frame for every tile (b200fe_plan_info(plan, 5) frames)
 * with at least one valid frame, in utterance order; returns the number of tiles (table may be NULL
 * to query the size), or a negative status. */
int b200fe_build_tile_table(const b200fe_plan* plan, const long long* nsamp_host, int batch, int* table_host, int capacity);
/* The same list plus PADDING tiles: an entry (utterance, -(row0 + 1)) makes the fused launch zero the rows
 * [row0, row0 + 256) of that utterance (clipped at max_frames), so the zero padding of the reference's collate
 * (lasr/data/dataset.py:18) is written by the same persistent CTAs, interleaved with the frame tiles, instead of by a
 * separate streaming kernel.  Pass the result with b200fe_fbank_args.tile_table_pads != 0. */
int b200fe_build_tile_table_padded(const b200fe_plan* plan, const long long* nsamp_host, int batch, int max_frames, int* table_host, int capacity);
/* The same work list built ON THE DEVICE from d_nsamp (one small kernel): no host-side table, no table upload.  d_table has room
 * for `capacity` entries (b200fe_tile_table_capacity gives a sufficient bound), *d_n_tiles receives the number of entries and
 * *d_work_counter is reset.  Pass d_table, n_tiles = capacity, d_n_tiles and d_work_counter to b200fe_fbank_fused (with
 * tile_table_pads = with_pads). */
int b200fe_tile_table_capacity(const b200fe_plan* plan, int batch, int max_frames, int with_pads);
int b200fe_build_tile_table_device(const b200fe_plan* plan, const long long* d_nsamp, int batch, int max_frames, int with_pads,
                                   int* d_table, int capacity, int* d_n_tiles, int* d_work_counter, void* stream);
/* The device-side builder with two options.  (1) d_zero / zero_bytes (optional, 16-byte aligned, multiple of 16): a buffer
 * the same kernel clears -- the statistics accumulators of the call -- which saves the memset in front of the fused launch.
 * (2) apply_lag > 0: CMVN-APPLY tiles (utterance CMVN inside the fused launch instead of b200fe_postpass, for the
 * `_subtract_column_mean` semantics of TA:220-226,644 and its mean/variance extension): besides the frame and padding tiles of
 * utterance u, slot u of the list carries the apply tiles of utterance u - apply_lag, entries (utterance | 0x40000000, row0) =
 * normalise rows [row0, row0 + R) of that utterance (R = b200fe_plan_info(plan, 8)) in place; the last apply_lag utterances' apply tiles close the list.  A
 * persistent CTA that takes an apply tile waits until all frame tiles of that utterance have been signalled in
 * d_utt_done[utterance] (int32 [batch + 1], zeroed by this call; element [batch] becomes non-zero if an apply tile ever gave up
 * waiting -- never expected, checked by the tests), reads the utterance's column sums from d_stats and rewrites the rows while
 * they are still L2-resident.  64 is a good lag: a CTA publishes the completions of eight tiles behind one fence and the
 * persistent grid holds about 900 claimed tiles, so the frame tiles an apply tile depends on have normally been signalled when it
 * is claimed.  Capacity with apply tiles: b200fe_tile_table_capacity(...) + batch * ceil(max_frames / R).  Pass d_utt_done and
 * apply_cmvn_mode to b200fe_fbank_fused.  Needs b200fe_plan_info(plan, 7) != 0.  Measured on B200 (BASELINE config 2): 0.380 ms
 * per step against 0.382 ms with b200fe_postpass -- opt-in, see DESIGN.md 5.3.
 * (3) apply_lag < 0: ONE completion tile per utterance (entry (utterance | 0x40000000, 0)), |apply_lag| utterances behind its frame
 * tiles, for apply_cmvn_mode 3 of b200fe_fbank_fused (SpecAugment mean fills inside the launch); capacity + batch. */
int b200fe_build_work_list_device(const b200fe_plan* plan, const long long* d_nsamp, int batch, int max_frames, int with_pads, int apply_lag,
                                  int* d_table, int capacity, int* d_n_tiles, int* d_work_counter, int* d_utt_done,
                                  void* d_zero, long long zero_bytes, void* stream);

/* Peak normalisation statistics: d_peak[b] = max |wav[b][0..nsamp[b])|.
 * Replaces the abs-max half of VoiceNorm (R/lasr/data/datatrans.py:22-27); the division is
 * fused into b200fe_fbank_fused through its `d_peak` argument. */
int b200fe_peak_absmax(const b200fe_plan* plan, const float* d_wav, long long wav_stride, const long long* d_wav_offsets,
                       const long long* d_nsamp, int batch, float* d_peak, void* stream);
/* Same for int16 PCM input: d_peak[b] = max |s16| / 2^15, the value max |x| takes after the reader's
 * int16 -> float conversion (R/lasr/data/reader.py:24, soundfile). */
int b200fe_peak_absmax_i16(const b200fe_plan* plan, const short* d_wav, long long wav_stride, const long long* d_wav_offsets,
                           const long long* d_nsamp, int batch, float* d_peak, void* stream);

/* One fused launch over a zero-padded batch of waveforms.
 * Replaces WavToKaldiFbank (R/lasr/data/datatrans.py:42-104 -> TA:514-645) for every
 * utterance of the batch and the padding collate batch_list (R/lasr/data/dataset.py:8-22):
 * features are written straight into the (batch, max_frames, num_mel_bins) layout, rows past an
 * utterance's frame count are zero (pad_audio = 0, R/example/asr_en/conf/config_baseline.yaml:70). */
typedef struct b200fe_fbank_args {
    /* = sizeof(b200fe_fbank_args) of the header the caller was built against: a binding whose struct is shorter or longer than
     * the library's (a stale copy of this declaration) is rejected with B200FE_EINVAL instead of being read past its end.
     * The same first field opens b200fe_post_args and b200fe_warp_args. */
    unsigned int struct_size;
    const float* d_wav;          /* [batch][wav_stride] float32 */
    long long wav_stride;        /* elements between utterances */
    const long long* d_nsamp;    /* [batch] valid samples */
    int batch;
    const float* d_peak;         /* NULL, or [batch] from b200fe_peak_absmax: x / (peak + 1e-9) first */
    float* d_out;                /* [batch][max_frames][num_mel_bins]; NULL = statistics only */
    long long* d_out_len;        /* optional [batch]: frames per utterance (the wav_len tensor, dataset.py:198,206) */
    int max_frames;              /* rows per utterance in d_out (>= max_b frames(nsamp[b]) is the caller's job) */
    /* CMVN applied in the epilogue, (x - mean) * istd in float32.  NULL = none.
     * cmvn_stride 0 = one global vector pair, num_mel_bins = per utterance. */
    const float* d_cmvn_mean;
    const float* d_cmvn_istd;
    long long cmvn_stride;
    /* SpecAugment rectangles per utterance: [n_freq_masks + n_time_masks][2] int32 (start, stop),
     * frequency masks first (R/lasr/utils/specaugment.py:47-106).  mask_zero = 1 zeroes them in this
     * launch (replace_with_zero=True); mask_zero = 0 leaves them to b200fe_postpass, which computes the mean fills. */
    const int* d_masks;
    int n_freq_masks, n_time_masks;
    int mask_zero;
    /* Column statistics of what is written (after CMVN), accumulated with fp64 atomics:
     * d_stats[u * stats_stride + c * num_mel_bins + d]: c < n_row_classes = sums over the rows of
     * class c, c == n_row_classes = sums of squares over all rows.  stats_stride 0 = one
     * accumulator for the whole batch (global CMVN, Kaldi compute-cmvn-stats).  Row classes are
     * the elementary intervals of d_row_bounds[u][n_row_classes-1] (sorted); NULL = one class.
     * The caller zeroes d_stats. */
    double* d_stats;
    long long stats_stride;
    const int* d_row_bounds;
    int n_row_classes;
    /* Optional compact work list for ragged batches (b200fe_build_tile_table): [n_tiles][2] int32
     * (utterance, first frame) of every tile that holds valid frames, consumed through the
     * atomic counter d_work_counter (4 bytes, reset by the call).  Padded rows are then zeroed by a
     * separate streaming kernel of the same call.  NULL = iterate the padded (batch x max_frames) grid. */
    const int* d_tile_table;
    int n_tiles;
    int* d_work_counter;
    /* Optional packed (ragged) input: utterance u starts at d_wav + d_wav_offsets[u] (elements) instead of
     * d_wav + u * wav_stride; wav_stride must still bound the longest utterance.  offsets_aligned != 0
     * promises that every offset is a multiple of 4 samples (enables the TMA loader). */
    const long long* d_wav_offsets;
    int offsets_aligned;
    /* Dither (plan option dither != 0): standard-normal noise per (utterance, frame, sample).  By default it
     * comes from a counter-based Philox4x32-10 generator keyed by dither_seed (use a fresh seed per call);
     * d_dither_noise [batch][max_frames][window_size] replaces it, e.g. with torch.randn drawn under the same
     * torch.manual_seed as the reference ("seeded identically"). */
    unsigned long long dither_seed;
    const float* d_dither_noise;
    /* Waveform element type: 0 = float32 in [-1, 1) (what the reader returns), 1 = int16 PCM (what the file
     * holds; d_wav then points to int16 and strides / offsets count int16 samples, offsets multiples of 8
     * for the TMA loader).  (float)s16 == float32 sample * 2^15 exactly, so both give identical features. */
    int wav_dtype;
    /* Lock-step streaming: non-zero promises that EVERY utterance of this call yields exactly max_frames frames
     * (d_nsamp[u] in [(max_frames-1)*shift + window, max_frames*shift + window) for all u).  With max_frames <= 16 the
     * fused kernel then packs several utterances into one 32-frame tile (8 streams x 4 frames for a 40 ms push).
     * Ignored together with peak normalisation, statistics, masks, per-utterance CMVN, dither or a tile table.
     * 2 = the weaker promise "AT MOST max_frames frames each" (independent streams, b200fe_stream_push): tiles are shared the
     * same way, frame slots without a frame are masked and their output rows are zero. */
    int uniform_frames;
    /* Packed feature output: [batch] first output ROW of every utterance, ascending; d_out is then
     * [sum of frames][num_mel_bins] with no padding rows (utterance u occupies rows [d_out_offsets[u],
     * d_out_offsets[u] + frames[u])).  NULL = the reference's zero-padded [batch][max_frames][num_mel_bins] layout. */
    const long long* d_out_offsets;
    /* Non-zero: d_tile_table also holds the padding tiles of b200fe_build_tile_table_padded; the separate zero-fill
     * kernel is then skipped. */
    int tile_table_pads;
    /* Optional: the number of valid entries of d_tile_table lives in device memory (b200fe_build_tile_table_device); n_tiles
     * is then only an upper bound used to size the grid, and the work counter is NOT reset by this call. */
    const int* d_n_tiles;
    /* Utterance CMVN applied INSIDE the launch by the apply tiles of b200fe_build_work_list_device: 0 = off, 1 = subtract the
     * utterance's column means (torchaudio's subtract_mean, TA:220-226), 2 = mean and variance ((x - mean) / std, variance
     * floored at 1e-20: the definition of b200fe_postpass).  Needs d_stats with stats_stride > 0 (per-utterance sums, one row
     * class), d_utt_done (int32 [batch + 1], zeroed by the list builder), the padded output layout and the default option set
     * (b200fe_plan_info(plan, 7) != 0); d_utt_mean / d_utt_istd (optional, [batch][num_mel_bins]) receive the vectors used. */
    int apply_cmvn_mode;
    int* d_utt_done;
    float* d_utt_mean;
    float* d_utt_istd;
    /* apply_cmvn_mode 3: SpecAugment MEAN FILLS inside the launch (the reference's default masks, replace_with_zero=False, on top
     * of global or no CMVN) by the completion tiles of b200fe_build_work_list_device(apply_lag < 0): the CTA that takes an
     * utterance's completion tile derives the fills from the row-class statistics as b200fe_postpass does and overwrites the
     * masked cells, so no b200fe_postpass call follows.  Needs d_masks with mask_zero = 0, per-utterance d_stats with the row
     * classes of the time masks (d_row_bounds), d_utt_done and the default option set; d_fills (optional,
     * [batch][n_freq_masks + n_time_masks]) receives the fills. */
    float* d_fills;
} b200fe_fbank_args;

int b200fe_fbank_fused(const b200fe_plan* plan, const b200fe_fbank_args* args, void* stream);

/* ---- Streaming front end (BASELINE config 5): S INDEPENDENT audio streams per handle.  The reference has no
 * streaming fbank -- ASRProcess.frontend (R/lasr/process/asrprocess.py:49-56) transforms whole utterances and its "online"
 * encoders chunk finished features (R/lasr/modules/net/online_transformer/encoder.py:143-176) -- so the contract is: the
 * concatenation of a stream's push outputs equals the offline fbank of everything pushed into it, frame for frame.
 * The handle owns the per-stream state on the device (the window - shift ... window - 1 samples the next frames still need),
 * allocated once by b200fe_stream_create on the current device; a push allocates nothing, never synchronises and reads no
 * length on the host, so a push with fixed pointers can be captured in a CUDA graph and replayed.  Streams of one push may
 * carry chunks of DIFFERENT lengths (0 = nothing arrived for that stream); different handles / CUDA streams are independent. */
typedef struct b200fe_stream b200fe_stream;
int b200fe_stream_create(const b200fe_plan* plan, int n_streams, int max_chunk /* samples per push and stream */, b200fe_stream** out);
void b200fe_stream_destroy(b200fe_stream* st);
/* Upper bound of the frames one push can complete for one stream: size d_out with it. */
int b200fe_stream_max_frames(const b200fe_stream* st);
/* Forgets the carry of the listed streams (d_ids device int32 [n]; NULL = streams 0 .. n-1). */
int b200fe_stream_reset(b200fe_stream* st, const int* d_ids, int n, void* stream);
/* Row i of the push feeds stream d_ids[i] (NULL: stream i) with d_chunk_len[i] samples (NULL: max_chunk for all) from
 * d_chunks + i * chunk_stride (float32, device).  Output row i: d_out[i][max_out_frames][num_mel_bins], the d_out_frames[i]
 * completed frames first, zero rows after them; optional global CMVN as in b200fe_fbank_fused.  Each stream may appear once
 * per push. */
int b200fe_stream_push(b200fe_stream* st, const int* d_ids, int n, const float* d_chunks, long long chunk_stride, const int* d_chunk_len,
                       const float* d_cmvn_mean, const float* d_cmvn_istd, float* d_out, int max_out_frames, long long* d_out_frames, void* stream);
/* Sticky error bits of the pushes so far (synchronises the given stream): 1 a chunk was longer than max_chunk and was
 * clipped, 2 a push completed more frames than max_out_frames, 4 a stream id was out of range. */
int b200fe_stream_flags(b200fe_stream* st, int* h_flags, void* stream);

/* Host -> device staging of a zero-padded HOST batch (what batch_list builds, dataset.py:8-22) into the
 * packed device layout: only the valid samples of every utterance cross PCIe (one cudaMemcpyAsync per
 * utterance on `stream`; h_wav should be pinned).  h_offsets[u] = destination offset (elements). */
int b200fe_h2d_ragged(const void* h_wav, long long h_stride, const long long* h_nsamp, const long long* h_offsets,
                      int batch, void* d_packed, int elem_bytes /* 4 float32, 2 int16 */, void* stream);

/* Device -> host copy of the first h_rows[u] rows of every utterance of a padded [batch][utt_rows][row_elems]
 * feature tensor into a host tensor of the same layout (one cudaMemcpyAsync per utterance).  The caller
 * passes, per utterance, the number of rows that can differ from what the host buffer already holds
 * (frames now valid, or valid in the previous batch and zero now), so zero padding is not re-sent. */
int b200fe_d2h_ragged(const float* d_feats, long long row_elems, long long utt_rows, const long long* h_rows, int batch,
                      float* h_feats, void* stream);

/* ---- Host staging (no device work): the reference's collate loop receives a LIST of 1-D host arrays, one per utterance
 * (R/lasr/data/dataset.py:190-206; float64 from soundfile.read, R/lasr/data/reader.py:24).  A persistent thread pool packs
 * them back to back into ONE (pinned) staging buffer -- the layout b200fe_fbank_fused reads through d_wav_offsets -- converting
 * float64 -> float32 on the way (the `.float()` of WavToKaldiFbank, R/lasr/data/datatrans.py:73), and zero-fills the padding
 * rows of the host feature batch (pad_audio = 0, dataset.py:18) while the GPU works.  Jobs are asynchronous: *_begin returns a
 * positive ticket, b200fe_host_wait blocks until that job is done (the waiting thread helps).  Tickets complete in FIFO order
 * of submission per pool. */
typedef struct b200fe_host_pool b200fe_host_pool;
int b200fe_host_pool_create(int n_threads /* <= 0: one per hardware thread */, b200fe_host_pool** pool);
void b200fe_host_pool_destroy(b200fe_host_pool* pool);
int b200fe_host_pool_threads(const b200fe_host_pool* pool);
/* Instruction set of the packing / zero-fill loops, chosen once at load time: 0 = SSE2, 1 = AVX-512F (64-byte loads, whole-line
 * non-temporal stores, L2 prefetch ahead of the loads; B200FE_HOST_ISA=sse2 in the environment forces the baseline). */
int b200fe_host_isa(void);
/* Data pointers of n NumPy arrays from the addresses of their Python objects (id(array)): the collate loop hands over a LIST of
 * ndarrays (R/lasr/data/dataset.py:190-206) and asking each one for its address through the interpreter costs ~2 us per utterance
 * -- half a millisecond of a C2 call before the first packing job can start.  data_offset = offset of PyArrayObject's `data` field
 * (the first field after the object header: 16 on 64-bit CPython); the caller verifies it on a probe array before relying on it. */
int b200fe_host_ndarray_data(const long long* py_objects, int n, void** data_out, int data_offset);
/* Utterance u: nsamp[u] elements from h_src[u] go to h_dst + dst_offsets[u] (elements of the DESTINATION type; starts must be
 * 16-byte aligned); the gap up to the next 16-byte boundary is cleared.  src_dtype: 0 float32 -> float32, 1 int16 -> int16
 * (PCM as the file holds it), 2 float64 -> float32, 3 float64 that HOLDS 16-bit PCM values (k / 32768, what
 * soundfile.read returns for PCM_16 files) -> int16 k: half the staging and PCIe bytes, bit-identical features; a sample that is
 * not such a value raises the job's flag (b200fe_host_wait_flag) and the caller repeats the batch with src_dtype 2.
 * dst_capacity = elements h_dst can hold. */
long long b200fe_host_pack_begin(b200fe_host_pool* pool, const void* const* h_src, const long long* nsamp, int batch, int src_dtype,
                                 void* h_dst, const long long* dst_offsets, long long dst_capacity);
/* b200fe_host_pack_begin plus the upload: the pool thread that finishes the job's last task issues ONE
 * cudaMemcpyAsync(d_dst + dst_offsets[0], h_dst + dst_offsets[0], copy_elems elements, copy_stream) on `device` and records
 * `copy_event` (a cudaEvent_t, optional) behind it, before the ticket completes -- so the DMA of a group starts the moment its
 * packing ends, independent of what the calling thread is doing.  h_dst must be pinned. */
long long b200fe_host_pack_copy_begin(b200fe_host_pool* pool, const void* const* h_src, const long long* nsamp, int batch, int src_dtype,
                                      void* h_dst, const long long* dst_offsets, long long dst_capacity,
                                      void* d_dst, long long copy_elems, int device, void* copy_stream, void* copy_event);
/* Clears rows [valid_rows[u], utt_rows) of every utterance of a host [batch][utt_rows][row_elems] float32 tensor. */
long long b200fe_host_zero_rows_begin(b200fe_host_pool* pool, float* h_feats, int batch, long long utt_rows, long long row_elems,
                                      const long long* valid_rows, int elem_bytes /* 4 float32, 2 bfloat16 */);
/* Clears n byte ranges [offsets[i], offsets[i] + nbytes[i]) of a host buffer (non-temporal stores): the parts of a recycled
 * batch buffer that still hold an earlier batch's rows and fall into this batch's padding. */
long long b200fe_host_zero_ranges_begin(b200fe_host_pool* pool, void* h_base, const long long* offsets, const long long* nbytes, int n);
int b200fe_host_wait(b200fe_host_pool* pool, long long ticket);
/* b200fe_host_wait that also returns the job's flag: 1 when a task of a src_dtype 3 packing job met a sample that is not a PCM16
 * value (the staged data of that job is then unusable), else 0. */
int b200fe_host_wait_flag(b200fe_host_pool* pool, long long ticket, int* flag);
/* 1 when `windows` windows of `window_len` float64 samples, spread evenly over every utterance, hold PCM16 values only (the cheap
 * test before a batch is packed with src_dtype 3), 0 otherwise, negative on error. */
int b200fe_host_pcm16_probe(const void* const* h_src, const long long* nsamp, int batch, int windows, int window_len);

/* Host-only SpecAugment planner (no device work): the rectangles of the reference's `freq_mask` / `time_mask` and the
 * (center, warped) pair of `time_warp` for a whole batch, utterances in order, drawn exactly as the reference draws them
 * (lasr/utils/specaugment.py:20-24,61-64,90-95): CPython `random.randrange` (MT19937, `_randbelow_with_getrandbits`) and
 * numpy's legacy `numpy.random.randint` (MT19937, masked rejection), each on the caller's generator state -- `py_key` /
 * `np_key` are the 624 key words of `random.getstate()` / `numpy.random.get_state()`, `*py_pos` / `*np_pos` their positions;
 * all four are advanced in place, so putting them back leaves both global generators where the reference would leave them.
 * frames[u] = number of feature frames.  masks [batch][n_freq_mask + n_time_mask][2] (start, stop; (0,0) = skipped),
 * row_bounds [batch][2 n_time_mask] sorted (row classes of the statistics), warps [batch][2] or NULL ((-1,-1) = not warped;
 * drawn only when draw_time_warp != 0). */
int b200fe_specaug_plan(unsigned int* py_key, int* py_pos, unsigned int* np_key, int* np_pos, const long long* frames, int batch,
                        int num_mel, int max_freq_width, int n_freq_mask, int max_time_width, int n_time_mask,
                        int draw_time_warp, int max_time_warp, int* masks, int* row_bounds, int* warps);

/* Encoder source mask without a host round trip.  The reference builds it on the CPU from `xlen.tolist()`
 * (lasr/model/e2e_ctc_att/e2e_base.py:19-20 -> make_pad_mask, lasr/utils/mask.py:5-45, then `~mask`, unsqueeze(-2)):
 *   subsample == 1:  d_mask [batch][1][max_frames]          = (t < frames[b])          -- the mask the encoder receives
 *   subsample == 4:  d_mask [batch][1][((T-1)/2-1)/2]       = (4 i < frames[b])        -- what Conv2dSubsampling returns
 *                    (x_mask[:, :, :-2:2][:, :, :-2:2], lasr/modules/net/transformer/subsampling.py:60)
 * d_len holds frame counts (the batch dict's wav_len) or, with len_is_samples != 0, sample counts (frames derived with the
 * plan's window / shift, TA:63-67).  d_out_len (optional) = number of true cells per utterance (hs_len of
 * E2E_CTC_ATT.subfunction, e2e_base.py:47-49).  Masks are bytes (torch.bool layout). */
int b200fe_src_mask(const b200fe_plan* plan, const long long* d_len, int len_is_samples, int batch, int max_frames,
                    int subsample, unsigned char* d_mask, long long* d_out_len, void* stream);

/* Ragged row copy in ONE kernel launch: row u = d_nbytes[u] bytes from src + d_src_off[u] to dst + d_dst_off[u] (byte
 * offsets; the three arrays are device-readable, max_bytes >= every d_nbytes[u]).  Either side may be PINNED host memory
 * (under unified addressing the SMs read / write it over PCIe), which replaces the per-utterance cudaMemcpyAsync of
 * b200fe_h2d_ragged / b200fe_d2h_ragged -- the GPU-side counterpart of the reference's collate copy loop
 * (lasr/data/dataset.py:8-22, one numpy row assignment per utterance).  16-byte aligned rows move as 128-bit accesses. */
int b200fe_copy_ragged(const void* src, const long long* d_src_off, void* dst, const long long* d_dst_off,
                       const long long* d_nbytes, int batch, long long max_bytes, void* stream);

/* Waveform ingest on the device.  b200fe_resample_poly: rational-ratio polyphase resampling of a ragged
 * batch, utterance u = d_n_in[u] samples at d_in + d_in_off[u] -> d_n_out[u] samples at d_out + d_out_off[u]:
 *     out[m] = scale * sum_j h[(m + pre_remove) * down - j * up] * in[j]
 * (the index arithmetic of scipy.signal.resample_poly / upfirdn; h = d_filter[filter_len], built by the caller).  Stands in for
 * ReSample (librosa `kaiser_fast`, R/lasr/data/datatrans.py:16-20) and, with up/down = 10/11 or 10/9, for SoxSpeedPt (sox `speed`,
 * datatrans.py:29-39); both reference filters live in libraries absent here -- see DESIGN.md for what is and is not pinned.
 * b200fe_avg_channels: y[i] = mean over c of x[i][c] for interleaved (n, channels) input (AverageChanl, datatrans.py:10-14). */
int b200fe_resample_poly(const float* d_in, const long long* d_in_off, const long long* d_n_in, int batch, float* d_out,
                         const long long* d_out_off, const long long* d_n_out, long long max_n_out, const float* d_filter, int filter_len,
                         int up, int down, int pre_remove, float scale, void* stream);
int b200fe_avg_channels(const float* d_in, float* d_out, long long n, int channels, void* stream);

/* bfloat16 feature emission: the encoder's first layer (Conv2dSubsampling,
 * R/lasr/modules/net/transformer/subsampling.py:53-57) runs in bfloat16 under autocast.  b200fe_cast_bf16 converts n float32
 * values (round to nearest even, = tensor.to(torch.bfloat16)); b200fe_copy_ragged_bf16 is b200fe_copy_ragged with the
 * conversion folded in (d_nbytes = bytes of float32 per row, multiples of 16; destination rows hold half as many bytes), so
 * features that go to the host cross PCIe as bfloat16 with no extra pass. */
int b200fe_cast_bf16(const float* d_in, void* d_out, long long n, void* stream);
int b200fe_copy_ragged_bf16(const float* src, const long long* d_src_off, void* dst, const long long* d_dst_off,
                            const long long* d_nbytes, int batch, long long max_bytes, void* stream);

/* Turns per-utterance statistics into (a) utterance CMVN vectors and (b) the SpecAugment mean
 * fills of R/lasr/utils/specaugment.py:71-74,102-105 (each mask is filled with the mean of the
 * CURRENT array, i.e. after CMVN and after all earlier masks), evaluated in closed form from the
 * row-class column sums.  cmvn_mode: 0 = statistics are already in the output domain,
 * 1 = subtract utterance mean, 2 = mean and variance (var floored at 1e-20). */
typedef struct b200fe_post_args {
    unsigned int struct_size;    /* = sizeof(b200fe_post_args) */
    float* d_feats;              /* [batch][max_frames][num_mel_bins], updated in place */
    const long long* d_nsamp;    /* [batch] (frame counts are derived exactly as in the fused launch) */
    int batch;
    int max_frames;
    const double* d_stats;       /* as written by b200fe_fbank_fused (stats_stride != 0) */
    long long stats_stride;
    const int* d_row_bounds;
    int n_row_classes;
    int cmvn_mode;
    float* d_cmvn_mean;          /* [batch][num_mel_bins] workspace/outputs (required when cmvn_mode != 0) */
    float* d_cmvn_istd;
    const int* d_masks;          /* as above; NULL = no SpecAugment */
    int n_freq_masks, n_time_masks;
    float* d_fills;              /* [batch][n_freq_masks + n_time_masks] outputs (required with masks) */
    int fill_zero;               /* 1 = replace_with_zero: masked cells become 0 instead of the running mean */
    const long long* d_feat_offsets; /* optional [batch]: d_feats is packed as written with d_out_offsets */
} b200fe_post_args;

int b200fe_postpass(const b200fe_plan* plan, const b200fe_post_args* args, void* stream);

/* SpecAugment time warp (R/lasr/utils/specaugment.py:4-32, the first step of the registry transform
 * `specaug`, R/lasr/data/datatrans.py:136): per utterance, rows [0, center) are resized to `warped` rows and
 * rows [center, T) to T - warped rows with Pillow's BICUBIC arithmetic (float64 coefficients and
 * accumulation, one rounding to float32) -- bit-identical to the reference.  (center, warped) are drawn on
 * the host with the reference's generator; center < 0 copies the utterance (T - W <= W, no draw).
 * Out of place: d_out gets the warped features (padded rows zero) and, optionally, their statistics in the
 * layout of b200fe_fbank_fused for the mean fills of the masks that follow -- or the masked result itself
 * (d_masks). */
typedef struct b200fe_warp_args {
    unsigned int struct_size;    /* = sizeof(b200fe_warp_args) */
    const float* d_in;           /* [batch][max_frames][num_mel_bins] */
    float* d_out;                /* same shape, must not alias d_in */
    const long long* d_nsamp;    /* [batch] */
    int batch;
    int max_frames;
    const int* d_warp;           /* [batch][2] int32 (center, warped) */
    double* d_stats;             /* optional, zeroed by the caller */
    long long stats_stride;
    const int* d_row_bounds;
    int n_row_classes;
    /* Optional: the SpecAugment masks that follow the warp (specaugment.py:47-106), applied in the same launch by the CTA that
     * completes an utterance -- fills derived as b200fe_postpass derives them (running mean of the current array, or zero),
     * later masks win.  Needs d_stats (+ d_row_bounds for mean fills); in this mode d_stats is a WORKSPACE: zero on entry, zero
     * again on exit, so one buffer serves every call.  NULL = warp only (call b200fe_postpass for the masks). */
    const int* d_masks;          /* [batch][n_freq_masks + n_time_masks][2] int32 (lo, hi) */
    int n_freq_masks, n_time_masks;
    float* d_fills;              /* [batch][n_freq_masks + n_time_masks] outputs */
    int fill_zero;               /* 1 = replace_with_zero */
    int* d_utt_done;             /* [batch] int32 workspace: zero before the first call, left zero by every call */
} b200fe_warp_args;

int b200fe_time_warp(const b200fe_plan* plan, const b200fe_warp_args* args, void* stream);

/* Global CMVN: [2][num_mel_bins + 1] Kaldi statistics (row 0 sums + count, row 1 sums of
 * squares + 0) on the HOST -> mean / inverse std vectors (float32, host). */
int b200fe_cmvn_from_stats(const double* stats, int num_mel_bins, int norm_vars, float* mean, float* istd);

#ifdef __cplusplus
}
#endif
#endif /* B200FE_H */

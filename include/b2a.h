/* b2a.h — C ABI of the B200-native Stage-2 audio feature-extraction path.
 *
 * Drop-in boundary for ONE hot path of gcpgarcias/audio-edge-ml-pipeline: the arithmetic behind
 * the `audio_mel_spec`, `audio_mfcc_seq` and `audio_cqt` extractors
 * (reference: src/preprocessing/feature_extraction/audio/deep.py:75-134, 196-260, 268-328),
 * i.e. everything those classes delegate to librosa 0.11.0 between "decoded, padded/trimmed
 * mono clip" and "normalised float32 feature matrix".
 *
 * Nothing in the reference calls C on this path (it is pure Python over librosa); the reference
 * FFI a maintainer would add is the ctypes stub shown in INTEGRATION.md, binding exactly the
 * symbols below.  Plain C symbols, plain pointers and sizes, no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B2A_E* code on failure;
 *     b2a_last_error() returns a thread-local, NUL-terminated description of the last failure.
 *   - one handle = one (device, configuration); one host thread uses a handle at a time;
 *     different handles / devices are independent.
 *   - the caller owns all sample and feature buffers; the library owns its constant tables
 *     (window, twiddles, banded mel weights, DCT matrix, CQT bases, decimator taps), all built on
 *     the host in double precision inside b2a_create().
 *   - clips are fixed length within a call: `n_samples` per clip, already padded/trimmed by the
 *     caller exactly as deep.py:52-53,58-61 does (right zero-pad / truncate).
 *   - input layout:  [n_clips][n_samples]  int16 PCM (value/32768 is applied on device,
 *                    deep.py:44-50 via librosa.load; model_to_c.py:577) or float32.
 *   - output layout: [n_clips][rows][frames] float32, C-contiguous — the `(N, rows, T)` layout
 *                    FeaturePipeline.save writes to features.npy (pipeline.py:125-170).
 */
#ifndef B2A_H_
#define B2A_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2A_ABI_VERSION 1

/* error codes */
#define B2A_OK            0
#define B2A_EINVAL       -1   /* bad argument / unsupported configuration            */
#define B2A_ECUDA        -2   /* CUDA runtime error (message in b2a_last_error)      */
#define B2A_ENOMEM       -3   /* host or device allocation failed                    */
#define B2A_ENODEVICE    -4   /* no usable CUDA device: there is NO CPU fallback     */

/* b2a_config.kind — which extractor's arithmetic (deep.py class in parentheses) */
#define B2A_KIND_MEL      0   /* AudioMelSpectrogram.extract  deep.py:112-134 */
#define B2A_KIND_MFCC     1   /* AudioMFCCSequence.extract    deep.py:304-328 */
#define B2A_KIND_CQT      2   /* AudioCQT.extract             deep.py:235-260 */
#define B2A_KIND_CLASSICAL 3  /* AudioClassicalExtractor._compute_features  classical.py:272-355:
                               * rows = 6 n_mfcc + 62, frames = 1.  A clip's vector holds EVERY feature group with
                               * both aggregations in the reference's canonical order — mfcc, delta_mfcc, delta2_mfcc
                               * (n_mfcc means then n_mfcc stds each), spectral_centroid, spectral_rolloff,
                               * spectral_bandwidth (mean, std), spectral_contrast (7 + 7), spectral_flatness (2),
                               * chroma (12 + 12), zcr (2), rms (2), tonnetz (6 + 6); a caller configured with a
                               * subset of `features` / `aggregations` (classical.py:152-180) selects columns.
                               * n_fft in {512, 1024, 2048}; n_samples >= max(n_fft, 8 hop_length)
                               * (classical.py:262-270 pads to that); constant padding only. */

/* b2a_config.input_dtype */
#define B2A_IN_I16        0
#define B2A_IN_F32        1

/* b2a_config.pad_mode — librosa.stft(center=True, pad_mode=...) */
#define B2A_PAD_CONSTANT  0   /* librosa >= 0.10 default: zeros (CLAUDE.md:90-91)   */
#define B2A_PAD_REFLECT   1   /* selectable, never the default                      */

typedef struct b2a_config {
    int32_t kind;             /* B2A_KIND_*                                                   */
    int32_t input_dtype;      /* B2A_IN_*                                                     */
    int32_t sample_rate;      /* Hz                                   (all kinds)             */
    int32_t n_samples;        /* samples per clip, = int(duration*sr) (all kinds)             */
    int32_t n_fft;            /* 256|512|1024|2048                    (mel, mfcc, classical)  */
    int32_t hop_length;       /*                                      (all kinds)             */
    int32_t n_mels;           /* mel bands (mfcc: librosa default 128)(mel, mfcc, classical)  */
    int32_t n_mfcc;           /* DCT rows kept                        (mfcc, classical)       */
    int32_t n_bins;           /* CQT bins                             (cqt)                   */
    int32_t bins_per_octave;  /*                                      (cqt)                   */
    double  fmin;             /* CQT lowest frequency; <=0 -> C1 = 32.7032 Hz   (cqt)         */
    int32_t pad_mode;         /* B2A_PAD_*                            (mel, mfcc)             */
    float   top_db;           /* dB floor below the peak; reference uses librosa's 80.0       */
    int32_t reserved[8];      /* must be zero                                                 */
} b2a_config;

typedef struct b2a_handle b2a_handle;

/* Fills *cfg with the reference defaults of `kind` (deep.py:98-105, 219-227, 290-297 plus
 * librosa's n_mels=128 for mfcc) and n_samples = 0; the caller then overrides fields. */
int b2a_default_config(int32_t kind, b2a_config* cfg);

/* Builds tables in double precision on the host, uploads them to `device`, allocates scratch.
 * Fails with B2A_ENODEVICE when no CUDA device is present — there is no CPU path. */
int b2a_create(const b2a_config* cfg, int32_t device, b2a_handle** out);
int b2a_destroy(b2a_handle* h);

/* rows = n_mels | n_mfcc | n_bins; frames = 1 + n_samples / hop_length  (CLAUDE.md:90) */
int b2a_out_shape(const b2a_handle* h, int32_t* rows, int32_t* frames);

/* classical handles: the tuning (in fractions of a semitone, -0.5 .. 0.49) that chroma_stft's estimate_tuning
 * step found for clips [0, n) of the LAST device launch on this handle (b2a_run_device, or the last chunk of
 * b2a_run_host) — a diagnostic: the estimate is the arg-max of a 100-bin histogram and therefore the one
 * discontinuous step of the extractor.  Synchronises the device. */
int b2a_classical_tunings(b2a_handle* h, float* out, int64_t n);

/* Device-resident path: d_clips and d_out are device pointers on the handle's device; the work
 * is enqueued on `stream` (a cudaStream_t, NULL = legacy default stream) and the call returns
 * without synchronising.  d_clips: [n_clips][n_samples] of input_dtype; d_out:
 * [n_clips][rows][frames] float32.  mfcc and cqt handles work in per-handle device scratch: two
 * calls on the SAME handle must not run concurrently (use one stream, or order the streams with an
 * event); different handles are independent. */
int b2a_run_device(b2a_handle* h, const void* d_clips, int64_t n_clips, float* d_out,
                   void* stream);

/* Host path: clips/out are HOST pointers (pinned or pageable).  The library chunks the batch,
 * overlaps H2D / kernels / D2H on its own streams and returns when `out` is complete. */
int b2a_run_host(b2a_handle* h, const void* clips, int64_t n_clips, float* out);

/* Diagnostic twin of b2a_run_host: the same chunked H2D / D2H schedule on the same streams and device
 * buffers with the kernels left out (`out` receives unspecified data).  Its duration is the transfer
 * ceiling of the host path on this machine; bench.py reports the end-to-end rate against it. */
int b2a_run_host_copy_only(b2a_handle* h, const void* clips, int64_t n_clips, float* out);

/* Ragged batches (`duration=None` in the reference: every clip keeps its own length and frame
 * count; deep.py:122-124 is skipped).  mel, mfcc and classical handles (a classical clip yields its `rows`
 * floats whatever its length and must hold at least 8 hops — the reference extractor takes whole files).  Clip i has lengths[i] samples,
 * n_fft <= lengths[i] <= cfg.n_samples (the handle's n_samples is the MAXIMUM length), stored at
 * element offset in_offsets[i] of `clips` (a multiple of 8 elements keeps the TMA staging path);
 * its (rows, 1 + lengths[i]/hop) float32 features are written at float offset out_offsets[i] of
 * `out`.  Host variant: all five pointers are host pointers, total_in / total_out are the element
 * counts of `clips` / `out`.  Device variant: all five are device pointers, asynchronous on `stream`. */
int b2a_run_host_ragged(b2a_handle* h, const void* clips, int64_t total_in, const int64_t* in_offsets,
                        const int32_t* lengths, const int64_t* out_offsets, int64_t n_clips,
                        float* out, int64_t total_out);
int b2a_run_device_ragged(b2a_handle* h, const void* d_clips, const int64_t* d_in_offsets,
                          const int32_t* d_lengths, const int64_t* d_out_offsets, int64_t n_clips,
                          float* d_out, void* stream);

/* Host-side front end for the common case (no GPU involved): decode n_files mono 16-bit PCM
 * RIFF/WAVE files that are already at `sample_rate` straight into dst[n_files][n_samples]
 * (typically pinned memory), applying what deep.py:30-61 applies per clip: optional segment
 * [offset_s[i], offset_s[i] + duration_s[i]) in native frames (NULL arrays = whole file,
 * duration < 0 = to the end), truncate to n_samples, right zero-pad.  status[i] receives
 * B2A_DEC_*; rows whose status is not B2A_DEC_OK are zero-filled and are the caller's to handle
 * (other formats -> a fuller decoder; otherwise skip the sample as base.py:204-206 does).
 * n_threads <= 0 picks min(32, hardware threads). */
#define B2A_DEC_OK            0
#define B2A_DEC_EIO           1   /* open/read failed                                   */
#define B2A_DEC_EFORMAT       2   /* not a RIFF/WAVE file or malformed                  */
#define B2A_DEC_EUNSUPPORTED  3   /* valid WAV, but not mono 16-bit PCM                 */
#define B2A_DEC_ERATE         4   /* file rate != sample_rate (the reference resamples) */
int b2a_decode_wav_pcm16_batch(const char* const* paths, int64_t n_files, int32_t sample_rate,
                               const double* offset_s, const double* duration_s, int32_t n_samples,
                               int16_t* dst, int32_t* status, int32_t n_threads);

/* The general case of the same front end (what librosa.load -> soundfile covers for RIFF/WAVE,
 * deep.py:44-50; header probe of dataset_loaders/audio_folder_loader.py:76-103): PCM 8/16/24/32-bit and
 * IEEE float 32/64, any channel count (float32 channel mean = librosa.to_mono), any rate.
 * b2a_probe_wav_batch reads headers only; every output array is optional except status.
 * b2a_decode_wav_batch decodes frames [offset, offset + duration) of each file AT THE FILE'S RATE, at most
 * max_frames of them, into row i of dst (dst_stride elements apart; the rest of the row is zero-filled):
 * out_dtype B2A_IN_I16 copies mono PCM16 as is (anything else: B2A_DEC_EUNSUPPORTED), B2A_IN_F32 converts
 * every supported format the way libsndfile scales it (PCM16 / 32768, ...).  rate[i] / n_out[i] receive the
 * file's rate and the frames written; files at another rate than the extractor's go on to the resampler
 * (b2a_run_host_resampled).  Non-WAV containers report B2A_DEC_EFORMAT and are the caller's to skip. */
int b2a_probe_wav_batch(const char* const* paths, int64_t n_files, int32_t* rate, int32_t* channels,
                        int32_t* bits, int32_t* format_tag, int64_t* n_frames, int32_t* status, int32_t n_threads);
int b2a_decode_wav_batch(const char* const* paths, int64_t n_files, const double* offset_s,
                         const double* duration_s, int64_t max_frames, int32_t out_dtype, void* dst,
                         int64_t dst_stride, int32_t* rate, int32_t* n_out, int32_t* status, int32_t n_threads);

/* Rational resampler for files whose rate differs from `sample_rate` — what librosa.load does
 * with soxr_hq inside `_load_segment` (deep.py:44-50) before any extractor sees the samples.
 * Zero-phase Kaiser-sinc low-pass (pass band to 0.913 x, stop band from 1.0 x the lower Nyquist,
 * 125 dB: the CQT decimator's specification), zero-extended edges, output length
 * ceil(n_in * target / orig) as librosa.resample fixes it, float32 out (int16 input is scaled by
 * 1/32768 like librosa.load).  One signal per call; `in_dtype` is B2A_IN_I16 or B2A_IN_F32.
 * A resampler belongs to one (device, orig, target); b2a_resampler_run_host may be called from several
 * threads (calls on one handle are serialised inside the library), run_device is the caller's to order;
 * no CPU path.
 * b2a_resampler_geometry exposes up/down (target/orig in lowest terms), the prototype's half
 * length and the polyphase table [up][taps_per_phase] for verification. */
typedef struct b2a_resampler b2a_resampler;
int b2a_resampler_create(int32_t orig_sr, int32_t target_sr, int32_t device, b2a_resampler** out);
int b2a_resampler_destroy(b2a_resampler* r);
int64_t b2a_resampler_out_len(const b2a_resampler* r, int64_t n_in);
int b2a_resampler_geometry(const b2a_resampler* r, int32_t* up, int32_t* down, int32_t* half_len,
                           int32_t* taps_per_phase, float* poly);
/* The same design without a device (host only): query sizes with poly == NULL, then pass a buffer
 * of up * taps_per_phase floats. */
int b2a_resampler_design(int32_t orig_sr, int32_t target_sr, int32_t* up, int32_t* down, int32_t* half_len,
                         int32_t* taps_per_phase, float* poly, int64_t poly_capacity);
int b2a_resampler_run_host(b2a_resampler* r, const void* in, int32_t in_dtype, int64_t n_in, float* out);
int b2a_resampler_run_device(b2a_resampler* r, const void* d_in, int32_t in_dtype, int64_t n_in,
                             float* d_out, void* stream);
/* Many signals per launch: clip i = in[i * in_stride, + in_len[i]) (d_in_len is a DEVICE array); row i of
 * d_out (out_stride floats apart) receives outputs [0, out_cap): the resampled signal while it lasts
 * (ceil(in_len * target / orig) samples), zeros after it — librosa.load followed by the pad / trim to a
 * fixed duration (deep.py:52-61).  Asynchronous on `stream`. */
int b2a_resampler_run_device_batch(b2a_resampler* r, const void* d_in, int32_t in_dtype, int64_t n_clips,
                                   int64_t in_stride, const int32_t* d_in_len, float* d_out, int64_t out_stride,
                                   int64_t out_cap, void* stream);
int b2a_resampler_rates(const b2a_resampler* r, int32_t* orig_sr, int32_t* target_sr, int32_t* device);
/* Host batch of clips at the FILES' rate -> features: H2D, batched resampling to the handle's sample rate
 * with pad / trim to its n_samples, the extractor's kernels, D2H; chunked and overlapped like b2a_run_host.
 * `h` must be a float32-input handle on the resampler's device whose sample_rate is the resampler's target.
 * clips: [n_clips][in_stride] of in_dtype (host), in_len[i] <= in_stride valid samples each (host array). */
int b2a_run_host_resampled(b2a_handle* h, b2a_resampler* r, const void* clips, int32_t in_dtype, int64_t in_stride,
                           const int32_t* in_len, int64_t n_clips, float* out);
const char* b2a_resampler_last_error(void);   /* same text b2a_last_error() returns after a resampler call */

/* Stage-1b waveform augmentation on the device (src/preprocessing/augment.py:88-212, 325-375): every output
 * row is one source clip pushed through a chain of steps whose random parameters the HOST has already drawn
 * with the reference's generator, in the reference's order (numpy default_rng(seed), augment.py:325), so the
 * result equals the reference's y_aug bit for bit.  Ragged: row r reads lengths[r] samples at element offset
 * src_off[r] of `src` and writes lengths[r] samples at element offset out_off[r] of `out`
 * (_preserve_length, augment.py:206-212, is the identity for these length-preserving steps).
 *   B2A_AUG_GAIN      y * a                               volume_scale :88-93 and the level match :344-345
 *   B2A_AUG_NOISE     clip(y + noise[noise_off + i] * a)  gaussian_noise :96-102 (noise = float32 of the host's draws) and
 *                                                         pdm_hiss :135-167 (noise = the host-synthesised pink row)
 *   B2A_AUG_ROLL      np.roll(y, shift)                   time_shift :121-126
 *   B2A_AUG_POLARITY  -y                                  polarity_inversion :129-132
 * A row's chain is steps[r * max_steps ...] up to the first op < 0.  out_dtype B2A_IN_F32 gives y_aug;
 * B2A_IN_I16 quantises it the way soundfile writes PCM_16 (lrintf(x * 32768), saturated) — the samples
 * Stage 2 reads back after the reference's WAV round trip.  time_stretch / pitch_shift are not built. */
#define B2A_AUG_END      -1
#define B2A_AUG_GAIN      0
#define B2A_AUG_NOISE     1
#define B2A_AUG_ROLL      2
#define B2A_AUG_POLARITY  3
typedef struct b2a_aug_step {
    int32_t op;          /* B2A_AUG_*                                   */
    float   a;           /* gain / noise amplitude (float32 of the draw) */
    int32_t shift;       /* roll: int(uniform * len)                     */
    int32_t reserved;
    int64_t noise_off;   /* element offset of this step's noise row      */
} b2a_aug_step;
int b2a_augment_host(int32_t device, const void* src, int32_t src_dtype, int64_t src_elems, const int64_t* src_off,
                     const int32_t* lengths, const int64_t* out_off, int64_t n_out, const b2a_aug_step* steps,
                     int32_t max_steps, const float* noise, int64_t noise_elems, void* out, int32_t out_dtype,
                     int64_t out_elems);

/* The two librosa-backed augmentors (augment.py:105-118), batched over ragged float32 rows; host pointers.
 * time_stretch: librosa.effects.time_stretch(y, rate) = istft(phase_vocoder(stft(y, n_fft 2048, hop 512), rate),
 *   length = round(n / rate)).  The caller passes out_len[i] = int(round(lengths[i] / rates[i])) (Python's rounding:
 *   it is the reference's) and receives row i at out + out_off[i].
 * pitch_shift: librosa.effects.pitch_shift(y, sr, n_steps) = fix_length(resample(time_stretch(y, rate),
 *   orig_sr = sr / rate, target_sr = sr), len(y)) with rate = 2 ** (-n_steps / 12); the caller passes rates[i],
 *   ratios[i] = target_sr / orig_sr as librosa forms it and mid_len[i] = int(round(lengths[i] / rates[i]));
 *   row i of the output has lengths[i] samples.  The resampling step stands in for libsoxr's variable-rate
 *   path (parity unpinned). */
int b2a_time_stretch_host(int32_t device, const float* src, int64_t src_elems, const int64_t* src_off,
                          const int32_t* lengths, const double* rates, int64_t n_rows, float* out,
                          int64_t out_elems, const int64_t* out_off, const int32_t* out_len);
int b2a_pitch_shift_host(int32_t device, const float* src, int64_t src_elems, const int64_t* src_off,
                         const int32_t* lengths, const double* rates, const double* ratios,
                         const int32_t* mid_len, int64_t n_rows, float* out, int64_t out_elems,
                         const int64_t* out_off);
/* Device variant: every pointer is a device pointer on the current device; asynchronous on `stream`. */
int b2a_augment_device(const void* d_src, int32_t src_dtype, const int64_t* d_src_off, const int32_t* d_lengths,
                       const int64_t* d_out_off, int64_t n_out, int32_t max_len, const b2a_aug_step* d_steps,
                       int32_t max_steps, const float* d_noise, void* d_out, int32_t out_dtype, void* stream);

/* Number of CUDA kernel launches the last b2a_run_* call on this handle enqueued. */
int64_t b2a_last_launch_count(const b2a_handle* h);

/* Pinned host memory for run_host callers (so H2D/D2H are true async DMA). */
int b2a_alloc_pinned(size_t bytes, void** out);
int b2a_free_pinned(void* p);

/* Introspection of the constant tables, for tests and integrators (float32 copies).
 *   which: B2A_TABLE_*; on entry *count = capacity of `dst` in floats (dst may be NULL to query),
 *   on exit *count = number of floats the table holds. */
#define B2A_TABLE_WINDOW      0   /* [n_fft]                     periodic Hann               */
#define B2A_TABLE_MEL_DENSE   1   /* [n_mels][1+n_fft/2]         Slaney filterbank, dense    */
#define B2A_TABLE_DCT         2   /* [n_mfcc][n_mels]            orthonormal DCT-II          */
#define B2A_TABLE_DECIM_TAPS  3   /* [383]                       2:1 decimator (cqt)         */
#define B2A_TABLE_CQT_LENGTHS 4   /* [n_bins]                    wavelet lengths (cqt)       */
#define B2A_TABLE_CQT_BASIS   5   /* [n_octaves][n_filters][1+n_fft_o/2][2] re,im; dense     */
#define B2A_TABLE_CHROMA      6   /* [100][12][1+n_fft/2]        chroma banks, tuning -0.5 + 0.01 i (classical) */
#define B2A_TABLE_TONNETZ     7   /* [6][12]                     tonnetz projection (classical) */
#define B2A_TABLE_CONTRAST_BANDS 8 /* [7] first bin, [7] bins, [7] q, then piptrack k0, k1 (classical) */
int b2a_get_table(const b2a_handle* h, int32_t which, float* dst, int64_t* count);

/* CQT geometry: per-octave FFT size / hop / signal length (cqt handles only). */
int b2a_cqt_geometry(const b2a_handle* h, int32_t* n_octaves, int32_t* n_filters,
                     int32_t* n_fft /*[n_octaves]*/, int32_t* hop /*[n_octaves]*/,
                     int32_t* sig_len /*[n_octaves]*/);

const char* b2a_last_error(void);
int b2a_abi_version(void);
int b2a_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B2A_H_ */

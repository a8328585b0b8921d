// classical.h — launch interface of the audio_classical kernel (classical.cu; reference classical.py:272-355).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace b2a {

struct ClassicalParams {
    const void* clips;          // [n_clips][n_samples] int16 or float32 (device)
    float* out;                 // [n_clips][6 n_mfcc + 62] (device): every group, [mean..., std...], canonical order
    const float* window;        // [n_fft]
    const float2* tw;           // [NC]      exp(-2 pi i k / NC)
    const float2* tw2;          // [NC/2+1]  exp(-i pi k / NC)
    const int* mel_k0;          // banded mel filterbank
    const int* mel_cnt;
    const int* mel_off;
    const float* mel_w;
    const float* dct;           // [n_mfcc][n_mels]
    const float* chroma;        // [100][12][1 + n_fft/2]: librosa.filters.chroma for tuning = -0.5 + 0.01 i
    const float* tonnetz;       // [6][12]
    const int* band_start;      // spectral_contrast: first bin, bin count, q of the seven bands
    const int* band_cnt;
    const int* band_q;
    float* scratch;             // [grid][scratch_per_cta] floats
    size_t scratch_per_cta;
    float* tuning_out;          // [n_clips] estimated tuning of every clip (diagnostic; may be null)
    // ragged batches (NULL for fixed-length): per-clip element offset into clips, length in samples, float offset
    // into out; n_samples / n_frames then hold the MAXIMA (scratch sizing)
    const long long* rag_in_off;
    const int* rag_len;
    const long long* rag_out_off;
    long long n_clips;
    int n_samples, hop, n_frames, n_mels, mel_nnz, n_mfcc, sample_rate;
    int pip_k0, pip_k1;         // piptrack bins: 150 Hz <= f < min(4 kHz, sr / 2)
    int cand_cap;               // capacity of the per-clip candidate list
    float top_db;
};

size_t classical_smem_bytes(int log2nc, int hop, int n_mels, int mel_nnz);
size_t classical_scratch_floats(int n_fft, int n_frames, int n_mels, int n_mfcc, int cand_cap);
cudaError_t launch_classical(const ClassicalParams& p, int log2nc, bool i16, int grid, cudaStream_t st);

}  // namespace b2a

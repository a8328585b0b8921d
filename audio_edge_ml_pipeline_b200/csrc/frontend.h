// frontend.h — launch interface of the fused STFT/mel kernels (frontend.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace b2a {

struct FrontParams {
    const void* clips;       // [n_clips][n_samples] int16 or float32 (device)
    float* out;              // [n_clips][rows][n_frames] (device)
    float* inter;            // mfcc only: [grid][n_mels][n_frames] raw-dB scratch (device)
    const float* window;     // [n_fft]
    const float2* tw;        // [NC]       exp(-2 pi i k / NC)
    const float2* tw2;       // [NC/2+1]   exp(-i pi k / NC)
    const int* mel_k0;       // banded mel filterbank
    const int* mel_cnt;
    const int* mel_off;
    const float* mel_w;
    // 4-padded, 0.25-prescaled banded weights + balanced band order (logmel512 kernel)
    const float* mel_wq;
    const int* mel_k0e;      // per band: first bin rounded down to even
    const int* mel_cnt4;     // per band: number of 4-bin steps (from mel_k0e)
    const int* mel_off4;     // per band: float offset into mel_wq (multiple of 4)
    const int* mel_order;    // band processed at position i (position i belongs to warp i % 8)
    int mel_wpad;            // floats in mel_wq
    int mel_special;         // 1: the generated straight-line mel code matches this configuration
    const float* dct;        // mfcc only: [n_mfcc][n_mels]
    // ragged batches (NULL for fixed-length): per-clip element offset into clips, length in samples,
    // float offset into out; n_samples/n_frames then hold the MAXIMA (scratch sizing)
    const long long* rag_in_off;
    const int* rag_len;
    const long long* rag_out_off;
    long long n_clips;
    int n_samples, hop, n_frames, n_mels, mel_nnz, n_mfcc, pad_mode;
    float top_db;
};

size_t front_smem_bytes(int log2nc, int hop, int n_mels, int mel_nnz);
int front_ctas_per_sm(int log2nc);
// kind: 0 = mel, 1 = mfcc
cudaError_t launch_front(const FrontParams& p, int log2nc, bool i16, int kind, int grid, cudaStream_t st);

// specialised n_fft = 512 kernel (logmel512.cu); requires an even hop
size_t logmel512_smem_bytes(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc);   // n_mfcc = 0: mel
bool logmel512_has_special(int sample_rate, int n_mels);
int logmel512_ctas_per_sm();     // persistent warp-specialised CTAs per SM (1)
int logmel512_mel_warps();       // mel (consumer) warps per CTA: table-driven bands are dealt round-robin to them
cudaError_t launch_logmel512(const FrontParams& p, bool i16, int kind, int grid, cudaStream_t st);

// specialised n_fft = 1024 kernel (logmel1024.cu): one warp per frame (two interleaved 512-sample real FFTs, then
// the frame's mel bands and DCT in parallel lanes); needs hop % 4 == 0; mel_order deals the bands to lanes by width
size_t logmel1024_smem_bytes(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc);  // n_mfcc = 0: mel
bool logmel1024_supports(int hop, int n_mels, int n_mfcc);
int logmel1024_pow_rows();       // bin-pair rows of the power tile (513 bins + padding): reads must stay below it
cudaError_t launch_logmel1024(const FrontParams& p, bool i16, int kind, int grid, cudaStream_t st);

}  // namespace b2a

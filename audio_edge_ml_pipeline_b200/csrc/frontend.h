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
    const float* dct;        // mfcc only: [n_mfcc][n_mels]
    long long n_clips;
    int n_samples, hop, n_frames, n_mels, mel_nnz, n_mfcc, pad_mode;
    float top_db;
};

size_t front_smem_bytes(int log2nc, int hop, int n_mels, int mel_nnz);
int front_ctas_per_sm(int log2nc);
// kind: 0 = mel, 1 = mfcc
cudaError_t launch_front(const FrontParams& p, int log2nc, bool i16, int kind, int grid, cudaStream_t st);

}  // namespace b2a

// cqt.h — device side of audio_cqt (reference: deep.py:235-260 -> librosa.cqt/vqt).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b2a.h"
#include "tables.h"

namespace b2a {

struct CqtOctaveDev {
    float2* coef = nullptr;    // time-domain wavelets [row block of 12][n_fft][12] (re, im)
    int n_blocks = 0;          // row blocks of 12
    size_t sig_off = 0;        // float offset of this octave's signal inside a clip's scratch
};

struct CqtDevice {

    int64_t chunk_clips = 0;
    size_t scratch_per_clip = 0;       // floats
    float* scratch = nullptr;          // [chunk_clips][scratch_per_clip]
    float* taps = nullptr;             // [383] even/odd interleaved as designed (natural order)
    float* inv_sqrt_len = nullptr;     // [n_bins]
    unsigned int* clip_max = nullptr;  // [chunk_clips] float bits of max |V|
    unsigned int* clip_min = nullptr;  // [chunk_clips] float bits of min |V|
    std::vector<CqtOctaveDev> oct;
    size_t early_off = 0;              // scratch offsets of the early-downsample chain outputs
    std::vector<size_t> early_offs;
    std::vector<int> early_lens;
    int sm_count = 0;
};

int cqt_device_init(const CqtPlan& plan, const b2a_config& cfg, int sm_count, size_t smem_optin,
                    CqtDevice* dev, std::string* err);
void cqt_device_free(CqtDevice* dev);
int cqt_run(const CqtPlan& plan, const b2a_config& cfg, CqtDevice* dev, const void* d_clips,
            int64_t n_clips, float* d_out, cudaStream_t st, int64_t* launches, std::string* err);

}  // namespace b2a

// api.cu — implementation of the C ABI declared in include/b2a.h.
//
// Reference interface each entry point stands in for (all in /root/reference):
//   b2a_create / b2a_default_config  <- extractor constructors
//        src/preprocessing/feature_extraction/audio/deep.py:98-110, 219-233, 290-302
//   b2a_run_host / b2a_run_device    <- the librosa calls inside the three extract() bodies
//        deep.py:126-134 (mel), 249-260 (cqt), 318-328 (mfcc)
//   b2a_out_shape                    <- "n_frames = 1 + n_samples // hop"  CLAUDE.md:90
#include "../../include/b2a.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "classical.h"
#include "cqt.h"
#include "frontend.h"
#include "tables.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,        \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

template <typename T>
cudaError_t upload(const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

int ilog2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

}  // namespace

// other translation units of the library (resample.cu) report through the same thread-local text
void b2a_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

struct b2a_handle {
    b2a_config cfg{};
    int device = 0, sm_count = 0;
    int rows = 0, frames = 0, log2nc = 0;
    size_t in_elem = 2;
    int64_t last_launches = 0;
    // host copies (introspection)
    std::vector<float> window, mel_dense, dct;
    b2a::BandedMel mel;
    // device tables (mel / mfcc)
    float* d_window = nullptr;
    float2* d_tw = nullptr;
    float2* d_tw2 = nullptr;
    int* d_k0 = nullptr;
    int* d_cnt = nullptr;
    int* d_off = nullptr;
    float* d_w = nullptr;
    float* d_wq = nullptr;      // 4-padded, 0.25-prescaled weights (logmel512)
    int* d_k0e = nullptr;
    int* d_cnt4 = nullptr;
    int* d_off4 = nullptr;
    int* d_order = nullptr;
    int mel_wpad = 0;
    bool use512 = false, use1024 = false;
    float* d_dct = nullptr;
    float* d_inter = nullptr;
    int grid_cap = 0;
    // classical
    b2a::ClassicalTables cls;
    float* d_chroma = nullptr;
    float* d_tonnetz = nullptr;
    int* d_bands = nullptr;          // [3][7]: start, count, q
    float* d_cls_scratch = nullptr;
    float* d_cls_tuning = nullptr;   // [cls_tuning_cap] tuning estimate of every clip of the last run
    int64_t cls_tuning_cap = 0, cls_last_n = 0;
    size_t cls_scratch_per_cta = 0;
    int cls_frames = 0, cls_cand_cap = 0;
    // cqt
    b2a::CqtPlan cqt;
    b2a::CqtDevice cqtdev;
    // run_host resources
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t ev_kernels[2] = {nullptr, nullptr};   // chunk c's kernels done (mfcc / cqt share one scratch per handle)
    void* d_in[2] = {nullptr, nullptr};
    float* d_out[2] = {nullptr, nullptr};
    int64_t chunk_clips = 0;
    bool host_ready = false;                          // set only after every run_host resource exists
    // run_host_ragged resources (grow-only, reused across calls)
    cudaStream_t rag_stream = nullptr;
    void* rag_in = nullptr; float* rag_out = nullptr; void* rag_meta = nullptr;
    size_t rag_in_cap = 0, rag_out_cap = 0, rag_meta_cap = 0;
    // run_host_resampled resources: raw clips at the files' rate + their lengths, one set per stream
    void* rs_raw[2] = {nullptr, nullptr};
    int* rs_len[2] = {nullptr, nullptr};
    size_t rs_raw_cap[2] = {0, 0};
};

extern "C" {

const char* b2a_last_error(void) { return g_err.c_str(); }
int b2a_abi_version(void) { return B2A_ABI_VERSION; }

int b2a_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b2a_default_config(int32_t kind, b2a_config* c) {
    if (!c) return fail(B2A_EINVAL, "cfg is NULL");
    std::memset(c, 0, sizeof(*c));
    c->kind = kind;
    c->input_dtype = B2A_IN_I16;
    c->pad_mode = B2A_PAD_CONSTANT;
    c->top_db = 80.0f;
    switch (kind) {
        case B2A_KIND_MEL:   // deep.py:98-105
            c->sample_rate = 16000; c->n_mels = 40; c->n_fft = 512; c->hop_length = 160; break;
        case B2A_KIND_MFCC:  // deep.py:290-297 (+ librosa.feature.mfcc n_mels default)
            c->sample_rate = 22050; c->n_mfcc = 40; c->n_fft = 1024; c->hop_length = 512; c->n_mels = 128; break;
        case B2A_KIND_CQT:   // deep.py:219-227
            c->sample_rate = 22050; c->hop_length = 512; c->n_bins = 84; c->bins_per_octave = 12; c->fmin = 0.0; break;
        case B2A_KIND_CLASSICAL:   // classical.py:139-150
            c->sample_rate = 22050; c->n_mfcc = 40; c->n_mels = 128; c->n_fft = 1024; c->hop_length = 512; break;
        default: return fail(B2A_EINVAL, "unknown kind");
    }
    return B2A_OK;
}

int b2a_destroy(b2a_handle* h) {
    if (!h) return B2A_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < 2; ++i) {
        if (h->streams[i]) { cudaStreamSynchronize(h->streams[i]); cudaStreamDestroy(h->streams[i]); }
        if (h->ev_kernels[i]) cudaEventDestroy(h->ev_kernels[i]);
        cudaFree(h->d_in[i]); cudaFree(h->d_out[i]);
    }
    if (h->rag_stream) { cudaStreamSynchronize(h->rag_stream); cudaStreamDestroy(h->rag_stream); }
    cudaFree(h->rag_in); cudaFree(h->rag_out); cudaFree(h->rag_meta);
    for (int i = 0; i < 2; ++i) { cudaFree(h->rs_raw[i]); cudaFree(h->rs_len[i]); }
    cudaFree(h->d_window); cudaFree(h->d_tw); cudaFree(h->d_tw2);
    cudaFree(h->d_k0); cudaFree(h->d_cnt); cudaFree(h->d_off); cudaFree(h->d_w);
    cudaFree(h->d_wq); cudaFree(h->d_k0e); cudaFree(h->d_cnt4); cudaFree(h->d_off4); cudaFree(h->d_order);
    cudaFree(h->d_dct); cudaFree(h->d_inter);
    cudaFree(h->d_chroma); cudaFree(h->d_tonnetz); cudaFree(h->d_bands); cudaFree(h->d_cls_scratch); cudaFree(h->d_cls_tuning);
    b2a::cqt_device_free(&h->cqtdev);
    delete h;
    return B2A_OK;
}

int b2a_create(const b2a_config* cfg, int32_t device, b2a_handle** out) {
    if (!cfg || !out) return fail(B2A_EINVAL, "cfg/out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(B2A_ENODEVICE, "no CUDA device visible: this library has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(B2A_EINVAL, "device index out of range");
    for (int r : cfg->reserved) if (r != 0) return fail(B2A_EINVAL, "reserved fields must be zero");
    if (cfg->input_dtype != B2A_IN_I16 && cfg->input_dtype != B2A_IN_F32) return fail(B2A_EINVAL, "input_dtype");
    if (cfg->sample_rate <= 0 || cfg->hop_length <= 0 || cfg->n_samples <= 0)
        return fail(B2A_EINVAL, "sample_rate, hop_length and n_samples must be positive");
    if (!(cfg->top_db > 0.f)) return fail(B2A_EINVAL, "top_db must be positive");

    b2a_handle* h = new (std::nothrow) b2a_handle();
    if (!h) return fail(B2A_ENOMEM, "host allocation failed");
    h->cfg = *cfg;
    h->device = device;
    h->in_elem = cfg->input_dtype == B2A_IN_I16 ? 2 : 4;
    auto bail = [&](int code, const std::string& m) { b2a_destroy(h); return fail(code, m); };
#define CU_TRY_H(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return bail(e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,              \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
    } while (0)

    CU_TRY_H(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY_H(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return bail(B2A_ENODEVICE, "device is not sm_100-class (kernels are built for sm_100a only)");
    h->sm_count = prop.multiProcessorCount;
    h->frames = 1 + cfg->n_samples / cfg->hop_length;

    if (cfg->kind == B2A_KIND_MEL || cfg->kind == B2A_KIND_MFCC) {
        const int n_fft = cfg->n_fft;
        if (n_fft != 256 && n_fft != 512 && n_fft != 1024 && n_fft != 2048)
            return bail(B2A_EINVAL, "n_fft must be one of 256, 512, 1024, 2048");
        if (cfg->n_samples < n_fft)      // deep.py:119 min_samples=n_fft
            return bail(B2A_EINVAL, "n_samples must be >= n_fft (deep.py pads to n_fft)");
        if (cfg->n_mels <= 0 || cfg->n_mels > 512) return bail(B2A_EINVAL, "n_mels out of range");
        if (cfg->pad_mode != B2A_PAD_CONSTANT && cfg->pad_mode != B2A_PAD_REFLECT) return bail(B2A_EINVAL, "pad_mode");
        const bool mfcc = cfg->kind == B2A_KIND_MFCC;
        if (mfcc && (cfg->n_mfcc <= 0 || cfg->n_mfcc > cfg->n_mels)) return bail(B2A_EINVAL, "n_mfcc must be in [1, n_mels]");
        h->log2nc = ilog2(n_fft / 2);
        h->rows = mfcc ? cfg->n_mfcc : cfg->n_mels;
        const int NC = n_fft / 2, n_bins = NC + 1;
        h->window = b2a::hann_periodic(n_fft);
        h->mel_dense = b2a::mel_filterbank(cfg->sample_rate, n_fft, cfg->n_mels);
        h->mel = b2a::band_mel(h->mel_dense, cfg->n_mels, n_bins);
        if (mfcc) {
            h->dct = b2a::dct2_ortho(cfg->n_mfcc, cfg->n_mels);
            const int F = h->log2nc <= 8 ? 32 : 16;
            if ((size_t)cfg->n_mels * 32 > (size_t)F * (NC + 1))
                return bail(B2A_EINVAL, "n_mels too large for this n_fft in the mfcc kernel");
        }
        const size_t smem = b2a::front_smem_bytes(h->log2nc, cfg->hop_length, cfg->n_mels, (int)h->mel.w.size());
        if (smem > (size_t)prop.sharedMemPerBlockOptin)
            return bail(B2A_EINVAL, "hop_length/n_fft/n_mels combination exceeds shared memory");
        std::vector<float> tw = b2a::twiddles(NC, NC);
        std::vector<float> tw2 = b2a::twiddles(2 * NC, NC / 2 + 1);
        CU_TRY_H(upload(h->window, &h->d_window));
        CU_TRY_H(upload(tw, (float**)&h->d_tw));
        CU_TRY_H(upload(tw2, (float**)&h->d_tw2));
        CU_TRY_H(upload(h->mel.k0, &h->d_k0));
        CU_TRY_H(upload(h->mel.cnt, &h->d_cnt));
        CU_TRY_H(upload(h->mel.off, &h->d_off));
        if (h->mel.w.empty()) h->mel.w.push_back(0.f);
        CU_TRY_H(upload(h->mel.w, &h->d_w));
        h->grid_cap = h->sm_count * b2a::front_ctas_per_sm(h->log2nc);
        const bool try512 = n_fft == 512 && (cfg->hop_length % 2) == 0;
        const bool try1024 = n_fft == 1024 && b2a::logmel1024_supports(cfg->hop_length, cfg->n_mels, mfcc ? cfg->n_mfcc : 0);
        if (try512 || try1024) {
            // tables of the specialised kernel: bands padded to float4 groups, |X|^2 -> 4|X|^2 folded
            // into the weights (x0.25 is exact), bands dealt to the mel warps in snake order by size
            std::vector<float> wq;
            std::vector<int> k0e(cfg->n_mels), cnt4(cfg->n_mels), off4(cfg->n_mels), order;
            for (int m = 0; m < cfg->n_mels; ++m) {
                off4[m] = (int)wq.size();
                k0e[m] = h->mel.k0[m] & ~1;                       // bands start on an even bin
                const int lead = h->mel.k0[m] - k0e[m];
                cnt4[m] = h->mel.cnt[m] ? (lead + h->mel.cnt[m] + 3) / 4 : 0;     // 4-bin steps
                for (int q = 0; q < cnt4[m] * 4; ++q) {
                    const int src = q - lead;
                    wq.push_back(src >= 0 && src < h->mel.cnt[m] ? 0.25f * h->mel.w[h->mel.off[m] + src] : 0.f);
                }
            }
            if (wq.empty()) wq.assign(4, 0.f);
            std::vector<int> idx(cfg->n_mels);
            for (int m = 0; m < cfg->n_mels; ++m) idx[m] = m;
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cnt4[a] > cnt4[b]; });
            order.assign(cfg->n_mels, 0);
            const int MW = b2a::logmel512_mel_warps();
            for (int i = 0; i < cfg->n_mels; ++i) {           // position i is served by mel warp i % MW
                const int rnd = i / MW, w = i % MW;
                const int src = rnd * MW + ((rnd & 1) ? MW - 1 - w : w);
                order[i] = idx[std::min(src, cfg->n_mels - 1)];
            }
            {   // the snake can alias at the ragged end; fall back to identity if not a permutation
                std::vector<int> seen(cfg->n_mels, 0);
                bool perm = true;
                for (int v : order) { if (seen[v]++) perm = false; }
                if (!perm) order = idx;
            }
            if (try1024) {
                // 1024 kernel: positions in descending width; positions 2i and 2i + 1 share a warp's two halves and
                // run the same number of 4-bin steps, so the narrower band is padded with zero weights to its
                // neighbour's length (leading zeros where trailing ones would read past the power tile)
                order = idx;
                std::vector<float> wq2;
                const int prow_cap = b2a::logmel1024_pow_rows();
                for (size_t i = 0; i < idx.size(); i += 2) {
                    const int steps = cnt4[idx[i]];
                    for (size_t e = i; e < std::min(i + 2, idx.size()); ++e) {
                        const int m = idx[e];
                        int start = k0e[m];
                        if (start / 2 + 2 * steps > prow_cap) start = 2 * (prow_cap - 2 * steps);
                        const int lead = h->mel.k0[m] - start;
                        off4[m] = (int)wq2.size();
                        for (int q = 0; q < 4 * steps; ++q) {
                            const int src = q - lead;
                            wq2.push_back(src >= 0 && src < h->mel.cnt[m] ? 0.25f * h->mel.w[h->mel.off[m] + src] : 0.f);
                        }
                        k0e[m] = start;
                        cnt4[m] = steps;
                    }
                }
                if (wq2.empty()) wq2.assign(4, 0.f);
                wq.swap(wq2);
            }
            h->mel_wpad = (int)wq.size();
            const size_t smem512 = try512
                ? b2a::logmel512_smem_bytes(cfg->hop_length, cfg->n_mels, h->mel_wpad, cfg->input_dtype == B2A_IN_I16, mfcc ? cfg->n_mfcc : 0)
                : b2a::logmel1024_smem_bytes(cfg->hop_length, cfg->n_mels, h->mel_wpad, cfg->input_dtype == B2A_IN_I16, mfcc ? cfg->n_mfcc : 0);
            const bool fits = smem512 <= (size_t)prop.sharedMemPerBlockOptin;
            if (fits) {
                CU_TRY_H(upload(wq, &h->d_wq));
                CU_TRY_H(upload(k0e, &h->d_k0e));
                CU_TRY_H(upload(cnt4, &h->d_cnt4));
                CU_TRY_H(upload(off4, &h->d_off4));
                CU_TRY_H(upload(order, &h->d_order));
                h->use512 = try512;
                h->use1024 = try1024;
                h->grid_cap = h->sm_count * b2a::logmel512_ctas_per_sm();
            }
        }
        if (mfcc) {
            CU_TRY_H(upload(h->dct, &h->d_dct));
            CU_TRY_H(cudaMalloc((void**)&h->d_inter, (size_t)h->grid_cap * cfg->n_mels * h->frames * sizeof(float)));
        }
    } else if (cfg->kind == B2A_KIND_CQT) {
        if (cfg->n_samples < 2 * cfg->hop_length)   // deep.py:242-243 min_samples = 2*hop
            return bail(B2A_EINVAL, "n_samples must be >= 2*hop_length (deep.py pads to 2*hop)");
        const char* perr = nullptr;
        if (!b2a::build_cqt_plan(cfg->sample_rate, cfg->hop_length, cfg->n_bins, cfg->bins_per_octave,
                                 cfg->fmin, cfg->n_samples, &h->cqt, &perr))
            return bail(B2A_EINVAL, perr ? perr : "cqt plan failed");
        h->rows = cfg->n_bins;
        h->frames = h->cqt.n_frames;
        std::string cerr;
        const int rc = b2a::cqt_device_init(h->cqt, *cfg, h->sm_count, (size_t)prop.sharedMemPerBlockOptin,
                                            &h->cqtdev, &cerr);
        if (rc != 0) return bail(rc, cerr);
    } else if (cfg->kind == B2A_KIND_CLASSICAL) {
        const int n_fft = cfg->n_fft;
        if (n_fft != 512 && n_fft != 1024 && n_fft != 2048) return bail(B2A_EINVAL, "classical: n_fft must be one of 512, 1024, 2048");
        // classical.py:262-270 pads every segment to max(min_duration * sr, n_fft, 8 * hop) samples (delta width 9)
        if (cfg->n_samples < std::max(n_fft, 8 * cfg->hop_length))
            return bail(B2A_EINVAL, "classical: n_samples must be >= max(n_fft, 8 * hop_length) (classical.py pads to it)");
        if (cfg->n_mels <= 0 || cfg->n_mels > 512) return bail(B2A_EINVAL, "n_mels out of range");
        if (cfg->n_mfcc <= 0 || cfg->n_mfcc > cfg->n_mels) return bail(B2A_EINVAL, "n_mfcc must be in [1, n_mels]");
        if (cfg->pad_mode != B2A_PAD_CONSTANT) return bail(B2A_EINVAL, "classical: pad_mode must be constant (librosa default)");
        h->log2nc = ilog2(n_fft / 2);
        h->rows = 6 * cfg->n_mfcc + 62;
        h->cls_frames = h->frames;
        h->frames = 1;
        const int NC = n_fft / 2, n_bins = NC + 1;
        h->window = b2a::hann_periodic(n_fft);
        h->mel_dense = b2a::mel_filterbank(cfg->sample_rate, n_fft, cfg->n_mels);
        h->mel = b2a::band_mel(h->mel_dense, cfg->n_mels, n_bins);
        h->dct = b2a::dct2_ortho(cfg->n_mfcc, cfg->n_mels);
        const char* perr = nullptr;
        if (!b2a::build_classical_tables(cfg->sample_rate, n_fft, &h->cls, &perr))
            return bail(B2A_EINVAL, perr ? perr : "classical tables failed");
        if (h->mel.w.empty()) h->mel.w.push_back(0.f);
        const size_t smem = b2a::classical_smem_bytes(h->log2nc, cfg->hop_length, cfg->n_mels, (int)h->mel.w.size());
        if (smem > (size_t)prop.sharedMemPerBlockOptin)
            return bail(B2A_EINVAL, "classical: hop_length/n_fft/n_mels combination exceeds shared memory");
        std::vector<float> tw = b2a::twiddles(NC, NC);
        std::vector<float> tw2 = b2a::twiddles(2 * NC, NC / 2 + 1);
        CU_TRY_H(upload(h->window, &h->d_window));
        CU_TRY_H(upload(tw, (float**)&h->d_tw));
        CU_TRY_H(upload(tw2, (float**)&h->d_tw2));
        CU_TRY_H(upload(h->mel.k0, &h->d_k0));
        CU_TRY_H(upload(h->mel.cnt, &h->d_cnt));
        CU_TRY_H(upload(h->mel.off, &h->d_off));
        CU_TRY_H(upload(h->mel.w, &h->d_w));
        CU_TRY_H(upload(h->dct, &h->d_dct));
        CU_TRY_H(upload(h->cls.chroma, &h->d_chroma));
        CU_TRY_H(upload(h->cls.tonnetz, &h->d_tonnetz));
        std::vector<int> bands;
        bands.insert(bands.end(), h->cls.band_start.begin(), h->cls.band_start.end());
        bands.insert(bands.end(), h->cls.band_cnt.begin(), h->cls.band_cnt.end());
        bands.insert(bands.end(), h->cls.band_q.begin(), h->cls.band_q.end());
        CU_TRY_H(upload(bands, &h->d_bands));
        // persistent CTAs (one clip at a time each); local maxima cannot be adjacent: at most every other piptrack bin
        h->grid_cap = h->sm_count;     // ~240 registers x 256 threads: one resident CTA per SM
        h->cls_cand_cap = h->cls_frames * ((h->cls.pip_k1 - h->cls.pip_k0 + 1) / 2 + 1);
        h->cls_scratch_per_cta = b2a::classical_scratch_floats(n_fft, h->cls_frames, cfg->n_mels, cfg->n_mfcc, h->cls_cand_cap);
        CU_TRY_H(cudaMalloc((void**)&h->d_cls_scratch, h->cls_scratch_per_cta * sizeof(float) * h->grid_cap));
    } else {
        return bail(B2A_EINVAL, "unknown kind");
    }
    *out = h;
    return B2A_OK;
#undef CU_TRY_H
}

int b2a_out_shape(const b2a_handle* h, int32_t* rows, int32_t* frames) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (rows) *rows = h->rows;
    if (frames) *frames = h->frames;
    return B2A_OK;
}

int64_t b2a_last_launch_count(const b2a_handle* h) { return h ? h->last_launches : 0; }

int b2a_classical_tunings(b2a_handle* h, float* out, int64_t n) {
    if (!h || !out) return fail(B2A_EINVAL, "handle/out is NULL");
    if (h->cfg.kind != B2A_KIND_CLASSICAL) return fail(B2A_EINVAL, "not a classical handle");
    if (n < 0 || n > h->cls_last_n) return fail(B2A_EINVAL, "n exceeds the clips of the last launch");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaDeviceSynchronize());
    if (n > 0) CU_TRY(cudaMemcpy(out, h->d_cls_tuning, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return B2A_OK;
}

static int run_device_impl(b2a_handle* h, const void* d_clips, int64_t n_clips, float* d_out,
                           cudaStream_t st, int64_t* launches, const long long* rag_in = nullptr,
                           const int* rag_len = nullptr, const long long* rag_out = nullptr) {
    if (n_clips == 0) return B2A_OK;
    if (h->cfg.kind == B2A_KIND_CQT) {
        std::string cerr;
        const int rc = b2a::cqt_run(h->cqt, h->cfg, &h->cqtdev, d_clips, n_clips, d_out, st, launches, &cerr);
        if (rc != 0) return fail(rc, cerr);
        return B2A_OK;
    }
    if (h->cfg.kind == B2A_KIND_CLASSICAL) {
        if (n_clips > h->cls_tuning_cap) {           // (grow-only; the previous run on this handle is ordered before by the caller)
            cudaFree(h->d_cls_tuning);
            h->d_cls_tuning = nullptr; h->cls_tuning_cap = 0;
            CU_TRY(cudaMalloc((void**)&h->d_cls_tuning, (size_t)n_clips * sizeof(float)));
            h->cls_tuning_cap = n_clips;
        }
        b2a::ClassicalParams c{};
        c.clips = d_clips; c.out = d_out;
        c.window = h->d_window; c.tw = h->d_tw; c.tw2 = h->d_tw2;
        c.mel_k0 = h->d_k0; c.mel_cnt = h->d_cnt; c.mel_off = h->d_off; c.mel_w = h->d_w;
        c.dct = h->d_dct; c.chroma = h->d_chroma; c.tonnetz = h->d_tonnetz;
        c.band_start = h->d_bands; c.band_cnt = h->d_bands + 7; c.band_q = h->d_bands + 14;
        c.scratch = h->d_cls_scratch; c.scratch_per_cta = h->cls_scratch_per_cta;
        c.tuning_out = h->d_cls_tuning;
        c.rag_in_off = rag_in; c.rag_len = rag_len; c.rag_out_off = rag_out;
        c.n_clips = n_clips; c.n_samples = h->cfg.n_samples; c.hop = h->cfg.hop_length; c.n_frames = h->cls_frames;
        c.n_mels = h->cfg.n_mels; c.mel_nnz = (int)h->mel.w.size(); c.n_mfcc = h->cfg.n_mfcc;
        c.sample_rate = h->cfg.sample_rate; c.pip_k0 = h->cls.pip_k0; c.pip_k1 = h->cls.pip_k1;
        c.cand_cap = h->cls_cand_cap; c.top_db = h->cfg.top_db;
        const int grid = (int)std::min<int64_t>(n_clips, h->grid_cap);
        CU_TRY(b2a::launch_classical(c, h->log2nc, h->cfg.input_dtype == B2A_IN_I16, grid, st));
        h->cls_last_n = n_clips;
        *launches += 1;
        return B2A_OK;
    }
    b2a::FrontParams p{};
    p.clips = d_clips; p.out = d_out; p.inter = h->d_inter;
    p.window = h->d_window; p.tw = h->d_tw; p.tw2 = h->d_tw2;
    p.mel_k0 = h->d_k0; p.mel_cnt = h->d_cnt; p.mel_off = h->d_off; p.mel_w = h->d_w;
    p.dct = h->d_dct;
    p.rag_in_off = rag_in; p.rag_len = rag_len; p.rag_out_off = rag_out;
    p.n_clips = n_clips; p.n_samples = h->cfg.n_samples; p.hop = h->cfg.hop_length;
    p.n_frames = h->frames; p.n_mels = h->cfg.n_mels; p.mel_nnz = (int)h->mel.w.size();
    p.n_mfcc = h->cfg.n_mfcc; p.pad_mode = h->cfg.pad_mode; p.top_db = h->cfg.top_db;
    p.mel_wq = h->d_wq; p.mel_k0e = h->d_k0e; p.mel_cnt4 = h->d_cnt4; p.mel_off4 = h->d_off4; p.mel_order = h->d_order;
    p.mel_wpad = h->mel_wpad;
    p.mel_special = (h->use512 && b2a::logmel512_has_special(h->cfg.sample_rate, h->cfg.n_mels)) ? 1 : 0;
    const int grid = (int)std::min<int64_t>(n_clips, h->grid_cap);
    const bool i16 = h->cfg.input_dtype == B2A_IN_I16;
    const int kind = h->cfg.kind == B2A_KIND_MFCC ? 1 : 0;
    if (h->use512) CU_TRY(b2a::launch_logmel512(p, i16, kind, grid, st));
    else if (h->use1024) CU_TRY(b2a::launch_logmel1024(p, i16, kind, grid, st));
    else CU_TRY(b2a::launch_front(p, h->log2nc, i16, kind, grid, st));
    *launches += 1;
    return B2A_OK;
}

int b2a_run_device(b2a_handle* h, const void* d_clips, int64_t n_clips, float* d_out, void* stream) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (n_clips < 0) return fail(B2A_EINVAL, "n_clips < 0");
    if (n_clips > 0 && (!d_clips || !d_out)) return fail(B2A_EINVAL, "NULL buffer");
    CU_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    return run_device_impl(h, d_clips, n_clips, d_out, (cudaStream_t)stream, &h->last_launches);
}

// Frees whatever a failed lazy init of the run_host resources left behind, so the next call retries.
static void host_teardown(b2a_handle* h) {
    for (int i = 0; i < 2; ++i) {
        if (h->streams[i]) { cudaStreamSynchronize(h->streams[i]); cudaStreamDestroy(h->streams[i]); h->streams[i] = nullptr; }
        if (h->ev_kernels[i]) { cudaEventDestroy(h->ev_kernels[i]); h->ev_kernels[i] = nullptr; }
        cudaFree(h->d_in[i]); h->d_in[i] = nullptr;
        cudaFree(h->d_out[i]); h->d_out[i] = nullptr;
    }
    h->host_ready = false;
}

static int host_init(b2a_handle* h, size_t in_clip, size_t out_clip) {
    // ~64 MiB of input per chunk, two chunks in flight.  mel / mfcc kernels are persistent (one clip
    // per CTA at a time, grid = min(n, grid_cap)): a chunk is a whole number of waves of grid_cap clips.
    int64_t cc = (int64_t)((64u << 20) / in_clip);
    cc = std::max<int64_t>(1, std::min<int64_t>(cc, 4096));
    if (h->cfg.kind == B2A_KIND_CQT) cc = std::max<int64_t>(cc, 1024);   // one full CQT chunk per copy
    else if (h->grid_cap > 0 && cc > h->grid_cap) cc = ((cc + h->grid_cap - 1) / h->grid_cap) * h->grid_cap;
    h->chunk_clips = cc;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_kernels[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_in[i], in_clip * cc);
        if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_out[i], out_clip * cc);
    }
    if (e != cudaSuccess) {
        host_teardown(h);            // the handle stays usable: the next call starts from scratch
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,
                    std::string("run_host resources: ") + cudaGetErrorString(e));
    }
    h->host_ready = true;
    return B2A_OK;
}

static int run_host_impl(b2a_handle* h, const void* clips, int64_t n_clips, float* out, bool copy_only);

int b2a_run_host(b2a_handle* h, const void* clips, int64_t n_clips, float* out) {
    return run_host_impl(h, clips, n_clips, out, false);
}

// Diagnostic: b2a_run_host's exact copy schedule (same chunks, streams, device buffers) with the kernels
// left out; `out` receives whatever the device buffers hold.  The time of this call is the host<->device
// transfer ceiling of the host path on this box (bench.py reports e2e against it).
int b2a_run_host_copy_only(b2a_handle* h, const void* clips, int64_t n_clips, float* out) {
    return run_host_impl(h, clips, n_clips, out, true);
}

static int run_host_impl(b2a_handle* h, const void* clips, int64_t n_clips, float* out, bool copy_only) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (n_clips < 0) return fail(B2A_EINVAL, "n_clips < 0");
    if (n_clips == 0) { h->last_launches = 0; return B2A_OK; }
    if (!clips || !out) return fail(B2A_EINVAL, "NULL buffer");
    CU_TRY(cudaSetDevice(h->device));
    const size_t in_clip = (size_t)h->cfg.n_samples * h->in_elem;
    const size_t out_clip = (size_t)h->rows * h->frames * sizeof(float);
    if (!h->host_ready) {
        const int irc = host_init(h, in_clip, out_clip);
        if (irc != B2A_OK) return irc;
    }
    h->last_launches = 0;
    int rc = B2A_OK;
    cudaError_t ce = cudaSuccess;
    const char* what = "";
    // Every failure inside the loop leaves through the common exit below, which drains both streams:
    // copies into the caller's `out` that were already enqueued must not outlive this call.
#define CU_STEP(expr) { ce = (expr); if (ce != cudaSuccess) { what = #expr; break; } }
    int64_t done = 0;
    for (int c = 0; done < n_clips; ++c) {
        const int s = c & 1;
        const int64_t nb = std::min<int64_t>(h->chunk_clips, n_clips - done);
        const unsigned char* src = (const unsigned char*)clips + (size_t)done * in_clip;
        float* dst = (float*)((unsigned char*)out + (size_t)done * out_clip);
        CU_STEP(cudaMemcpyAsync(h->d_in[s], src, in_clip * nb, cudaMemcpyHostToDevice, h->streams[s]));
        // The two streams overlap copies with kernels, but the kernels of consecutive chunks must not
        // overlap each other when they work in the handle's scratch (mfcc: raw-dB rows per CTA index;
        // cqt: decimated signals and per-clip extrema): chunk c+1's kernels wait for chunk c's.
        const bool shared_scratch = h->cfg.kind != B2A_KIND_MEL;
        if (shared_scratch && c > 0) CU_STEP(cudaStreamWaitEvent(h->streams[s], h->ev_kernels[s ^ 1], 0));
        if (!copy_only) rc = run_device_impl(h, h->d_in[s], nb, h->d_out[s], h->streams[s], &h->last_launches);
        if (rc != B2A_OK) break;
        if (shared_scratch) CU_STEP(cudaEventRecord(h->ev_kernels[s], h->streams[s]));
        CU_STEP(cudaMemcpyAsync(dst, h->d_out[s], out_clip * nb, cudaMemcpyDeviceToHost, h->streams[s]));
        done += nb;
    }
#undef CU_STEP
    const cudaError_t e0 = cudaStreamSynchronize(h->streams[0]);
    const cudaError_t e1 = cudaStreamSynchronize(h->streams[1]);
    if (rc != B2A_OK) return rc;                     // run_device_impl already set the message
    if (ce != cudaSuccess)
        return fail(ce == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    CU_TRY(e0);
    CU_TRY(e1);
    return B2A_OK;
}

// Files recorded at another rate (deep.py:44-50: librosa.load resamples, then deep.py:52-61 pads / trims):
// raw clips at the file rate -> H2D -> batched resampler writing [chunk][n_samples] float32 rows (trimmed or
// zero-padded to the handle's n_samples) -> the extractor's kernels -> D2H, chunked over the two streams
// exactly like b2a_run_host.  The intermediate rows live in the handle's own input buffers.
int b2a_run_host_resampled(b2a_handle* h, b2a_resampler* r, const void* clips, int32_t in_dtype, int64_t in_stride,
                           const int32_t* in_len, int64_t n_clips, float* out) {
    if (!h || !r) return fail(B2A_EINVAL, "handle/resampler is NULL");
    if (h->cfg.input_dtype != B2A_IN_F32) return fail(B2A_EINVAL, "run_host_resampled needs a float32-input handle (the resampler's output type)");
    if (in_dtype != B2A_IN_I16 && in_dtype != B2A_IN_F32) return fail(B2A_EINVAL, "in_dtype");
    if (n_clips < 0 || in_stride <= 0) return fail(B2A_EINVAL, "bad batch geometry");
    int32_t orig = 0, target = 0, rdev = 0;
    b2a_resampler_rates(r, &orig, &target, &rdev);
    if (target != h->cfg.sample_rate || rdev != h->device)
        return fail(B2A_EINVAL, "resampler target rate / device do not match the handle");
    if (n_clips == 0) { h->last_launches = 0; return B2A_OK; }
    if (!clips || !in_len || !out) return fail(B2A_EINVAL, "NULL buffer");
    for (int64_t i = 0; i < n_clips; ++i)
        if (in_len[i] < 0 || in_len[i] > in_stride) return fail(B2A_EINVAL, "in_len outside [0, in_stride]");
    CU_TRY(cudaSetDevice(h->device));
    const size_t mid_clip = (size_t)h->cfg.n_samples * sizeof(float);
    const size_t out_clip = (size_t)h->rows * h->frames * sizeof(float);
    const size_t raw_elem = in_dtype == B2A_IN_I16 ? 2 : 4;
    if (!h->host_ready) {
        const int irc = host_init(h, mid_clip, out_clip);
        if (irc != B2A_OK) return irc;
    }
    // raw chunks of <= 128 MiB: not more clips than the intermediate buffers hold
    int64_t cc = std::max<int64_t>(1, (int64_t)((128u << 20) / ((size_t)in_stride * raw_elem)));
    cc = std::min<int64_t>(cc, h->chunk_clips);
    for (int i = 0; i < 2; ++i) {
        const size_t need = (size_t)cc * in_stride * raw_elem;
        if (need > h->rs_raw_cap[i]) {
            cudaFree(h->rs_raw[i]); h->rs_raw[i] = nullptr; h->rs_raw_cap[i] = 0;
            CU_TRY(cudaMalloc(&h->rs_raw[i], need));
            h->rs_raw_cap[i] = need;
        }
        if (!h->rs_len[i]) CU_TRY(cudaMalloc((void**)&h->rs_len[i], (size_t)h->chunk_clips * sizeof(int)));
    }
    h->last_launches = 0;
    int rc = B2A_OK;
    cudaError_t ce = cudaSuccess;
    const char* what = "";
#define CU_STEP(expr) { ce = (expr); if (ce != cudaSuccess) { what = #expr; break; } }
    int64_t done = 0;
    for (int c = 0; done < n_clips; ++c) {
        const int s = c & 1;
        const int64_t nb = std::min<int64_t>(cc, n_clips - done);
        const unsigned char* src = (const unsigned char*)clips + (size_t)done * in_stride * raw_elem;
        float* dst = (float*)((unsigned char*)out + (size_t)done * out_clip);
        CU_STEP(cudaMemcpyAsync(h->rs_raw[s], src, (size_t)nb * in_stride * raw_elem, cudaMemcpyHostToDevice, h->streams[s]));
        CU_STEP(cudaMemcpyAsync(h->rs_len[s], in_len + done, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, h->streams[s]));
        rc = b2a_resampler_run_device_batch(r, h->rs_raw[s], in_dtype, nb, in_stride, h->rs_len[s], (float*)h->d_in[s],
                                            h->cfg.n_samples, h->cfg.n_samples, h->streams[s]);
        if (rc != B2A_OK) break;
        h->last_launches += 1;
        const bool shared_scratch = h->cfg.kind != B2A_KIND_MEL;
        if (shared_scratch && c > 0) CU_STEP(cudaStreamWaitEvent(h->streams[s], h->ev_kernels[s ^ 1], 0));
        rc = run_device_impl(h, h->d_in[s], nb, h->d_out[s], h->streams[s], &h->last_launches);
        if (rc != B2A_OK) break;
        if (shared_scratch) CU_STEP(cudaEventRecord(h->ev_kernels[s], h->streams[s]));
        CU_STEP(cudaMemcpyAsync(dst, h->d_out[s], out_clip * nb, cudaMemcpyDeviceToHost, h->streams[s]));
        done += nb;
    }
#undef CU_STEP
    const cudaError_t e0 = cudaStreamSynchronize(h->streams[0]);
    const cudaError_t e1 = cudaStreamSynchronize(h->streams[1]);
    if (rc != B2A_OK) return rc;
    if (ce != cudaSuccess)
        return fail(ce == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    CU_TRY(e0);
    CU_TRY(e1);
    return B2A_OK;
}

int b2a_run_device_ragged(b2a_handle* h, const void* d_clips, const int64_t* d_in_offsets,
                          const int32_t* d_lengths, const int64_t* d_out_offsets, int64_t n_clips,
                          float* d_out, void* stream) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (h->cfg.kind == B2A_KIND_CQT) return fail(B2A_EINVAL, "ragged batches: mel, mfcc and classical handles only");
    if (n_clips < 0) return fail(B2A_EINVAL, "n_clips < 0");
    if (n_clips > 0 && (!d_clips || !d_out || !d_in_offsets || !d_lengths || !d_out_offsets))
        return fail(B2A_EINVAL, "NULL buffer");
    CU_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
    return run_device_impl(h, d_clips, n_clips, d_out, (cudaStream_t)stream, &h->last_launches,
                           (const long long*)d_in_offsets, (const int*)d_lengths, (const long long*)d_out_offsets);
}

int b2a_run_host_ragged(b2a_handle* h, const void* clips, int64_t total_in, const int64_t* in_offsets,
                        const int32_t* lengths, const int64_t* out_offsets, int64_t n_clips,
                        float* out, int64_t total_out) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (h->cfg.kind == B2A_KIND_CQT) return fail(B2A_EINVAL, "ragged batches: mel, mfcc and classical handles only");
    if (n_clips < 0 || total_in < 0 || total_out < 0) return fail(B2A_EINVAL, "negative size");
    if (n_clips == 0) { h->last_launches = 0; return B2A_OK; }
    if (!clips || !out || !in_offsets || !lengths || !out_offsets) return fail(B2A_EINVAL, "NULL buffer");
    const int hop = h->cfg.hop_length;
    const bool classical = h->cfg.kind == B2A_KIND_CLASSICAL;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t L = lengths[i];
        if (L < h->cfg.n_fft || L > h->cfg.n_samples) return fail(B2A_EINVAL, "clip length outside [n_fft, cfg.n_samples]");
        if (classical && L < 8 * hop) return fail(B2A_EINVAL, "classical: clip shorter than 8 hops (classical.py:262-270 pads to it)");
        if (in_offsets[i] < 0 || in_offsets[i] + L > total_in) return fail(B2A_EINVAL, "input offset out of range");
        const int64_t need = classical ? (int64_t)h->rows : (int64_t)h->rows * (1 + L / hop);
        if (out_offsets[i] < 0 || out_offsets[i] + need > total_out) return fail(B2A_EINVAL, "output offset out of range");
    }
    CU_TRY(cudaSetDevice(h->device));
    // Per-handle scratch, grown on demand and kept (no allocation in the steady state), on the handle's
    // own non-blocking stream.  Offsets and lengths travel in one block: [in_off | out_off | len].
    if (!h->rag_stream) CU_TRY(cudaStreamCreateWithFlags(&h->rag_stream, cudaStreamNonBlocking));
    auto grow = [](void** p, size_t* cap, size_t need) -> cudaError_t {
        if (need <= *cap) return cudaSuccess;
        cudaFree(*p); *p = nullptr; *cap = 0;
        const size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(p, want);
        if (e == cudaSuccess) *cap = want;
        return e;
    };
    const size_t meta_bytes = (size_t)n_clips * (8 + 8 + 4);
    CU_TRY(grow(&h->rag_in, &h->rag_in_cap, (size_t)total_in * h->in_elem));
    CU_TRY(grow((void**)&h->rag_out, &h->rag_out_cap, (size_t)total_out * sizeof(float)));
    CU_TRY(grow(&h->rag_meta, &h->rag_meta_cap, meta_bytes));
    long long* const d_io = (long long*)h->rag_meta;
    long long* const d_oo = d_io + n_clips;
    int* const d_len = (int*)(d_oo + n_clips);
    cudaStream_t st = h->rag_stream;
    int rc = B2A_OK;
    cudaError_t ce = cudaSuccess;
    const char* what = "";
#define CU_STEP(expr) { ce = (expr); if (ce != cudaSuccess) { what = #expr; break; } }
    do {
        CU_STEP(cudaMemcpyAsync(h->rag_in, clips, (size_t)total_in * h->in_elem, cudaMemcpyHostToDevice, st));
        CU_STEP(cudaMemcpyAsync(d_io, in_offsets, (size_t)n_clips * 8, cudaMemcpyHostToDevice, st));
        CU_STEP(cudaMemcpyAsync(d_oo, out_offsets, (size_t)n_clips * 8, cudaMemcpyHostToDevice, st));
        CU_STEP(cudaMemcpyAsync(d_len, lengths, (size_t)n_clips * 4, cudaMemcpyHostToDevice, st));
        h->last_launches = 0;
        rc = run_device_impl(h, h->rag_in, n_clips, h->rag_out, st, &h->last_launches, d_io, d_len, d_oo);
        if (rc != B2A_OK) break;
        CU_STEP(cudaMemcpyAsync(out, h->rag_out, (size_t)total_out * sizeof(float), cudaMemcpyDeviceToHost, st));
    } while (0);
#undef CU_STEP
    const cudaError_t es = cudaStreamSynchronize(st);     // also on failure: nothing may still touch the caller's buffers
    if (rc != B2A_OK) return rc;
    if (ce != cudaSuccess)
        return fail(ce == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    CU_TRY(es);
    return B2A_OK;
}

int b2a_alloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(B2A_EINVAL, "out is NULL");
    *out = nullptr;
    CU_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return B2A_OK;
}

int b2a_free_pinned(void* p) {
    if (p) CU_TRY(cudaFreeHost(p));
    return B2A_OK;
}

int b2a_get_table(const b2a_handle* h, int32_t which, float* dst, int64_t* count) {
    if (!h || !count) return fail(B2A_EINVAL, "handle/count is NULL");
    std::vector<float> tmp;
    const std::vector<float>* src = nullptr;
    switch (which) {
        case B2A_TABLE_WINDOW: src = &h->window; break;
        case B2A_TABLE_MEL_DENSE: src = &h->mel_dense; break;
        case B2A_TABLE_DCT: src = &h->dct; break;
        case B2A_TABLE_DECIM_TAPS: {
            for (double v : b2a::decimator_taps()) tmp.push_back((float)v);
            src = &tmp; break;
        }
        case B2A_TABLE_CQT_LENGTHS: {
            for (double v : h->cqt.lengths) tmp.push_back((float)v);
            src = &tmp; break;
        }
        case B2A_TABLE_CQT_BASIS: {
            for (const auto& o : h->cqt.oct) tmp.insert(tmp.end(), o.basis.begin(), o.basis.end());
            src = &tmp; break;
        }
        case B2A_TABLE_CHROMA: src = &h->cls.chroma; break;
        case B2A_TABLE_TONNETZ: src = &h->cls.tonnetz; break;
        case B2A_TABLE_CONTRAST_BANDS: {
            for (int v : h->cls.band_start) tmp.push_back((float)v);
            for (int v : h->cls.band_cnt) tmp.push_back((float)v);
            for (int v : h->cls.band_q) tmp.push_back((float)v);
            tmp.push_back((float)h->cls.pip_k0); tmp.push_back((float)h->cls.pip_k1);
            src = &tmp; break;
        }
        default: return fail(B2A_EINVAL, "unknown table");
    }
    const int64_t need = (int64_t)src->size();
    if (dst) {
        if (*count < need) { *count = need; return fail(B2A_EINVAL, "destination too small"); }
        std::memcpy(dst, src->data(), (size_t)need * sizeof(float));
    }
    *count = need;
    return B2A_OK;
}

int b2a_cqt_geometry(const b2a_handle* h, int32_t* n_octaves, int32_t* n_filters, int32_t* n_fft,
                     int32_t* hop, int32_t* sig_len) {
    if (!h) return fail(B2A_EINVAL, "handle is NULL");
    if (h->cfg.kind != B2A_KIND_CQT) return fail(B2A_EINVAL, "not a cqt handle");
    if (n_octaves) *n_octaves = h->cqt.n_octaves;
    if (n_filters) *n_filters = h->cqt.n_filters;
    for (int i = 0; i < h->cqt.n_octaves; ++i) {
        if (n_fft) n_fft[i] = h->cqt.oct[i].n_fft;
        if (hop) hop[i] = h->cqt.oct[i].hop;
        if (sig_len) sig_len[i] = h->cqt.oct[i].sig_len;
    }
    return B2A_OK;
}

}  // extern "C"

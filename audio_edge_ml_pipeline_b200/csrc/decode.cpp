// decode.cpp — batched PCM16 WAV decode straight into a (pinned) fixed-length batch.
//
// Host-side front end of the path: what `_load_segment` + `_pad_or_trim` (reference
// src/preprocessing/feature_extraction/audio/deep.py:30-61) do per clip through librosa.load /
// soundfile, for the common case of mono 16-bit PCM RIFF/WAVE files already at the target rate.
// Anything else (other sample formats, channels, containers, rates) is reported per file through
// `status` and left to the caller (the Python decoder handles more formats, or skips the sample).
// Plain C++17 + POSIX I/O, no CUDA; a small thread pool because one clip is one independent file.
#include "../../include/b2a.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct WavInfo {
    uint16_t tag = 0, ch = 0, align = 0, bits = 0;
    uint32_t sr = 0;
    int64_t data_off = -1, data_size = 0;
};

// Opens `path`, walks the RIFF chunk list up to the data chunk.  Returns the fd (>= 0) or -(B2A_DEC_*).
int open_wav(const char* path, WavInfo* w) {
    const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return -B2A_DEC_EIO;
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return -B2A_DEC_EIO; }
    unsigned char head[12];
    const ssize_t got = ::pread(fd, head, sizeof(head), 0);
    if (got < 12 || std::memcmp(head, "RIFF", 4) != 0 || std::memcmp(head + 8, "WAVE", 4) != 0) {
        ::close(fd);
        return -B2A_DEC_EFORMAT;
    }
    int64_t pos = 12;
    bool have_fmt = false;
    while (pos + 8 <= st.st_size) {
        unsigned char hdr[8 + 40];
        const ssize_t g = ::pread(fd, hdr, sizeof(hdr), pos);
        if (g < 8) break;
        const uint32_t size = rd32(hdr + 4);
        if (std::memcmp(hdr, "fmt ", 4) == 0 && g >= 8 + 16) {
            w->tag = rd16(hdr + 8); w->ch = rd16(hdr + 10); w->sr = rd32(hdr + 12);
            w->align = rd16(hdr + 20); w->bits = rd16(hdr + 22);
            if (w->tag == 0xFFFE && size >= 26 && g >= 8 + 26) w->tag = rd16(hdr + 8 + 24);   // WAVE_FORMAT_EXTENSIBLE
            have_fmt = true;
        } else if (std::memcmp(hdr, "data", 4) == 0) {
            w->data_off = pos + 8;
            w->data_size = std::min<int64_t>(size, st.st_size - w->data_off);
            break;
        }
        pos += 8 + (int64_t)size + (size & 1);
    }
    if (!have_fmt || w->data_off < 0 || w->ch < 1 || w->align < 1 || w->sr == 0) { ::close(fd); return -B2A_DEC_EFORMAT; }
    return fd;
}

bool read_fully(int fd, void* dst, size_t bytes, int64_t off) {
    size_t done = 0;
    while (done < bytes) {
        const ssize_t r = ::pread(fd, (unsigned char*)dst + done, bytes - done, off + (int64_t)done);
        if (r <= 0) return false;
        done += (size_t)r;
    }
    return true;
}

// One sample of any supported format as float32, scaled the way libsndfile hands it to librosa.load:
// PCM16 / 32768, PCM24 / 2^23, PCM32 / 2^31, unsigned PCM8 (x - 128) / 128, IEEE float as is.
inline float sample_f32(const unsigned char* p, int tag, int bits) {
    if (tag == 1) {
        switch (bits) {
            case 16: return (float)(int16_t)rd16(p) * (1.0f / 32768.0f);
            case 8:  return ((float)p[0] - 128.0f) * (1.0f / 128.0f);
            case 24: { int32_t v = p[0] | (p[1] << 8) | (p[2] << 16); v = (v ^ 0x800000) - 0x800000;
                       return (float)((double)v / 8388608.0); }
            default: return (float)((double)(int32_t)rd32(p) / 2147483648.0);     // 32
        }
    }
    if (bits == 32) { float f; std::memcpy(&f, p, 4); return f; }
    double d; std::memcpy(&d, p, 8); return (float)d;
}

bool format_supported(const WavInfo& w) {
    if (w.tag == 1) return (w.bits == 8 || w.bits == 16 || w.bits == 24 || w.bits == 32) && w.align == w.ch * (w.bits / 8);
    if (w.tag == 3) return (w.bits == 32 || w.bits == 64) && w.align == w.ch * (w.bits / 8);
    return false;
}

// General decode of one file at its native rate: frames [start, start + want) -> dst (int16 only for mono
// PCM16, else float32 with the channel mean librosa.to_mono takes), the rest of the row zero-filled.
int decode_general(const char* path, double offset_s, double duration_s, int64_t max_frames, int out_dtype,
                   void* dst, int32_t* rate, int32_t* n_out, int32_t* channels) {
    const size_t esz = out_dtype == B2A_IN_I16 ? 2 : 4;
    std::memset(dst, 0, (size_t)max_frames * esz);
    *rate = 0; *n_out = 0;
    if (channels) *channels = 0;
    WavInfo w;
    const int fd = open_wav(path, &w);
    if (fd < 0) return -fd;
    *rate = (int32_t)w.sr;
    if (channels) *channels = w.ch;
    if (!format_supported(w)) { ::close(fd); return B2A_DEC_EUNSUPPORTED; }
    const bool mono16 = w.tag == 1 && w.bits == 16 && w.ch == 1;
    if (out_dtype == B2A_IN_I16 && !mono16) { ::close(fd); return B2A_DEC_EUNSUPPORTED; }
    const int64_t n_frames = w.data_size / w.align;
    int64_t start = std::min<int64_t>((int64_t)(offset_s * (double)w.sr), n_frames);
    if (start < 0) start = 0;
    int64_t stop = n_frames;
    if (duration_s >= 0) stop = std::min<int64_t>(n_frames, start + (int64_t)(duration_s * (double)w.sr));
    const int64_t want = std::min<int64_t>(std::max<int64_t>(stop - start, 0), max_frames);
    bool ok = true;
    if (out_dtype == B2A_IN_I16) {
        ok = read_fully(fd, dst, (size_t)want * 2, w.data_off + start * 2);
    } else {
        constexpr int64_t kBlock = 16384;                       // frames per read
        std::vector<unsigned char> buf((size_t)std::min<int64_t>(kBlock, std::max<int64_t>(want, 1)) * w.align);
        float* out = (float*)dst;
        const int bps = w.bits / 8;
        const float inv_ch = 1.0f;                              // (true division below, as numpy's mean)
        (void)inv_ch;
        for (int64_t f0 = 0; f0 < want && ok; f0 += kBlock) {
            const int64_t nf = std::min<int64_t>(kBlock, want - f0);
            ok = read_fully(fd, buf.data(), (size_t)nf * w.align, w.data_off + (start + f0) * w.align);
            if (!ok) break;
            const unsigned char* p = buf.data();
            if (w.ch == 1) {
                for (int64_t i = 0; i < nf; ++i, p += bps) out[f0 + i] = sample_f32(p, w.tag, w.bits);
            } else {
                for (int64_t i = 0; i < nf; ++i) {
                    float acc = 0.f;                            // np.mean over the channel axis, float32
                    for (int c = 0; c < w.ch; ++c, p += bps) acc += sample_f32(p, w.tag, w.bits);
                    out[f0 + i] = acc / (float)w.ch;
                }
            }
        }
    }
    ::close(fd);
    if (!ok) return B2A_DEC_EIO;
    *n_out = (int32_t)want;
    return B2A_DEC_OK;
}

int decode_one(const char* path, int sample_rate, double offset_s, double duration_s, int n_samples,
               int16_t* dst) {
    std::memset(dst, 0, (size_t)n_samples * sizeof(int16_t));
    const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return B2A_DEC_EIO;
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return B2A_DEC_EIO; }
    unsigned char head[4096];
    const ssize_t got = ::pread(fd, head, sizeof(head), 0);
    if (got < 12 || std::memcmp(head, "RIFF", 4) != 0 || std::memcmp(head + 8, "WAVE", 4) != 0) {
        ::close(fd);
        return B2A_DEC_EFORMAT;
    }
    // walk the chunk list (re-reading a header window when a chunk starts beyond what we hold)
    int64_t pos = 12;
    bool have_fmt = false;
    uint16_t tag = 0, ch = 0, align = 0, bits = 0;
    uint32_t sr = 0;
    int64_t data_off = -1, data_size = 0;
    while (pos + 8 <= st.st_size) {
        unsigned char hdr[8 + 40];
        const ssize_t g = ::pread(fd, hdr, sizeof(hdr), pos);
        if (g < 8) break;
        const uint32_t size = rd32(hdr + 4);
        if (std::memcmp(hdr, "fmt ", 4) == 0 && g >= 8 + 16) {
            tag = rd16(hdr + 8); ch = rd16(hdr + 10); sr = rd32(hdr + 12); align = rd16(hdr + 20); bits = rd16(hdr + 22);
            if (tag == 0xFFFE && size >= 26 && g >= 8 + 26) tag = rd16(hdr + 8 + 24);   // WAVE_FORMAT_EXTENSIBLE
            have_fmt = true;
        } else if (std::memcmp(hdr, "data", 4) == 0) {
            data_off = pos + 8;
            data_size = std::min<int64_t>(size, st.st_size - data_off);
            break;
        }
        pos += 8 + (int64_t)size + (size & 1);
    }
    if (!have_fmt || data_off < 0) { ::close(fd); return B2A_DEC_EFORMAT; }
    if (tag != 1 || bits != 16 || ch != 1 || align != 2) { ::close(fd); return B2A_DEC_EUNSUPPORTED; }
    if ((int)sr != sample_rate) { ::close(fd); return B2A_DEC_ERATE; }
    const int64_t n_frames = data_size / 2;
    // librosa.load(offset, duration): seek int(offset*sr) frames, read int(duration*sr) frames
    int64_t start = std::min<int64_t>((int64_t)(offset_s * (double)sr), n_frames);
    if (start < 0) start = 0;
    int64_t stop = n_frames;
    if (duration_s >= 0) stop = std::min<int64_t>(n_frames, start + (int64_t)(duration_s * (double)sr));
    const int64_t want = std::min<int64_t>(std::max<int64_t>(stop - start, 0), n_samples);
    int64_t done = 0;
    while (done < want) {
        const ssize_t r = ::pread(fd, (unsigned char*)dst + done * 2, (size_t)(want - done) * 2, data_off + (start + done) * 2);
        if (r <= 0) break;
        done += r / 2;
    }
    ::close(fd);
    return done == want ? B2A_DEC_OK : B2A_DEC_EIO;
}

}  // namespace

extern "C" int b2a_decode_wav_pcm16_batch(const char* const* paths, int64_t n_files, int32_t sample_rate,
                                          const double* offset_s, const double* duration_s,
                                          int32_t n_samples, int16_t* dst, int32_t* status,
                                          int32_t n_threads) {
    if (n_files < 0 || n_samples <= 0 || (n_files > 0 && (!paths || !dst || !status))) return B2A_EINVAL;
    int nt = n_threads > 0 ? n_threads : (int)std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()));
    nt = (int)std::min<int64_t>(nt, std::max<int64_t>(n_files, 1));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_files) return;
            status[i] = decode_one(paths[i], sample_rate, offset_s ? offset_s[i] : 0.0,
                                   duration_s ? duration_s[i] : -1.0, n_samples, dst + (size_t)i * n_samples);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return B2A_OK;
}

namespace {
void parallel_for(int64_t n, int n_threads, const std::function<void(int64_t)>& fn) {
    int nt = n_threads > 0 ? n_threads : (int)std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()));
    nt = (int)std::min<int64_t>(nt, std::max<int64_t>(n, 1));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) return;
            fn(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}
}  // namespace

extern "C" int b2a_probe_wav_batch(const char* const* paths, int64_t n_files, int32_t* rate, int32_t* channels,
                                   int32_t* bits, int32_t* format_tag, int64_t* n_frames, int32_t* status,
                                   int32_t n_threads) {
    if (n_files < 0 || (n_files > 0 && (!paths || !status))) return B2A_EINVAL;
    parallel_for(n_files, n_threads, [&](int64_t i) {
        WavInfo w;
        const int fd = open_wav(paths[i], &w);
        if (fd >= 0) ::close(fd);
        status[i] = fd < 0 ? -fd : (format_supported(w) ? B2A_DEC_OK : B2A_DEC_EUNSUPPORTED);
        const bool have = fd >= 0;
        if (rate) rate[i] = have ? (int32_t)w.sr : 0;
        if (channels) channels[i] = have ? w.ch : 0;
        if (bits) bits[i] = have ? w.bits : 0;
        if (format_tag) format_tag[i] = have ? w.tag : 0;
        if (n_frames) n_frames[i] = have ? w.data_size / w.align : 0;
    });
    return B2A_OK;
}

extern "C" int b2a_decode_wav_batch(const char* const* paths, int64_t n_files, const double* offset_s,
                                    const double* duration_s, int64_t max_frames, int32_t out_dtype, void* dst,
                                    int64_t dst_stride, int32_t* rate, int32_t* n_out, int32_t* status,
                                    int32_t n_threads) {
    if (n_files < 0 || max_frames <= 0 || dst_stride < max_frames || (out_dtype != B2A_IN_I16 && out_dtype != B2A_IN_F32) ||
        (n_files > 0 && (!paths || !dst || !status || !rate || !n_out)))
        return B2A_EINVAL;
    const size_t esz = out_dtype == B2A_IN_I16 ? 2 : 4;
    parallel_for(n_files, n_threads, [&](int64_t i) {
        status[i] = decode_general(paths[i], offset_s ? offset_s[i] : 0.0, duration_s ? duration_s[i] : -1.0, max_frames,
                                   out_dtype, (unsigned char*)dst + (size_t)i * (size_t)dst_stride * esz, &rate[i], &n_out[i],
                                   nullptr);
    });
    return B2A_OK;
}

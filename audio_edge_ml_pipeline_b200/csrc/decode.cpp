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
#include <thread>
#include <vector>

namespace {

inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

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

// tables.h — host-side constant tables, built in double precision at b2a_create().
//
// Each builder restates the published formula of the librosa 0.11.0 function the reference
// calls on this path (reference call sites: src/preprocessing/feature_extraction/audio/deep.py
// :126-133, :249-259, :318-324).  They are independent re-implementations in C++; the tests
// compare them against the numpy oracle through b2a_get_table().
#pragma once
#include <cstdint>
#include <vector>

namespace b2a {

// scipy.signal.get_window("hann", n, fftbins=True): 0.5 - 0.5 cos(2 pi i / n)
std::vector<float> hann_periodic(int n);

// exp(-2 pi i k / n), k = 0..count-1, interleaved re,im (float32, evaluated in double)
std::vector<float> twiddles(int n, int count);

// librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm="slaney",
// dtype=float32) -> dense [n_mels][1+n_fft/2]
std::vector<float> mel_filterbank(int sr, int n_fft, int n_mels);

// Banded form of a dense filterbank: per band first bin, count, offset into `w`.
struct BandedMel {
    std::vector<int32_t> k0, cnt, off;
    std::vector<float> w;
    int max_cnt = 0;
};
BandedMel band_mel(const std::vector<float>& dense, int n_mels, int n_bins);

// scipy.fft.dct(type=2, norm="ortho") as a matrix [n_out][n_in] (export_svm.py:69-79)
std::vector<float> dct2_ortho(int n_out, int n_in);

// 2:1 decimator standing in for soxr_hq (see DESIGN.md "CQT decimator"): 383-tap Kaiser sinc
constexpr int kDecimTaps = 383;
std::vector<double> decimator_taps();

// Rational resampler standing in for soxr_hq inside librosa.load (deep.py:44-50): the 2:1
// decimator's specification (pass band to 0.913 x, stop band from 1.0 x the lower Nyquist,
// 125 dB, linear phase, zero-extended edges) scaled to up/down = target/orig.  The prototype
// low-pass runs at up x orig Hz with half-length round(95.5 * up * orig / min(orig, target)),
// so 2:1 gives exactly decimator_taps().  poly[p][i] = up * g[p + up * i]  (float32, rows padded to
// a multiple of 4 taps):
//   y[m] = sum_i poly[(m*down + half_len) % up][i] * x[(m*down + half_len) / up - i]
struct ResamplerDesign {
    int up = 1, down = 1;        // target/orig in lowest terms
    int half_len = 0;            // centre of the prototype (taps = 2 * half_len + 1)
    int taps_per_phase = 0;      // K, multiple of 4
    std::vector<float> poly;     // [up][K]
};
constexpr int kResampleMaxUp = 4096;
bool design_resampler(int orig_sr, int target_sr, ResamplerDesign* d, const char** err);

// ---- CQT plan (librosa.cqt = vqt(gamma=0, intervals="equal")) ---------------------------
struct CqtOctave {
    int n_fft = 0;          // power of two >= longest wavelet of this octave
    int hop = 0;            // hop at this octave's rate
    int sig_len = 0;        // samples of the (decimated) signal feeding this octave
    double sr = 0;          // rate of that signal
    bool decimate_after = false;
    int row0 = 0;           // first output row (bin index) of this octave
    int n_rows = 0;         // filters in this octave that land inside [0, n_bins)
    int filt0 = 0;          // index of the first used filter inside the octave's basis
    // dense complex basis [n_filters][1+n_fft/2] (re,im), already sparsified (zeros outside
    // the kept support), scaled by lengths/n_fft and sqrt(sr0/sr_o)
    std::vector<float> basis;
};
struct CqtPlan {
    int n_octaves = 0, n_filters = 0, n_early = 0, n_frames = 0;
    std::vector<CqtOctave> oct;
    std::vector<double> lengths;   // wavelet lengths at the (early-downsampled) base rate
    std::vector<double> freqs;
};
// returns false and fills err on invalid configuration (e.g. filter cut-off above Nyquist)
bool build_cqt_plan(int sr, int hop, int n_bins, int bpo, double fmin, int n_samples,
                    CqtPlan* plan, const char** err);

// ---- audio_classical (classical.py:272-355) --------------------------------------------------------
// librosa.filters.chroma(sr, n_fft, tuning, n_chroma=12, ctroct=5, octwidth=2, norm=2, base_c=True) for the 100
// tunings estimate_tuning can return (np.linspace(-0.5, 0.5, 101)[:-1]), the spectral_contrast bands
// (n_bands=6, fmin=200, quantile=0.02), the tonnetz projection and the piptrack bin range.
struct ClassicalTables {
    std::vector<float> chroma;                       // [100][12][1 + n_fft/2]
    std::vector<int32_t> band_start, band_cnt, band_q;
    std::vector<float> tonnetz;                      // [6][12]
    int pip_k0 = 0, pip_k1 = 0;
};
bool build_classical_tables(int sr, int n_fft, ClassicalTables* t, const char** err);

}  // namespace b2a

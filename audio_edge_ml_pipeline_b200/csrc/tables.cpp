// tables.cpp — see tables.h.  Plain C++17, no CUDA.
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>

namespace b2a {

static const double kPi = 3.14159265358979323846;

std::vector<float> hann_periodic(int n) {
    std::vector<float> w(n);
    for (int i = 0; i < n; ++i) w[i] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * i / n));
    return w;
}

std::vector<float> twiddles(int n, int count) {
    std::vector<float> t(2 * (size_t)count);
    for (int k = 0; k < count; ++k) {
        double a = -2.0 * kPi * k / n;
        t[2 * k] = (float)std::cos(a);
        t[2 * k + 1] = (float)std::sin(a);
    }
    return t;
}

// ---- Slaney mel scale (librosa.hz_to_mel / mel_to_hz, htk=False) -------------------------
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

std::vector<float> mel_filterbank(int sr, int n_fft, int n_mels) {
    const int n_bins = 1 + n_fft / 2;
    const double fmin = 0.0, fmax = (double)sr / 2;
    std::vector<double> fftfreqs(n_bins);
    const double val = 1.0 / (n_fft * (1.0 / sr));           // np.fft.rfftfreq
    for (int k = 0; k < n_bins; ++k) fftfreqs[k] = k * val;
    std::vector<double> mel_f(n_mels + 2);
    const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
    const double step = (m1 - m0) / (n_mels + 1);            // np.linspace
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : i * step + m0);
    std::vector<float> w((size_t)n_mels * n_bins, 0.f);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < n_bins; ++k) {
            const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
            const float tri = (float)std::max(0.0, std::min(lower, upper));   // stored float32
            w[(size_t)i * n_bins + k] = (float)((double)tri * enorm);          // weights *= enorm
        }
    }
    return w;
}

BandedMel band_mel(const std::vector<float>& dense, int n_mels, int n_bins) {
    BandedMel b;
    b.k0.resize(n_mels); b.cnt.resize(n_mels); b.off.resize(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int lo = n_bins, hi = -1;
        for (int k = 0; k < n_bins; ++k)
            if (dense[(size_t)m * n_bins + k] != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k); }
        b.off[m] = (int32_t)b.w.size();
        if (hi < 0) { b.k0[m] = 0; b.cnt[m] = 0; continue; }
        b.k0[m] = lo; b.cnt[m] = hi - lo + 1;
        for (int k = lo; k <= hi; ++k) b.w.push_back(dense[(size_t)m * n_bins + k]);
        b.max_cnt = std::max(b.max_cnt, hi - lo + 1);
    }
    return b;
}

std::vector<float> dct2_ortho(int n_out, int n_in) {
    std::vector<float> d((size_t)n_out * n_in);
    for (int k = 0; k < n_out; ++k)
        for (int n = 0; n < n_in; ++n) {
            double v = std::cos(kPi / n_in * (n + 0.5) * k) * std::sqrt(2.0 / n_in);
            if (k == 0) v /= std::sqrt(2.0);
            d[(size_t)k * n_in + n] = (float)v;
        }
    return d;
}

// ---- decimator ---------------------------------------------------------------------------
static double bessel_i0(double x) {
    double s = 1.0, t = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        t *= q / ((double)k * k);
        s += t;
        if (t < 1e-18 * s) break;
    }
    return s;
}

std::vector<double> decimator_taps() {
    const int n = kDecimTaps;
    const double atten = 125.0, pass = 0.913, stop = 1.0;
    const double beta = 0.1102 * (atten - 8.7);
    const double fc = 0.5 * (pass + stop) * 0.25;
    std::vector<double> h(n);
    double sum = 0;
    const double i0b = bessel_i0(beta);
    for (int i = 0; i < n; ++i) {
        const double m = i - (n - 1) / 2.0;
        const double x = 2.0 * fc * m;
        const double sinc = (x == 0.0) ? 1.0 : std::sin(kPi * x) / (kPi * x);
        const double r = 2.0 * m / (n - 1);
        const double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
        h[i] = 2.0 * fc * sinc * w;
        sum += h[i];
    }
    for (auto& v : h) v /= sum;
    return h;
}

// ---- rational resampler ------------------------------------------------------------------
bool design_resampler(int orig_sr, int target_sr, ResamplerDesign* d, const char** err) {
    if (orig_sr <= 0 || target_sr <= 0) { *err = "resampler: sample rates must be positive"; return false; }
    int a = orig_sr, b = target_sr;
    while (b) { const int t = a % b; a = b; b = t; }
    d->up = target_sr / a;
    d->down = orig_sr / a;
    if (d->up > kResampleMaxUp) { *err = "resampler: target/orig reduces to an up-factor above 4096"; return false; }
    const double lo = (double)std::min(orig_sr, target_sr);
    const double fs_up = (double)d->up * (double)orig_sr;            // rate the prototype runs at
    d->half_len = (int)std::llround(95.5 * fs_up / lo);
    const long long ntot = 2LL * d->half_len + 1;
    if (ntot > (1LL << 26)) { *err = "resampler: filter too long for this ratio"; return false; }
    const double atten = 125.0, pass = 0.913, stop = 1.0;
    const double beta = 0.1102 * (atten - 8.7);
    const double fc = 0.5 * (pass + stop) * 0.5 * lo / fs_up;        // cycles per prototype sample
    std::vector<double> g((size_t)ntot);
    const double i0b = bessel_i0(beta);
    double sum = 0;
    for (long long i = 0; i < ntot; ++i) {
        const double m = (double)(i - d->half_len);
        const double x = 2.0 * fc * m;
        const double sinc = (x == 0.0) ? 1.0 : std::sin(kPi * x) / (kPi * x);
        const double r = m / (double)d->half_len;
        const double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
        g[(size_t)i] = 2.0 * fc * sinc * w;
        sum += g[(size_t)i];
    }
    const int K = (int)(((ntot + d->up - 1) / d->up + 3) & ~3LL);
    d->taps_per_phase = K;
    d->poly.assign((size_t)d->up * K, 0.f);
    for (int p = 0; p < d->up; ++p)
        for (int i = 0; i < K; ++i) {
            const long long j = p + (long long)d->up * i;
            if (j < ntot) d->poly[(size_t)p * K + i] = (float)((double)d->up * g[(size_t)j] / sum);
        }
    return true;
}

// ---- CQT plan ----------------------------------------------------------------------------
static const double kHannBandwidth = 1.50018310546875;   // librosa window_bandwidth("hann")
static const double kC1 = 32.70319566257483;             // librosa.note_to_hz("C1")

static int num_two_factors(int x) {
    if (x <= 0) return 0;
    int n = 0;
    while (x % 2 == 0) { ++n; x /= 2; }
    return n;
}

// one octave's basis: filters.wavelet (complex64) -> *lengths/n_fft -> FFT -> sparsify 1%
static void octave_basis(const double* freqs, const double* alpha, int nf, double sr,
                         double scale, CqtOctave* o) {
    std::vector<double> len(nf);
    double max_len = 0;
    for (int i = 0; i < nf; ++i) { len[i] = (1.0 / alpha[i]) * sr / freqs[i]; max_len = std::max(max_len, len[i]); }
    const int n_fft = (int)std::pow(2.0, std::ceil(std::log2(max_len)));
    const int n_bins = n_fft / 2 + 1;
    o->n_fft = n_fft;
    o->basis.assign((size_t)nf * n_bins * 2, 0.f);
    std::vector<std::complex<float>> filt(n_fft);
    std::vector<std::complex<float>> G(n_bins);
    std::vector<float> mags(n_bins), srt(n_bins);
    for (int i = 0; i < nf; ++i) {
        const double ilen = len[i];
        const double start = std::floor(-ilen / 2.0), stop = std::floor(ilen / 2.0);
        const int L = (int)(stop - start);                    // np.arange(-ilen//2, ilen//2)
        std::vector<std::complex<double>> sig(L);
        double l1 = 0;
        for (int n = 0; n < L; ++n) {
            const double t = start + n;
            const double ph = t * 2 * kPi * freqs[i] / sr;
            const double w = 0.5 - 0.5 * std::cos(2.0 * kPi * n / L);   // periodic Hann(L)
            sig[n] = std::complex<double>(std::cos(ph), std::sin(ph)) * w;
            l1 += std::abs(sig[n]);
        }
        std::fill(filt.begin(), filt.end(), std::complex<float>(0.f, 0.f));
        const int lpad = (n_fft - L) / 2;                     // util.pad_center
        for (int n = 0; n < L; ++n) {
            const std::complex<double> v = sig[n] / l1;       // util.normalize(norm=1)
            filt[lpad + n] = std::complex<float>((float)v.real(), (float)v.imag());   // complex64
        }
        const double rs = ilen / (double)n_fft;               // basis *= lengths / n_fft
        for (int n = 0; n < n_fft; ++n) {
            const std::complex<double> v(filt[n].real() * rs, filt[n].imag() * rs);
            filt[n] = std::complex<float>((float)v.real(), (float)v.imag());
        }
        // DFT (double accumulation of the float32-valued filter), keep bins 0..n_fft/2
        for (int k = 0; k < n_bins; ++k) {
            std::complex<double> acc(0, 0);
            for (int n = lpad; n < lpad + L; ++n) {
                const double a = -2.0 * kPi * (double)(((int64_t)k * n) % n_fft) / n_fft;
                acc += std::complex<double>(filt[n].real(), filt[n].imag()) *
                       std::complex<double>(std::cos(a), std::sin(a));
            }
            G[k] = std::complex<float>((float)acc.real(), (float)acc.imag());
        }
        // util.sparsify_rows(quantile=0.01), float32 arithmetic as numpy does on complex64
        float norm = 0.f;
        {
            double s = 0;
            for (int k = 0; k < n_bins; ++k) { mags[k] = std::hypot(G[k].real(), G[k].imag()); s += mags[k]; }
            norm = (float)s;
        }
        srt = mags;
        std::sort(srt.begin(), srt.end());
        float cum = 0.f;
        int j = 0;
        for (int k = 0; k < n_bins; ++k) {
            cum += srt[k] / norm;
            if (!(cum < 0.01f)) { j = k; break; }
        }
        const float thr = srt[j];
        for (int k = 0; k < n_bins; ++k)
            if (mags[k] >= thr) {
                o->basis[((size_t)i * n_bins + k) * 2 + 0] = (float)((double)G[k].real() * scale);
                o->basis[((size_t)i * n_bins + k) * 2 + 1] = (float)((double)G[k].imag() * scale);
            }
    }
}

bool build_cqt_plan(int sr_in, int hop_in, int n_bins, int bpo, double fmin, int n_samples,
                    CqtPlan* plan, const char** err) {
    if (n_bins < 2 || bpo < 1 || hop_in < 1) { *err = "cqt: need n_bins >= 2, bins_per_octave >= 1, hop >= 1"; return false; }
    if (fmin <= 0) fmin = kC1;
    const int n_oct = (n_bins + bpo - 1) / bpo;
    const int n_filters = std::min(bpo, n_bins);
    std::vector<double> freqs(n_bins), alpha(n_bins), logf(n_bins);
    for (int b = 0; b < n_bins; ++b) { freqs[b] = fmin * std::pow(2.0, (double)b / bpo); logf[b] = std::log2(freqs[b]); }
    for (int b = 0; b < n_bins; ++b) {
        double r;
        if (b == 0) r = 1.0 / (logf[1] - logf[0]);
        else if (b == n_bins - 1) r = 1.0 / (logf[b] - logf[b - 1]);
        else r = 2.0 / (logf[b + 1] - logf[b - 1]);
        const double p = std::pow(2.0, 2.0 / r);
        alpha[b] = (p - 1) / (p + 1);
    }
    double sr = sr_in;
    int hop = hop_in;
    double cutoff = 0;
    for (int b = 0; b < n_bins; ++b) cutoff = std::max(cutoff, freqs[b] * (1 + 0.5 * kHannBandwidth * alpha[b]));
    const double nyq = sr / 2.0;
    if (cutoff > nyq) { *err = "cqt: wavelet basis would exceed the Nyquist frequency; reduce n_bins"; return false; }
    const int c1 = std::max(0, (int)std::ceil(std::log2(nyq / cutoff)) - 1 - 1);
    const int c2 = std::max(0, num_two_factors(hop) - n_oct + 1);
    const int n_early = std::min(c1, c2);
    int len = n_samples;
    for (int e = 0; e < n_early; ++e) { hop /= 2; sr /= 2.0; len = (len + 1) / 2; }
    plan->n_octaves = n_oct; plan->n_filters = n_filters; plan->n_early = n_early;
    plan->freqs = freqs;
    plan->lengths.resize(n_bins);
    for (int b = 0; b < n_bins; ++b) plan->lengths[b] = (1.0 / alpha[b]) * sr / freqs[b];
    plan->oct.assign(n_oct, CqtOctave());
    const double sr0 = sr;
    double my_sr = sr0;
    int my_hop = hop;
    int min_frames = 1 << 30;
    for (int i = 0; i < n_oct; ++i) {
        CqtOctave& o = plan->oct[i];
        const int hi = n_bins - n_filters * i;
        const int lo = std::max(0, n_bins - n_filters * (i + 1));
        o.row0 = lo; o.n_rows = hi - lo; o.filt0 = 0;
        o.hop = my_hop; o.sr = my_sr; o.sig_len = len;
        octave_basis(&freqs[lo], &alpha[lo], hi - lo, my_sr, std::sqrt(sr0 / my_sr), &o);
        min_frames = std::min(min_frames, 1 + len / my_hop);
        o.decimate_after = (my_hop % 2 == 0);
        if (o.decimate_after) { my_hop /= 2; my_sr /= 2.0; len = (len + 1) / 2; }
    }
    plan->n_frames = min_frames;
    return true;
}

// ---- audio_classical tables ---------------------------------------------------------------------------
static std::vector<float> chroma_bank(int sr, int n_fft, double tuning) {
    const int nc = 12, nb = 1 + n_fft / 2;
    std::vector<double> frq(n_fft), bw(n_fft);
    const double a440 = 440.0 * std::pow(2.0, tuning / nc);
    const double step = (double)sr / n_fft;                      // np.linspace(0, sr, n_fft, endpoint=False)
    for (int i = 1; i < n_fft; ++i) frq[i] = nc * std::log2((i * step) / (a440 / 16));
    frq[0] = frq[1] - 1.5 * nc;                                  // "make up a value for the 0 Hz bin"
    for (int i = 0; i + 1 < n_fft; ++i) bw[i] = std::max(frq[i + 1] - frq[i], 1.0);
    bw[n_fft - 1] = 1.0;
    std::vector<double> w((size_t)nc * n_fft);
    const double half = std::nearbyint(nc / 2.0);
    for (int i = 0; i < n_fft; ++i) {
        double nrm = 0.0;
        for (int c = 0; c < nc; ++c) {
            double d = frq[i] - c + half + 10 * nc;
            d = d - nc * std::floor(d / nc) - half;              // np.remainder(., n_chroma) - n_chroma2
            const double g = std::exp(-0.5 * std::pow(2 * d / bw[i], 2));
            w[(size_t)c * n_fft + i] = g;
            nrm += g * g;
        }
        nrm = std::sqrt(nrm);
        if (nrm < 2.2250738585072014e-308) nrm = 1.0;
        const double oct = std::exp(-0.5 * std::pow((frq[i] / nc - 5.0) / 2.0, 2));
        for (int c = 0; c < nc; ++c) w[(size_t)c * n_fft + i] = w[(size_t)c * n_fft + i] / nrm * oct;
    }
    std::vector<float> out((size_t)nc * nb);
    for (int c = 0; c < nc; ++c)                                 // np.roll(wts, -3, axis=0): row c <- row c + 3
        for (int k = 0; k < nb; ++k) out[(size_t)c * nb + k] = (float)w[(size_t)((c + 3) % nc) * n_fft + k];
    return out;
}

bool build_classical_tables(int sr, int n_fft, ClassicalTables* t, const char** err) {
    const int nb = 1 + n_fft / 2, n_bands = 6;
    const double fmin = 200.0, quantile = 0.02;
    const double val = 1.0 / (n_fft * (1.0 / sr));               // np.fft.rfftfreq
    std::vector<double> freq(nb);
    for (int k = 0; k < nb; ++k) freq[k] = k * val;
    // spectral_contrast bands
    std::vector<double> octa(n_bands + 2, 0.0);
    for (int i = 0; i <= n_bands; ++i) octa[i + 1] = fmin * std::pow(2.0, i);
    for (int i = 0; i <= n_bands; ++i)
        if (octa[i] >= 0.5 * sr) { *err = "spectral_contrast: a frequency band exceeds Nyquist (sample_rate too low)"; return false; }
    t->band_start.clear(); t->band_cnt.clear(); t->band_q.clear();
    for (int b = 0; b <= n_bands; ++b) {
        int first = -1, last = -1;
        for (int k = 0; k < nb; ++k)
            if (freq[k] >= octa[b] && freq[k] <= octa[b + 1]) { if (first < 0) first = k; last = k; }
        if (first < 0) { *err = "spectral_contrast: empty frequency band (n_fft too small)"; return false; }
        if (b > 0) first -= 1;
        if (b == n_bands) last = nb - 1;
        const int n_in = last - first + 1;
        int cnt = n_in;
        if (b < n_bands) cnt -= 1;
        if (first < 0 || cnt < 1) { *err = "spectral_contrast: degenerate frequency band"; return false; }
        const int q = (int)std::max(std::nearbyint(quantile * n_in), 1.0);
        if (q > cnt) { *err = "spectral_contrast: quantile exceeds the band"; return false; }
        t->band_start.push_back(first); t->band_cnt.push_back(cnt); t->band_q.push_back(q);
    }
    // piptrack: fmin = 150, fmax = min(4000, sr / 2)
    const double pmin = 150.0, pmax = std::min(4000.0, sr / 2.0);
    t->pip_k0 = nb; t->pip_k1 = 0;
    for (int k = 0; k < nb; ++k)
        if (pmin <= freq[k] && freq[k] < pmax) { t->pip_k0 = std::min(t->pip_k0, k); t->pip_k1 = std::max(t->pip_k1, k + 1); }
    if (t->pip_k0 >= t->pip_k1) { t->pip_k0 = t->pip_k1 = 1; }
    if (t->pip_k0 < 1) t->pip_k0 = 1;                            // (the parabolic shift is defined as 0 on the edge bins)
    if (t->pip_k1 > nb - 1) t->pip_k1 = nb - 1;
    // tonnetz projection
    const double scale[6] = {7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3};
    const double R[6] = {1, 1, 1, 1, 0.5, 0.5};
    t->tonnetz.assign(72, 0.f);
    const double pi = 3.14159265358979323846;
    for (int p = 0; p < 6; ++p)
        for (int c = 0; c < 12; ++c) {
            double v = scale[p] * c;                             // dim_map = linspace(0, 12, 12, endpoint=False) = c
            if ((p & 1) == 0) v -= 0.5;
            t->tonnetz[p * 12 + c] = (float)(R[p] * std::cos(pi * v));
        }
    // chroma banks for np.linspace(-0.5, 0.5, 101)[i], i < 100
    t->chroma.clear();
    t->chroma.reserve((size_t)100 * 12 * nb);
    for (int i = 0; i < 100; ++i) {
        const double tuning = i * 0.01 + (-0.5);
        std::vector<float> fb = chroma_bank(sr, n_fft, tuning);
        t->chroma.insert(t->chroma.end(), fb.begin(), fb.end());
    }
    return true;
}

}  // namespace b2a

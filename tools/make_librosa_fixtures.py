#!/usr/bin/env python
"""Pin the oracle (and through it the CUDA path) to REAL librosa — for anyone who has it installed.

librosa 0.11.0 / soxr 1.0.0 / soundfile (the reference's requirements.txt:55,114,115) are not
installable in the authoring container (no network), so parity of `audio_cqt`, of `audio_mfcc_seq`
and of the load-time resampler against the real library is "unpinned" (DESIGN.md section 2): every
tolerance is measured against oracle/librosa_restated.py.  This script closes that gap where the
libraries exist: it runs the reference's own call sequence (deep.py:44-50, 126-134, 249-260,
318-328) on seeded synthetic clips and writes

    tests/golden/librosa_pin.npz    inputs + librosa outputs + the versions that produced them

tests/test_librosa_pin.py then asserts BOTH the oracle (CPU, `-m "not gpu"`) and the CUDA path
(`-m gpu`) against that file; without the file those tests skip and say why.

    pip install librosa==0.11.0 soxr==1.0.0 soundfile==0.13.1
    python tools/make_librosa_fixtures.py            # a few seconds; commit the .npz it writes
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden" / "librosa_pin.npz"


def _normalize(x: np.ndarray) -> np.ndarray:          # deep.py:64-67
    lo, hi = x.min(), x.max()
    return (x - lo) / (hi - lo + 1e-8)


def main() -> int:
    try:
        import librosa
        import soxr
    except ImportError as exc:
        print(f"make_librosa_fixtures: {exc}; install librosa==0.11.0 soxr==1.0.0 and re-run", file=sys.stderr)
        return 2
    from audio_edge_ml_pipeline_b200 import synth

    out: dict = {"versions": np.array([f"librosa {librosa.__version__}", f"soxr {soxr.__version__}",
                                        f"numpy {np.__version__}"])}
    # --- audio_mel_spec / audio_mfcc_seq: one clip of each synthetic family, 5 s @ 16 kHz -------------
    pcm16 = synth.make_suite(synth.N_FAMILIES, 16000, 80000, seed=1234)
    out["pcm_16k"] = pcm16
    mel, mf13, = [], []
    for c in pcm16:
        y = c.astype(np.float32) / np.float32(32768.0)               # librosa.load of PCM16 (deep.py:44-50)
        m = librosa.feature.melspectrogram(y=y, sr=16000, n_fft=512, hop_length=160, n_mels=40)   # deep.py:126-132
        mel.append(_normalize(librosa.power_to_db(m, ref=np.max)).astype(np.float32))               # :133-134
        # config 2 (n_mels is this project's extension key; librosa.feature.mfcc forwards it)
        f = librosa.feature.mfcc(y=y, sr=16000, n_mfcc=13, n_fft=512, hop_length=160, n_mels=40)
        mf13.append(((f - f.mean(axis=1, keepdims=True)) / (f.std(axis=1, keepdims=True) + 1e-8)).astype(np.float32))
    out["mel_16k"], out["mfcc13_16k"] = np.stack(mel), np.stack(mf13)
    # --- reference defaults at 22.05 kHz: mfcc (40, 216) and cqt (84, 216) ------------------------------
    pcm22 = synth.make_suite(synth.N_FAMILIES, 22050, 110250, seed=1234)
    out["pcm_22k"] = pcm22
    mf40, cq, cq_lin = [], [], []
    for c in pcm22:
        y = c.astype(np.float32) / np.float32(32768.0)
        f = librosa.feature.mfcc(y=y, sr=22050, n_mfcc=40, n_fft=1024, hop_length=512)              # deep.py:318-324
        mf40.append(((f - f.mean(axis=1, keepdims=True)) / (f.std(axis=1, keepdims=True) + 1e-8)).astype(np.float32))
        C = np.abs(librosa.cqt(y, sr=22050, hop_length=512, n_bins=84, bins_per_octave=12,
                               fmin=librosa.note_to_hz("C1")))                                       # deep.py:249-258
        cq_lin.append(C.astype(np.float32))
        cq.append(_normalize(librosa.amplitude_to_db(C, ref=np.max)).astype(np.float32))            # :259-260
    out["mfcc40_22k"], out["cqt_22k"], out["cqt_lin_22k"] = np.stack(mf40), np.stack(cq), np.stack(cq_lin)
    # --- the load-time resampler: librosa.load(sr=16000) of 44.1 kHz / 22.05 kHz / 48 kHz / 8 kHz audio ------
    rng = np.random.default_rng(7)
    for orig in (44100, 22050, 48000, 8000):
        n = int(1.0 * orig)
        t = np.arange(n) / orig
        y = (0.3 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 0.3 * min(orig, 16000) * t)).astype(np.float32)
        out[f"rs_in_{orig}"] = y
        out[f"rs_out_{orig}_16000"] = librosa.resample(y, orig_sr=orig, target_sr=16000, res_type="soxr_hq").astype(np.float32)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes): {sorted(k for k in out)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""SURVEY 8(f) N4 on the CPU: (1) known-answer tests that pin oracle/classical_restated.py (librosa itself is
not installable here: **parity unpinned** against it — pure tones, noise and ramps have closed forms instead),
(2) the C++ table builders of the product (csrc/tables.cpp, compiled for the host with g++) against the
oracle's chroma filterbank / contrast bands / tonnetz projection, (3) the host mirror's column selection
(classical.py:152-207).  No GPU."""
import shutil
import subprocess

import numpy as np
import pytest
import scipy.signal

from audio_edge_ml_pipeline_b200.build import CSRC
from audio_edge_ml_pipeline_b200.extractors import AudioClassicalExtractor
from oracle import classical_restated as C

SR, NFFT, HOP = 22050, 1024, 512


def _tone(freq, secs=2.0, amp=0.5, sr=SR):
    t = np.arange(int(secs * sr)) / sr
    return (amp * np.sin(2 * np.pi * freq * t)).astype(np.float32)


def test_pure_tone_spectral_shape():
    f0 = 100 * SR / NFFT                                   # centre of bin 100
    ff = C.frame_features(_tone(f0))
    mid = slice(3, -3)                                     # frames that do not touch the zero padding
    assert np.allclose(ff["spectral_centroid"][0, mid], f0, rtol=1e-3)      # Hann leakage is symmetric around the bin
    assert np.allclose(ff["spectral_rolloff"][0, mid], f0, atol=SR / NFFT)  # 85 % of |S| is reached inside the main lobe
    assert np.all(ff["spectral_bandwidth"][0, mid] < 2 * SR / NFFT)
    assert np.all(ff["spectral_flatness"][0, mid] < 1e-6)
    assert np.allclose(ff["rms"][0, mid], 0.5 / np.sqrt(2), rtol=1e-3)
    assert np.allclose(ff["zcr"][0, 3:-3], 2 * f0 / SR, atol=2.0 / 2048)    # two crossings per period


def test_white_noise_spectral_shape():
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal(4 * SR)).astype(np.float32)
    ff = C.frame_features(y)
    mid = slice(3, -3)
    assert abs(ff["spectral_centroid"][0, mid].mean() - SR / 4) < 150          # flat spectrum: mean frequency
    assert abs(ff["spectral_rolloff"][0, mid].mean() - 0.85 * SR / 2) < 250
    # |X|^2 of Gaussian noise is exponentially distributed: geometric / arithmetic mean = exp(-gamma) = 0.5615
    assert abs(ff["spectral_flatness"][0, mid].mean() - 0.5615) < 0.03
    assert abs(ff["rms"][0, mid].mean() - 0.1) < 0.003
    assert abs(ff["zcr"][0, mid].mean() - 0.5) < 0.02


def test_chroma_class_and_tuning_of_tones():
    for midi, cls in [(69, 9), (60, 0), (64, 4)]:          # A4, C4, E4 (base_c=True: class 0 is C)
        f = 440.0 * 2 ** ((midi - 69) / 12)
        chroma, tuning = C.chroma_stft(_tone(f), SR, NFFT, HOP, return_tuning=True)
        assert chroma[:, 3:-3].mean(axis=1).argmax() == cls
        assert abs(tuning) <= 0.15                        # parabolic interpolation on the POWER of a Hann main lobe: up to ~0.1 bin of bias
        assert np.isclose(chroma.max(axis=0)[3:-3], 1.0).all()                # norm=inf
    f = 440.0 * 2 ** (0.30 / 12)                          # 30 cents sharp
    _, tuning = C.chroma_stft(_tone(f), SR, NFFT, HOP, return_tuning=True)
    assert abs(tuning - 0.30) <= 0.15
    # silence: no pitch candidates -> tuning 0.0, chroma columns left undivided (all zero)
    chroma, tuning = C.chroma_stft(np.zeros(SR, dtype=np.float32), SR, NFFT, HOP, return_tuning=True)
    assert tuning == 0.0 and not chroma.any()


def test_chroma_filterbank_properties():
    fb = C.chroma_filterbank(SR, NFFT, tuning=0.0)
    assert fb.shape == (12, 513) and fb.dtype == np.float32
    k = int(round(440.0 * NFFT / SR))
    assert fb[:, k].argmax() == 9                          # A
    fb2 = C.chroma_filterbank(SR, NFFT, tuning=0.5)        # half a semitone up: A's bin sits between A and G#
    assert fb2[9, k] < fb[9, k]


def test_tonnetz_matrix_and_projection():
    phi = C.tonnetz_matrix()
    assert phi.shape == (6, 12)
    assert np.allclose(phi[0], np.sin(np.pi * 7 / 6 * np.arange(12)))       # fifths: even rows are sines
    assert np.allclose(phi[1], np.cos(np.pi * 7 / 6 * np.arange(12)))
    assert np.allclose(phi[4], 0.5 * np.sin(np.pi * 2 / 3 * np.arange(12)))
    chroma = np.zeros((12, 3)); chroma[0] = 1.0                               # a lone C
    assert np.allclose(C.tonnetz(chroma)[:, 0], phi[:, 0])


def test_delta_edges_are_constant_and_interior_is_the_ls_slope():
    """What the CUDA kernel relies on: with polyorder == deriv the 'interp' edge fit has a constant derivative, so
    the first / last four frames repeat the nearest interior value; interior weights are i/60 and (3i^2-20)/462."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 40)).astype(np.float32)
    d1, d2 = C.delta(x, 1), C.delta(x, 2)
    i = np.arange(-4, 5)
    for t in range(4, 36):
        assert np.allclose(d1[:, t], (x[:, t - 4:t + 5] * (i / 60.0)).sum(axis=1), atol=1e-5)
        assert np.allclose(d2[:, t], (x[:, t - 4:t + 5] * ((3 * i * i - 20) / 462.0)).sum(axis=1), atol=1e-5)
    for t in range(4):
        assert np.allclose(d1[:, t], d1[:, 4], atol=1e-5) and np.allclose(d1[:, 39 - t], d1[:, 35], atol=1e-5)
        assert np.allclose(d2[:, t], d2[:, 4], atol=1e-5) and np.allclose(d2[:, 39 - t], d2[:, 35], atol=1e-5)
    assert np.allclose(C.delta(np.arange(20, dtype=np.float32)[None] * 0.5, 1), 0.5)     # a ramp's slope


def test_contrast_bands_and_values():
    bands = C.contrast_bands(SR, NFFT)
    assert [len(b) for b, _ in bands] == [9, 9, 19, 37, 74, 149, 216]
    assert [q for _, q in bands] == [1, 1, 1, 1, 2, 3, 4]
    assert bands[0][0][0] == 0 and bands[-1][0][-1] == 512
    for (a, _), (b, _) in zip(bands[:-1], bands[1:]):
        assert b[0] == a[-1] + 1                          # contiguous: the neighbour rule and the dropped last bin cancel
    S = np.abs(np.random.default_rng(1).standard_normal((513, 4))).astype(np.float32)
    got = C.spectral_contrast(S, SR, NFFT)
    b, q = bands[5]
    srt = np.sort(S[b], axis=0)
    want = 10 * np.log10(srt[-q:].mean(axis=0)) - 10 * np.log10(srt[:q].mean(axis=0))
    assert np.allclose(got[5], want, atol=1e-6)


def test_vector_layout_and_dims():
    y = _tone(440.0, secs=1.0) + 0.01 * np.random.default_rng(2).standard_normal(SR).astype(np.float32)
    v = C.audio_classical(y)
    assert v.shape == (302,) and v.dtype == np.float32
    ff = C.frame_features(y)
    assert np.allclose(v[:40], ff["mfcc"].mean(axis=1), rtol=1e-6)
    assert np.allclose(v[40:80], ff["mfcc"].std(axis=1), rtol=1e-6)
    assert np.isclose(v[240], ff["spectral_centroid"].mean()) and np.isclose(v[241], ff["spectral_centroid"].std())
    assert np.allclose(v[262:274], ff["chroma"].mean(axis=1)) and np.allclose(v[296:302], ff["tonnetz"].std(axis=1))
    lean = ["mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_contrast",
            "spectral_flatness", "chroma", "zcr", "rms"]
    assert C.audio_classical(y, aggregations=["mean"]).shape == (151,)        # classical.py:31-41 (docstring counts)
    assert C.audio_classical(y, features=lean).shape == (130,)
    assert C.audio_classical(y, features=lean, aggregations=["mean"]).shape == (65,)


def test_host_mirror_selects_the_reference_columns():
    y = _tone(330.0, secs=1.0)
    full = C.audio_classical(y)
    lean = ["rms", "mfcc", "chroma"]                       # any order in: canonical order out (classical.py:166-167)
    ex = AudioClassicalExtractor(features=lean, aggregations=["std"])
    assert ex.features == ["mfcc", "chroma", "rms"] and ex.feature_dim == 40 + 12 + 1
    assert np.array_equal(full[ex._columns], C.audio_classical(y, features=lean, aggregations=["std"]))
    assert AudioClassicalExtractor().feature_dim == 302
    assert AudioClassicalExtractor(aggregations=["mean"]).feature_dim == 151
    with pytest.raises(ValueError):
        AudioClassicalExtractor(features=["mfcc", "nope"])
    with pytest.raises(ValueError):
        AudioClassicalExtractor(aggregations=[])
    assert AudioClassicalExtractor()._min_samples() == 4096 and C.min_samples(SR, NFFT, HOP) == 4096


HOST_TABLES = r'''
#include <cstdio>
#include "tables.h"
int main(int argc, char** argv) {
    const int sr = std::atoi(argv[1]), n_fft = std::atoi(argv[2]);
    b2a::ClassicalTables t; const char* err = nullptr;
    if (!b2a::build_classical_tables(sr, n_fft, &t, &err)) { std::printf("ERR %s\n", err); return 0; }
    for (int i = 0; i < 7; ++i) std::printf("%d %d %d\n", t.band_start[i], t.band_cnt[i], t.band_q[i]);
    std::printf("%d %d\n", t.pip_k0, t.pip_k1);
    for (float v : t.tonnetz) std::printf("%a\n", (double)v);
    const int nb = 1 + n_fft / 2;
    for (int bank : {0, 37, 50, 99})
        for (int i = 0; i < 12 * nb; ++i) std::printf("%a\n", (double)t.chroma[(size_t)bank * 12 * nb + i]);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("sr,n_fft", [(22050, 1024), (16000, 512)])
def test_product_tables_match_the_oracle(tmp_path, sr, n_fft):
    src = tmp_path / "cls_tables.cpp"
    src.write_text(HOST_TABLES.replace("#include <cstdio>", "#include <cstdio>\n#include <cstdlib>"))
    exe = tmp_path / "cls_tables"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(CSRC), str(src), str(CSRC / "tables.cpp"), "-o", str(exe)], check=True)
    tok = subprocess.run([str(exe), str(sr), str(n_fft)], capture_output=True, text=True, check=True).stdout.split()
    assert tok[0] != "ERR", tok
    bands = C.contrast_bands(sr, n_fft)
    for i, (b, q) in enumerate(bands):
        assert [int(tok[3 * i]), int(tok[3 * i + 1]), int(tok[3 * i + 2])] == [int(b[0]), len(b), q]
    freqs = C.fft_frequencies(sr, n_fft)
    mask = np.flatnonzero((freqs >= 150.0) & (freqs < min(4000.0, sr / 2)))
    assert [int(tok[21]), int(tok[22])] == [int(mask[0]), int(mask[-1]) + 1]
    vals = np.array([float.fromhex(t) for t in tok[23:]])
    assert np.allclose(vals[:72].reshape(6, 12), C.tonnetz_matrix(), atol=1e-7)
    nb = 1 + n_fft // 2
    edges = np.linspace(-0.5, 0.5, 101)
    for n, bank in enumerate([0, 37, 50, 99]):
        got = vals[72 + n * 12 * nb: 72 + (n + 1) * 12 * nb].reshape(12, nb)
        ref = C.chroma_filterbank(sr, n_fft, tuning=float(edges[bank]))
        assert np.abs(got - ref).max() <= 2e-7, (bank, np.abs(got - ref).max())


def test_product_rejects_a_sample_rate_below_the_contrast_bands(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    src = tmp_path / "cls_tables.cpp"
    src.write_text(HOST_TABLES.replace("#include <cstdio>", "#include <cstdio>\n#include <cstdlib>"))
    exe = tmp_path / "cls_tables"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(CSRC), str(src), str(CSRC / "tables.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "8000", "512"], capture_output=True, text=True, check=True).stdout
    assert out.startswith("ERR")                          # librosa raises ParameterError for the same input
    with pytest.raises(ValueError):
        C.contrast_bands(8000, 512)

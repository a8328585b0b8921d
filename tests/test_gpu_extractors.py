"""GPU: the registered extractors end to end (WAV files -> decode -> C ABI -> features.npy)
against the oracle, including the in-process multi-device sharder on however many GPUs exist."""
import json
import logging

import numpy as np
import pytest

import audio_edge_ml_pipeline_b200 as P
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import pipeline, synth, wavio
from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
from oracle import librosa_restated as L

pytestmark = pytest.mark.gpu


def _dataset(root, sr, n_lo, n_hi, classes=3, per=5, seed=0):
    rng = np.random.default_rng(seed)
    pcm = {}
    for c in range(classes):
        d = root / f"class_{c:02d}"
        d.mkdir(parents=True)
        for i in range(per):
            x = synth.to_pcm16(synth.make_clip(rng, (c * per + i) % 5, sr, int(rng.integers(n_lo, n_hi))))
            wavio.write_wav_pcm16(d / f"c{i:02d}.wav", x, sr)
            pcm[(d.name, f"c{i:02d}.wav")] = x
    return pcm


def test_fsc22_shaped_experiment_from_yaml(tmp_path, monkeypatch):
    """config 1/5 in miniature: the reference YAML, WAV files on disk, features.npy out."""
    pcm = _dataset(tmp_path / "ds", 16000, 60000, 100000)
    (tmp_path / "ds" / "class_01" / "broken.wav").write_bytes(b"zzz")
    cfg = f"""
dataset: {tmp_path / 'ds'}
experiments:
  - name:      melspec
    extractor: audio_mel_spec
    loader:    audio_folder
    split:     ""
    extractor_params: {{duration: 5.0, n_mels: 40, sample_rate: 16000, n_fft: 512, hop_length: 160}}
"""
    (tmp_path / "c.yaml").write_text(cfg)
    monkeypatch.chdir(tmp_path)
    (fs,) = pipeline.run_config(tmp_path / "c.yaml")
    out = tmp_path / "data/processed/melspec"
    f = np.load(out / "features.npy")
    assert f.shape == (15, 40, 501) and f.dtype == np.float32
    meta = json.loads((out / "metadata.json").read_text())
    assert np.load(out / "labels.npy").tolist() == [0] * 5 + [1] * 5 + [2] * 5
    for k, m in enumerate(meta):
        ref = L.audio_mel_spec(L.pcm16_to_float(pcm[(m["class_dir"], m["filename"])]), duration=5.0)
        assert np.abs(f[k] - ref).max() <= 1e-4


def test_variable_length_clips_without_duration(tmp_path):
    """duration=None: every clip keeps its own frame count (grouped by length on the host)."""
    pcm = _dataset(tmp_path / "ds", 16000, 3000, 9000, classes=1, per=6, seed=3)
    ex = P.AudioMelSpectrogram()
    for (cd, fn), x in pcm.items():
        got = ex.extract(tmp_path / "ds" / cd / fn)
        ref = L.audio_mel_spec(L.pcm16_to_float(x))
        assert got.shape == ref.shape == (40, 1 + len(x) // 160)
        assert np.abs(got - ref).max() <= 1e-4
    with pytest.raises(ValueError):                # np.stack of ragged features, as in the reference
        ex.extract_dataset(AudioFolderLoader(tmp_path / "ds"))


def test_mfcc_and_cqt_extractors_on_files(tmp_path):
    pcm = _dataset(tmp_path / "ds", 22050, 100000, 120000, classes=2, per=3, seed=5)
    loader = AudioFolderLoader(tmp_path / "ds")
    fs = P.AudioMFCCSequence(duration=5.0).extract_dataset(loader)
    assert fs.features.shape == (6, 40, 216)
    fc = P.AudioCQT(duration=5.0).extract_dataset(loader)
    assert fc.features.shape == (6, 84, 216)
    for k, m in enumerate(fs.metadata):
        y = L.pcm16_to_float(pcm[(m["class_dir"], m["filename"])])
        assert np.abs(fs.features[k] - L.audio_mfcc_seq(y, duration=5.0)).max() <= 1e-3
        assert np.abs(fc.features[k] - L.audio_cqt(y, duration=5.0)).max() <= 1.25e-4      # tests/test_gpu_parity.py: CQT_TOL


def test_bad_cqt_config_skips_every_sample_like_the_reference(tmp_path, caplog):
    _dataset(tmp_path / "ds", 16000, 20000, 21000, classes=1, per=2)
    with caplog.at_level(logging.WARNING), pytest.raises(RuntimeError, match="No features were successfully"):
        P.AudioCQT(sample_rate=16000, n_bins=120).extract_dataset(AudioFolderLoader(tmp_path / "ds"))
    assert any("Nyquist" in r.message for r in caplog.records)


def test_sharded_over_all_visible_gpus_is_bit_identical():
    pcm = synth.make_noise_batch(257, 80000, seed=17)
    one = P.AudioMelSpectrogram(duration=5.0, devices=[0]).extract_batch(pcm)
    allg = P.AudioMelSpectrogram(duration=5.0, devices="all").extract_batch(pcm)
    assert B.device_count() >= 1 and np.array_equal(one, allg)


def test_smoke_entry_point():
    import __graft_entry__ as G
    G.smoke()

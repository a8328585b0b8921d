"""Host-side mirror of the reference's plugin interface: registry, extractor constructors,
extract / extract_dataset semantics, WAV decode, persistence layout, YAML config.  The compute
engine is replaced by tests/fake_engine.py (oracle-backed) — these tests run without a GPU."""
import inspect
import json
import logging
import sys
from pathlib import Path

import numpy as np
import pytest

import audio_edge_ml_pipeline_b200 as P
from audio_edge_ml_pipeline_b200 import extractors, pipeline, registry, synth, wavio
from audio_edge_ml_pipeline_b200.base import BaseFeatureExtractor
from fake_engine import FakeEngine
from oracle import librosa_restated as L

REF = Path("/root/reference")


@pytest.fixture()
def fake(monkeypatch):
    FakeEngine.calls = []
    monkeypatch.setattr(extractors, "_make_engine", lambda cfg, dev: FakeEngine(cfg, dev))
    return FakeEngine


def _make_dataset(root: Path, sr=16000, n=8000, classes=("bird", "axe", "rain"), per=3, seed=0, broken=()):
    rng = np.random.default_rng(seed)
    clips = {}
    for c in classes:
        (root / c).mkdir(parents=True)
        for i in range(per):
            pcm = synth.to_pcm16(rng.standard_normal(int(n * rng.uniform(0.7, 1.2))) * 0.1)
            p = root / c / f"clip_{i}.wav"
            if (c, i) in broken:
                p.write_bytes(b"not a wav file")
            else:
                wavio.write_wav_pcm16(p, pcm, sr)
            clips[(c, i)] = pcm
    return clips


# ---- registry / constructors -------------------------------------------------------------------

def test_registry_semantics_match_reference():
    assert P.list_extractors() == ["audio_classical", "audio_cqt", "audio_mel_spec", "audio_mfcc_seq"]
    assert P.get("audio_mel_spec") is P.AudioMelSpectrogram
    with pytest.raises(KeyError):
        P.get("nope")
    with pytest.raises(ValueError):
        registry.register(P.AudioMelSpectrogram)          # duplicate name

    class NoName(BaseFeatureExtractor):
        def extract(self, sample_path, **kw):
            return np.zeros(1)
    with pytest.raises(TypeError):
        registry.register(NoName)


def test_constructor_signatures_are_the_reference_ones():
    def pos(cls):
        return [(n, p.default) for n, p in inspect.signature(cls.__init__).parameters.items()
                if n != "self" and p.kind == p.POSITIONAL_OR_KEYWORD]
    assert pos(P.AudioMelSpectrogram) == [("sample_rate", 16000), ("n_mels", 40), ("n_fft", 512),
                                          ("hop_length", 160), ("duration", None)]          # deep.py:98-105
    assert pos(P.AudioMFCCSequence) == [("sample_rate", 22050), ("n_mfcc", 40), ("n_fft", 1024),
                                        ("hop_length", 512), ("duration", None)]            # deep.py:290-297
    assert pos(P.AudioCQT) == [("sample_rate", 22050), ("hop_length", 512), ("n_bins", 84),
                               ("bins_per_octave", 12), ("fmin", None), ("duration", None)]  # deep.py:219-227
    for cls, nm in ((P.AudioMelSpectrogram, "audio_mel_spec"), (P.AudioMFCCSequence, "audio_mfcc_seq"),
                    (P.AudioCQT, "audio_cqt")):
        assert (cls.name, cls.feature_type, cls.modality) == (nm, "deep", "audio")
    with pytest.raises(TypeError):                          # pipeline.py:524 splat: unknown key fails
        P.AudioMelSpectrogram(bogus=1)
    # the YAML keys of config/feature_extraction.yaml:65-70
    P.AudioMelSpectrogram(duration=5.0, n_mels=40, sample_rate=16000, n_fft=512, hop_length=160)


@pytest.mark.skipif(not REF.exists(), reason="authoring container only")
def test_constructor_signatures_against_reference_source():
    import ast
    tree = ast.parse((REF / "src/preprocessing/feature_extraction/audio/deep.py").read_text())
    want = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == "__init__":
                    args = [a.arg for a in f.args.args[1:]]
                    defs = [ast.literal_eval(d) for d in f.args.defaults]
                    want[node.name] = list(zip(args, defs))
    for cls in (P.AudioMelSpectrogram, P.AudioMFCCSequence, P.AudioCQT):
        got = [(n, p.default) for n, p in inspect.signature(cls.__init__).parameters.items()
               if n != "self" and p.kind == p.POSITIONAL_OR_KEYWORD]
        assert got == want[cls.__name__]


# ---- extract / extract_dataset -------------------------------------------------------------------

def test_extract_swallows_loader_metadata_and_matches_oracle(tmp_path, fake):
    pcm = synth.to_pcm16(np.random.default_rng(1).standard_normal(20000) * 0.1)
    wavio.write_wav_pcm16(tmp_path / "a.wav", pcm, 16000)
    ex = P.AudioMelSpectrogram(duration=1.0)
    got = ex.extract(tmp_path / "a.wav", filename="a.wav", class_dir="x", duration=1.25, sample_rate=16000,
                     n_channels=1)                        # audio_folder_loader.py:181-185 metadata
    ref = L.audio_mel_spec(L.pcm16_to_float(pcm), duration=1.0)
    assert got.shape == (40, 101) and got.dtype == np.float32 and got.flags.c_contiguous
    assert np.array_equal(got, ref)
    seg = ex.extract(tmp_path / "a.wav", start_time=0.25, end_time=0.75)
    ref = L.audio_mel_spec(L.pcm16_to_float(pcm[4000:12000]), duration=1.0)
    assert np.array_equal(seg, ref)


def test_short_clip_is_padded_to_min_samples(tmp_path, fake):
    wavio.write_wav_pcm16(tmp_path / "s.wav", np.arange(100, dtype=np.int16), 16000)
    assert P.AudioMelSpectrogram().extract(tmp_path / "s.wav").shape == (40, 1 + 512 // 160)      # n_fft
    assert P.AudioCQT(sample_rate=16000, n_bins=24).extract(tmp_path / "s.wav").shape == (24, 3)  # 2*hop


def test_extract_dataset_reproduces_reference_loop(tmp_path, fake, caplog):
    clips = _make_dataset(tmp_path / "ds", broken={("axe", 1)})
    loader = P.loaders.AudioFolderLoader(tmp_path / "ds") if hasattr(P, "loaders") else None
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    loader = AudioFolderLoader(tmp_path / "ds")
    ex = P.AudioMelSpectrogram(duration=0.5)
    with caplog.at_level(logging.WARNING):
        fs = ex.extract_dataset(loader)
    serial = BaseFeatureExtractor.extract_dataset(ex, loader)        # the reference's loop, verbatim semantics
    assert np.array_equal(fs.features, serial.features) and fs.features.shape == (8, 40, 51)
    assert np.array_equal(fs.labels, serial.labels) and fs.labels.dtype == np.int32
    assert fs.label_names == serial.label_names == ["axe", "bird", "rain"]     # sorted dirs, first-seen order
    assert fs.metadata == serial.metadata and len(fs.metadata) == 8
    assert any("Skipping" in r.message for r in caplog.records)
    assert fs.features.flags.c_contiguous and fs.features.dtype == np.float32
    # max_samples cuts by enumeration index, failed samples included (base.py:199-201)
    fs4 = ex.extract_dataset(loader, max_samples=5)
    assert fs4.n_samples == 4 and fs4.label_names == ["axe", "bird"]


def test_extract_dataset_first_class_all_broken_shifts_label_indices(tmp_path, fake):
    _make_dataset(tmp_path / "ds", classes=("a", "b"), per=2, broken={("a", 0), ("a", 1)})
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    fs = P.AudioMelSpectrogram(duration=0.5).extract_dataset(AudioFolderLoader(tmp_path / "ds"))
    assert fs.label_names == ["b"] and fs.labels.tolist() == [0, 0]


def test_extract_dataset_nothing_extracted_raises(tmp_path, fake):
    _make_dataset(tmp_path / "ds", classes=("a",), per=2, broken={("a", 0), ("a", 1)})
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    with pytest.raises(RuntimeError, match="No features were successfully extracted."):
        P.AudioMelSpectrogram().extract_dataset(AudioFolderLoader(tmp_path / "ds"))


def test_rate_mismatch_is_a_per_sample_skip(tmp_path, fake):
    (tmp_path / "ds" / "a").mkdir(parents=True)
    wavio.write_wav_pcm16(tmp_path / "ds" / "a" / "x.wav", np.zeros(4000, np.int16), 8000)
    wavio.write_wav_pcm16(tmp_path / "ds" / "a" / "y.wav", np.zeros(4000, np.int16), 16000)
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    fs = P.AudioMelSpectrogram(duration=0.25).extract_dataset(AudioFolderLoader(tmp_path / "ds"))
    assert fs.n_samples == 1 and fs.metadata[0]["filename"] == "y.wav"


def test_duration_none_goes_through_ragged_calls_per_length_bucket(tmp_path, fake):
    """duration=None (reference default): every clip keeps its own frame count."""
    clips = _make_dataset(tmp_path / "ds", classes=("a",), per=5)
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    ex = P.AudioMelSpectrogram()
    FakeEngine.calls = []
    with pytest.raises(ValueError):                    # np.stack of ragged features, as in the reference
        ex.extract_dataset(AudioFolderLoader(tmp_path / "ds"))
    buckets = {int(np.ceil(np.log2(len(p_)))) for p_ in clips.values()}      # one launch per power-of-two length bucket
    assert len(FakeEngine.calls) == len(buckets) <= 2 and sum(n for _d, n in FakeEngine.calls) == 5
    for (c, i), pcm in clips.items():
        got = ex.extract(tmp_path / "ds" / c / f"clip_{i}.wav")
        assert got.shape == (40, 1 + len(pcm) // 160)
        assert np.array_equal(got, L.audio_mel_spec(L.pcm16_to_float(pcm)))


def test_multi_device_sharding_is_contiguous_and_ordered(fake):
    pcm = synth.make_suite(10, 16000, 4000, seed=3)
    one = P.AudioMelSpectrogram(devices=[0]).extract_batch(pcm)
    FakeEngine.calls = []
    four = P.AudioMelSpectrogram(devices=[0, 1, 2, 3]).extract_batch(pcm)
    assert np.array_equal(one, four)
    assert sorted(FakeEngine.calls) == [(0, 2), (1, 3), (2, 2), (3, 3)]


def test_mfcc_and_cqt_extractors_match_oracle(tmp_path, fake):
    pcm = synth.to_pcm16(np.random.default_rng(5).standard_normal(22050) * 0.1)
    wavio.write_wav_pcm16(tmp_path / "a.wav", pcm, 22050)
    y = L.pcm16_to_float(pcm)
    assert np.array_equal(P.AudioMFCCSequence(duration=1.0).extract(tmp_path / "a.wav"),
                          L.audio_mfcc_seq(y, duration=1.0))
    assert np.array_equal(P.AudioMFCCSequence(sample_rate=22050, n_mfcc=13, n_fft=512, hop_length=160, n_mels=40)
                          .extract(tmp_path / "a.wav"), L.audio_mfcc_seq(y, 22050, 13, 512, 160, None, n_mels=40))
    assert np.array_equal(P.AudioCQT(duration=1.0).extract(tmp_path / "a.wav"), L.audio_cqt(y, duration=1.0))


# ---- WAV decode ------------------------------------------------------------------------------------

def test_wav_decode_variants(tmp_path):
    import scipy.io.wavfile as wf
    rng = np.random.default_rng(7)
    x16 = synth.to_pcm16(rng.standard_normal(3000) * 0.2)
    wavio.write_wav_pcm16(tmp_path / "m16.wav", x16, 16000)
    a, sr = wavio.decode_wav(tmp_path / "m16.wav")
    assert sr == 16000 and a.dtype == np.int16 and np.array_equal(a, x16)
    assert np.array_equal(wf.read(tmp_path / "m16.wav")[1], x16)           # our writer is a valid WAV
    st = np.stack([x16, x16[::-1]], axis=1)
    wf.write(tmp_path / "s16.wav", 16000, st)
    a, _ = wavio.decode_wav(tmp_path / "s16.wav")
    assert a.dtype == np.float32 and np.allclose(a, (st.astype(np.float32) / 32768).mean(axis=1))
    f32 = (rng.standard_normal(3000) * 0.1).astype(np.float32)
    wf.write(tmp_path / "f32.wav", 22050, f32)
    a, sr = wavio.decode_wav(tmp_path / "f32.wav")
    assert sr == 22050 and np.array_equal(a, f32)
    i32 = (rng.integers(-2**31, 2**31 - 1, 1000)).astype(np.int32)
    wf.write(tmp_path / "i32.wav", 8000, i32)
    a, _ = wavio.decode_wav(tmp_path / "i32.wav")
    assert np.allclose(a, i32 / 2147483648.0, atol=1e-7)
    a, _ = wavio.decode_wav(tmp_path / "m16.wav", offset=0.05, duration=0.1)
    assert np.array_equal(a, x16[800:2400])
    info = wavio.wav_info(tmp_path / "s16.wav")
    assert info == {"duration": 3000 / 16000, "sample_rate": 16000, "n_channels": 2}
    assert wavio.wav_info(tmp_path / "missing.wav") == {"duration": 0.0, "sample_rate": 0, "n_channels": 0}
    with pytest.raises(wavio.AudioDecodeError):
        (tmp_path / "bad.wav").write_bytes(b"garbage")
        wavio.decode_wav(tmp_path / "bad.wav")


def test_native_batch_decoder_matches_python_decoder(tmp_path, lib_built):
    from audio_edge_ml_pipeline_b200 import _lib as B
    import scipy.io.wavfile as wf
    rng = np.random.default_rng(11)
    xs = [synth.to_pcm16(rng.standard_normal(n) * 0.2) for n in (700, 16000, 23456)]
    for i, x in enumerate(xs):
        wavio.write_wav_pcm16(tmp_path / f"{i}.wav", x, 16000)
    (tmp_path / "junk.wav").write_bytes(b"RIFFxxxxjunk")
    wavio.write_wav_pcm16(tmp_path / "r8k.wav", xs[0], 8000)
    wf.write(tmp_path / "stereo.wav", 16000, np.stack([xs[1], xs[1]], axis=1))
    wf.write(tmp_path / "f32.wav", 16000, (xs[1] / 32768).astype(np.float32))
    names = ["0.wav", "1.wav", "2.wav", "junk.wav", "r8k.wav", "stereo.wav", "f32.wav", "nope.wav"]
    out = np.full((len(names), 16000), 123, np.int16)
    st = B.decode_wav_pcm16_batch([tmp_path / n for n in names], 16000, 16000, out, n_threads=3)
    assert st.tolist() == [B.DEC_OK, B.DEC_OK, B.DEC_OK, B.DEC_EFORMAT, B.DEC_ERATE, B.DEC_EUNSUPPORTED,
                           B.DEC_EUNSUPPORTED, B.DEC_EIO]
    for k in range(3):                                  # same samples as load_segment + pad_or_trim
        ref = wavio.pad_or_trim(wavio.load_segment(tmp_path / names[k], 16000, None, None, min_samples=512), 16000)
        assert np.array_equal(out[k], ref)
    assert (out[3:] == 0).all()
    st = B.decode_wav_pcm16_batch([tmp_path / "2.wav"] * 2, 16000, 16000, out, offsets=[0.5, 1.4], durations=[0.25, -1.0])
    assert st.tolist() == [0, 0]
    assert np.array_equal(out[0], wavio.pad_or_trim(wavio.load_segment(tmp_path / "2.wav", 16000, 0.5, 0.75), 16000))
    assert np.array_equal(out[1], wavio.pad_or_trim(wavio.load_segment(tmp_path / "2.wav", 16000, 1.4, None), 16000))


def _write_wav_raw(path, fmt_tag, channels, rate, bits, payload: bytes, extensible=False):
    import struct
    align = channels * bits // 8
    if extensible:
        guid = struct.pack("<H", fmt_tag) + bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, rate, rate * align, align, bits, 22, bits, 0) + guid
    else:
        fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * align, align, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" \
        + b"data" + struct.pack("<I", len(payload)) + payload
    Path(path).write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)


def test_native_general_decoder_matches_python_decoder(tmp_path, lib_built):
    """b2a_probe_wav_batch / b2a_decode_wav_batch (stereo, 8/24/32-bit PCM, float32/64, extensible headers,
    any rate) against wavio.decode_wav, which restates what librosa.load -> soundfile returns (deep.py:44-50)."""
    from audio_edge_ml_pipeline_b200 import _lib as B
    import scipy.io.wavfile as wf
    rng = np.random.default_rng(13)
    n = 5000
    x16 = synth.to_pcm16(rng.standard_normal(n) * 0.2)
    files = {}
    wavio.write_wav_pcm16(tmp_path / "m16_16k.wav", x16, 16000); files["m16_16k.wav"] = (16000, 1, 16, 1)
    wavio.write_wav_pcm16(tmp_path / "m16_44k.wav", x16, 44100); files["m16_44k.wav"] = (44100, 1, 16, 1)
    wf.write(tmp_path / "s16.wav", 22050, np.stack([x16, x16[::-1]], axis=1)); files["s16.wav"] = (22050, 2, 16, 1)
    wf.write(tmp_path / "f32.wav", 48000, (rng.standard_normal(n) * 0.1).astype(np.float32)); files["f32.wav"] = (48000, 1, 32, 3)
    wf.write(tmp_path / "f64s.wav", 8000, rng.standard_normal((n, 2)) * 0.1); files["f64s.wav"] = (8000, 2, 64, 3)
    wf.write(tmp_path / "i32.wav", 16000, rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)); files["i32.wav"] = (16000, 1, 32, 1)
    wf.write(tmp_path / "u8.wav", 16000, rng.integers(0, 256, n).astype(np.uint8)); files["u8.wav"] = (16000, 1, 8, 1)
    p24 = rng.integers(-2**23, 2**23 - 1, (n, 3))
    _write_wav_raw(tmp_path / "i24x3.wav", 1, 3, 32000, 24,
                   b"".join(int(v).to_bytes(3, "little", signed=True) for v in p24.reshape(-1)), extensible=True)
    files["i24x3.wav"] = (32000, 3, 24, 1)
    (tmp_path / "junk.wav").write_bytes(b"RIFFxxxxjunk")
    _write_wav_raw(tmp_path / "adpcm.wav", 2, 1, 16000, 8, b"\0" * 64)
    names = list(files) + ["junk.wav", "adpcm.wav", "missing.wav"]
    info = B.probe_wav_batch([tmp_path / k for k in names], n_threads=3)
    for i, k in enumerate(files):
        assert (info["rate"][i], info["channels"][i], info["bits"][i], info["format_tag"][i]) == files[k], k
        assert info["status"][i] == B.DEC_OK and info["n_frames"][i] == n
    assert info["status"][len(files):].tolist() == [B.DEC_EFORMAT, B.DEC_EUNSUPPORTED, B.DEC_EIO]
    # float32 output: every format, channel mean, zero-filled tail
    out = np.full((len(names), n + 100), 7.0, np.float32)
    rate, n_out, status = B.decode_wav_batch([tmp_path / k for k in names], n + 100, out, n_threads=2)
    assert status.tolist() == [0] * len(files) + [B.DEC_EFORMAT, B.DEC_EUNSUPPORTED, B.DEC_EIO]
    for i, k in enumerate(files):
        ref, sr = wavio.decode_wav(tmp_path / k)
        ref = L.pcm16_to_float(ref) if ref.dtype == np.int16 else ref
        assert rate[i] == sr == files[k][0] and n_out[i] == n
        assert np.array_equal(out[i, :n], ref), k
        assert not out[i, n:].any()
    assert not out[len(files):].any()
    # int16 output: mono PCM16 only; segments in native frames; truncation to max_frames
    o16 = np.zeros((3, 2000), np.int16)
    rate, n_out, status = B.decode_wav_batch([tmp_path / "m16_44k.wav", tmp_path / "s16.wav", tmp_path / "m16_16k.wav"],
                                             2000, o16, offsets=[0.01, 0.0, 0.2], durations=[0.02, -1.0, -1.0])
    assert status.tolist() == [0, B.DEC_EUNSUPPORTED, 0] and rate.tolist() == [44100, 22050, 16000]
    assert n_out.tolist() == [882, 0, 1800]
    assert np.array_equal(o16[0, :882], x16[441:441 + 882]) and not o16[0, 882:].any()
    assert np.array_equal(o16[2, :1800], x16[3200:5000])


def test_native_and_python_decode_paths_give_the_same_dataset(tmp_path, fake, monkeypatch):
    _make_dataset(tmp_path / "ds", n=8000, broken={("axe", 1)})
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    loader = AudioFolderLoader(tmp_path / "ds")
    a = P.AudioMelSpectrogram(duration=0.5).extract_dataset(loader)
    monkeypatch.setattr(extractors, "NATIVE_DECODE", False)
    b = P.AudioMelSpectrogram(duration=0.5).extract_dataset(loader)
    assert np.array_equal(a.features, b.features) and np.array_equal(a.labels, b.labels)
    assert a.metadata == b.metadata and a.label_names == b.label_names


# ---- persistence + config ---------------------------------------------------------------------------

def test_save_layout_and_roundtrip(tmp_path, fake):
    _make_dataset(tmp_path / "ds")
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    fs = P.FeaturePipeline(AudioFolderLoader(tmp_path / "ds"), P.AudioMelSpectrogram(duration=0.5)).run()
    P.FeaturePipeline.save(fs, tmp_path / "out")
    assert sorted(p.name for p in (tmp_path / "out").iterdir()) == ["features.npy", "info.json", "label_names.json",
                                                                   "labels.npy", "metadata.json"]
    f = np.load(tmp_path / "out" / "features.npy")
    assert f.shape == (9, 40, 51) and f.dtype == np.float32 and f.flags.c_contiguous
    assert np.load(tmp_path / "out" / "labels.npy").dtype == np.int32
    info = json.loads((tmp_path / "out" / "info.json").read_text())
    assert info == {"feature_type": "deep", "modality": "audio", "n_samples": 9, "feature_shape": [40, 51],
                    "n_classes": 3, "is_supervised": True}
    back = P.FeaturePipeline.load(tmp_path / "out")
    assert np.array_equal(back.features, fs.features) and back.label_names == fs.label_names
    with pytest.raises(FileNotFoundError):
        P.FeaturePipeline.load(tmp_path / "nowhere")


def test_features_are_written_in_place_when_the_output_is_known(tmp_path, fake, lib_built):
    """run(output_dir=...) + save: features.npy is the memory-mapped array the run filled (no second copy), byte for
    byte what np.save writes; a skipped sample makes the run fall back to the ordinary array + np.save."""
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    _make_dataset(tmp_path / "ds", n=8000)
    loader = AudioFolderLoader(tmp_path / "ds")
    ref = P.FeaturePipeline(loader, P.AudioMelSpectrogram(duration=0.5)).run()
    P.FeaturePipeline.save(ref, tmp_path / "plain")
    fs = P.FeaturePipeline(loader, P.AudioMelSpectrogram(duration=0.5)).run(output_dir=tmp_path / "out")
    assert isinstance(fs.features, np.memmap) and fs.features_file == tmp_path / "out" / "features.npy"
    P.FeaturePipeline.save(fs, tmp_path / "out")
    assert (tmp_path / "out" / "features.npy").read_bytes() == (tmp_path / "plain" / "features.npy").read_bytes()
    assert np.array_equal(P.FeaturePipeline.load(tmp_path / "out").features, ref.features)
    # saving somewhere else still works (plain np.save of the mapped rows)
    P.FeaturePipeline.save(fs, tmp_path / "elsewhere")
    assert np.array_equal(np.load(tmp_path / "elsewhere" / "features.npy"), ref.features)
    # a broken file: rows are compacted in memory, the half-written file is dropped and save() writes the real one
    _make_dataset(tmp_path / "ds2", n=8000, broken={("axe", 1)})
    loader2 = AudioFolderLoader(tmp_path / "ds2")
    fs2 = P.FeaturePipeline(loader2, P.AudioMelSpectrogram(duration=0.5)).run(output_dir=tmp_path / "out2")
    assert not isinstance(fs2.features, np.memmap) and not (tmp_path / "out2" / "features.npy").exists()
    P.FeaturePipeline.save(fs2, tmp_path / "out2")
    assert np.load(tmp_path / "out2" / "features.npy").shape[0] == len(loader2) - 1


def test_output_staging_path_gives_the_same_dataset(tmp_path, fake, lib_built, monkeypatch):
    """B2A_OUT_STAGING=1: rows travel through staging buffers and helper-thread copies; windows of 4 clips so that both
    buffers are used and reused, one broken file so that the compaction has to wait for the copies."""
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    _make_dataset(tmp_path / "ds", n=8000, per=5, broken={("axe", 2)})
    loader = AudioFolderLoader(tmp_path / "ds")
    plain = P.AudioMelSpectrogram(duration=0.5).extract_dataset(loader)
    monkeypatch.setattr(extractors, "OUT_STAGING", True)
    monkeypatch.setattr(extractors, "HOST_BATCH_CLIPS", 4)
    ex = P.AudioMelSpectrogram(duration=0.5)
    staged = ex.extract_dataset(loader)
    ex.close()
    assert np.array_equal(plain.features, staged.features) and np.array_equal(plain.labels, staged.labels)
    assert plain.metadata == staged.metadata and len(staged.features) == len(loader) - 1


def test_loader_metadata_same_through_native_and_python_probe(tmp_path, lib_built, monkeypatch):
    from audio_edge_ml_pipeline_b200 import loaders
    _make_dataset(tmp_path / "ds", n=8000, broken={("axe", 1)})
    a = list(loaders.AudioFolderLoader(tmp_path / "ds"))
    monkeypatch.setattr(loaders, "_probe_all", lambda paths: [loaders.wav_info(p) for p in paths])
    b = list(loaders.AudioFolderLoader(tmp_path / "ds"))
    assert a == b and any(m["sample_rate"] == 0 for _p, _l, m in a)          # the broken file: zeros either way


def test_reference_yaml_runs_unchanged(tmp_path, fake, monkeypatch):
    """config/feature_extraction.yaml:60-70, copied verbatim (keys + extractor_params)."""
    cfg = """
dataset: DS
experiments:
  - name:           fsc22_device_augmented_melspec_train
    extractor:      audio_mel_spec
    loader:         audio_folder
    dataset:        DS
    split:          ""
    unknown_key:    ignored
    extractor_params:
      duration:    5.0
      n_mels:      40
      sample_rate: 16000
      n_fft:       512
      hop_length:  160
""".replace("DS", str(tmp_path / "ds"))
    _make_dataset(tmp_path / "ds", n=80000, classes=("a", "b"), per=2)
    (tmp_path / "c.yaml").write_text(cfg)
    monkeypatch.chdir(tmp_path)
    (fs,) = pipeline.run_config(tmp_path / "c.yaml")
    out = tmp_path / "data/processed/fsc22_device_augmented_melspec_train"
    assert np.load(out / "features.npy").shape == (4, 40, 501)
    assert (out / "config.yaml").read_text() == cfg
    if REF.exists():
        import yaml
        ref_cfg = yaml.safe_load((REF / "config/feature_extraction.yaml").read_text())
        exps = pipeline.resolve_experiments(ref_cfg)
        assert exps[0]["extractor"] == "audio_mel_spec" and exps[0]["loader"] == "audio_folder"
        assert exps[0]["extractor_params"] == {"duration": 5.0, "n_mels": 40, "sample_rate": 16000,
                                               "n_fft": 512, "hop_length": 160}
        P.get(exps[0]["extractor"])(**exps[0]["extractor_params"])

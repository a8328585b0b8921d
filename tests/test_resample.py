"""Rational resampler: what librosa.load does with soxr_hq when a file's rate differs
from the extractor's sample_rate (deep.py:44-50).  libsoxr is absent offline, so parity against its
exact output is unpinned; these tests pin (i) the library's table against the oracle's closed form,
(ii) the oracle against the properties the soxr_hq specification promises, (iii) the CUDA kernel
against the oracle, (iv) the host plumbing.  CPU tests use the oracle only as the checker."""
import numpy as np
import pytest

from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth, wavio
from oracle import librosa_restated as L

RATIOS = [(44100, 16000), (48000, 16000), (22050, 16000), (8000, 16000), (44100, 22050), (32000, 22050), (11025, 22050)]


@pytest.fixture(scope="module")
def lib_built():
    from audio_edge_ml_pipeline_b200.build import build_lib
    build_lib()
    return B.load_library()


@pytest.mark.parametrize("orig,target", RATIOS)
def test_library_table_is_the_oracle_closed_form(lib_built, orig, target):
    up, down, half, poly = B.resampler_design(orig, target)
    u2, d2, h2, g = L.resampler_prototype(orig, target)
    assert (up, down, half) == (u2, d2, h2)
    K = poly.shape[1]
    assert K % 4 == 0 and up * K >= len(g)
    ref = np.zeros(up * K)
    ref[:len(g)] = g * up
    assert np.array_equal(poly, ref.reshape(K, up).T.astype(np.float32))      # bit-exact float32


def test_two_to_one_is_the_cqt_decimator(lib_built):
    up, down, half, poly = B.resampler_design(2, 1)
    assert (up, down, half) == (1, 2, 191)
    assert np.array_equal(poly[0, :383], L.halfband_taps().astype(np.float32)) and not poly[0, 383:].any()
    y = np.random.default_rng(0).standard_normal(5000).astype(np.float32)
    a = L.resample_restated(y, 2, 1).astype(np.float64) * np.sqrt(2.0)
    assert np.abs(a - L.decimate2(y)).max() <= 5e-7                            # same filter, same alignment


def test_bad_ratio_is_rejected(lib_built):
    with pytest.raises(B.B2AError):
        B.resampler_design(44100, 16001)       # up = 16001 > 4096
    with pytest.raises(B.B2AError):
        B.resampler_design(0, 16000)


@pytest.mark.parametrize("orig,target", [(44100, 16000), (22050, 16000), (8000, 16000)])
def test_oracle_meets_the_soxr_hq_specification(orig, target):
    n = orig                                    # one second
    t = np.arange(n) / orig
    nyq = min(orig, target) / 2
    out_len = int(np.ceil(n * target / orig))
    tt = np.arange(out_len) / target
    edge = 2000
    # pass band (up to 0.913 x the lower Nyquist): amplitude and phase preserved
    for f0 in (100.0, 0.5 * nyq, 0.9 * nyq):
        z = L.resample_restated((0.5 * np.sin(2 * np.pi * f0 * t)).astype(np.float32), orig, target)
        assert len(z) == out_len and z.dtype == np.float32
        assert np.abs(z[edge:-edge] - 0.5 * np.sin(2 * np.pi * f0 * tt[edge:-edge])).max() <= 2e-6
    # stop band (from the lower Nyquist up): gone, >= 120 dB
    if orig > target:
        for f0 in (1.001 * nyq, 1.3 * nyq, 0.45 * orig):
            z = L.resample_restated((0.5 * np.sin(2 * np.pi * f0 * t)).astype(np.float32), orig, target)
            assert 20 * np.log10(np.abs(z[edge:-edge]).max() / 0.5 + 1e-30) <= -120.0
    # DC gain 1, zero-extended edges, same rate = copy
    z = L.resample_restated(np.ones(n, np.float32), orig, target)
    assert np.abs(z[edge:-edge] - 1.0).max() <= 1e-6
    y = np.arange(7, dtype=np.float32)
    assert np.array_equal(L.resample_restated(y, target, target), y)


def test_load_segment_resamples_like_librosa_load(tmp_path, monkeypatch):
    """Host plumbing on a box without a GPU: the resampler hook is replaced by the oracle (the way
    tests/fake_engine.py replaces the engine); offset/duration are applied at the native rate
    first, then the segment is resampled (librosa.load order)."""
    calls = []

    def fake(audio, orig_sr, target_sr, device=0):
        calls.append((len(audio), orig_sr, target_sr))
        y = L.pcm16_to_float(audio) if audio.dtype == np.int16 else audio
        return L.resample_restated(y, orig_sr, target_sr)

    monkeypatch.setattr(wavio, "resample_audio", fake)
    pcm = (np.sin(2 * np.pi * 440 * np.arange(44100) / 44100) * 12000).astype(np.int16)
    wavio.write_wav_pcm16(tmp_path / "a.wav", pcm, 44100)
    y = wavio.load_segment(tmp_path / "a.wav", 16000, None, None)
    assert y.dtype == np.float32 and len(y) == 16000 and calls == [(44100, 44100, 16000)]
    y2 = wavio.load_segment(tmp_path / "a.wav", 16000, 0.25, 0.75)
    assert len(y2) == 8000 and calls[-1] == (22050, 44100, 16000)
    ref = L.resample_restated(L.pcm16_to_float(pcm[11025:33075]), 44100, 16000)
    assert np.array_equal(y2, ref)
    same = wavio.load_segment(tmp_path / "a.wav", 44100, None, None)             # no resampling needed
    assert same.dtype == np.int16 and len(calls) == 2


def test_resampler_without_a_gpu_fails_loudly(lib_built):
    if B.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(B.B2AError) as ei:
        B.Resampler(44100, 16000)
    assert ei.value.code == -4                  # B2A_ENODEVICE: no CPU path


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("orig,target", RATIOS)
@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_gpu_resampler_matches_oracle(orig, target, dtype):
    rng = np.random.default_rng(orig + target)
    n = int(1.3 * orig) + 17
    t = np.arange(n) / orig
    y = 0.3 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 0.3 * min(orig, target) * t)
    x = np.clip(np.round(y * 32768), -32768, 32767).astype(np.int16) if dtype == np.int16 else y.astype(np.float32)
    with B.Resampler(orig, target) as r:
        got = r.run_host(x)
        assert len(got) == r.out_len(n) == int(np.ceil(n * target / orig))
        tiny = r.run_host(x[:1])                                                  # shorter than the filter
        assert len(tiny) == int(np.ceil(target / orig))
    ref = L.resample_restated(L.pcm16_to_float(x) if dtype == np.int16 else x, orig, target)
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6, float(np.abs(got - ref).max())       # unit-scale signal, fp32 sums
    ref1 = L.resample_restated(L.pcm16_to_float(x[:1]) if dtype == np.int16 else x[:1], orig, target)
    assert np.abs(tiny - ref1).max() <= 1e-6


@pytest.mark.gpu
def test_gpu_mixed_rate_dataset_goes_through_the_resampler(tmp_path):
    """A class folder with 44.1 kHz, 8 kHz and native 16 kHz files: the reference loads all three
    (librosa.load resamples); so does this package, and each row equals oracle(resample -> log-mel)."""
    import audio_edge_ml_pipeline_b200 as P
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    rng = np.random.default_rng(5)
    d = tmp_path / "ds" / "birds"
    d.mkdir(parents=True)
    clips = {}
    for name, sr in (("a.wav", 44100), ("b.wav", 8000), ("c.wav", 16000)):
        n = int(1.5 * sr)
        pcm = np.clip(np.round((0.2 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 900 * np.arange(n) / sr)) * 32768),
                      -32768, 32767).astype(np.int16)
        wavio.write_wav_pcm16(d / name, pcm, sr)
        clips[name] = (pcm, sr)
    fs = P.AudioMelSpectrogram(duration=1.0).extract_dataset(AudioFolderLoader(tmp_path / "ds"))
    assert fs.n_samples == 3 and fs.features.shape == (3, 40, 101)
    for row, meta in zip(fs.features, fs.metadata):
        pcm, sr = clips[meta["filename"]]
        y = L.pcm16_to_float(pcm)
        if sr != 16000:
            y = L.resample_restated(y, sr, 16000)
        ref = L.audio_mel_spec(y, 16000, 40, 512, 160, 1.0)
        assert np.abs(row - ref).max() <= 1e-4, (meta["filename"], float(np.abs(row - ref).max()))


@pytest.mark.gpu
def test_shared_resampler_is_safe_from_many_threads():
    """extract_dataset decodes on a thread pool and every worker whose file is at another rate goes
    through the process-wide resampler of that (orig, target, device): concurrent run_host calls on one
    handle must each return what a serial call returns (the library serialises them)."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(11)
    clips = [np.clip(np.round(0.3 * rng.standard_normal(20000 + 3001 * k) * 32768), -32768, 32767).astype(np.int16)
             for k in range(12)]
    serial = [B.resample(c, 44100, 16000) for c in clips]
    for _ in range(3):
        with ThreadPoolExecutor(max_workers=12) as pool:
            got = list(pool.map(lambda c: B.resample(c, 44100, 16000), clips))
        for g, s in zip(got, serial):
            assert g.shape == s.shape and np.array_equal(g, s)
    assert len([k for k in B._resamplers if k[:2] == (44100, 16000)]) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("orig,dtype", [(44100, np.int16), (22050, np.float32), (48000, np.int16), (8000, np.int16)])
def test_batched_resample_then_extract_equals_per_clip_path(orig, dtype):
    """b2a_run_host_resampled: ragged clips at the file rate -> device resampler (many clips per launch,
    pad / trim to n_samples) -> log-mel, against the single-clip resampler + extract_batch, bit for bit,
    and against the oracle within the mel tolerance."""
    import audio_edge_ml_pipeline_b200 as P
    rng = np.random.default_rng(orig)
    n_t = 16000                                             # 1 s at 16 kHz
    lens = [int(f * orig) for f in (0.31, 1.0, 1.7, 0.05, 1.02)] + [1]
    stride = max(lens)
    raw = np.zeros((len(lens), stride), dtype)
    for i, n in enumerate(lens):
        y = 0.3 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 700 * np.arange(n) / orig)
        raw[i, :n] = np.clip(np.round(y * 32768), -32768, 32767).astype(np.int16) if dtype == np.int16 else y
    mel = P.AudioMelSpectrogram(duration=1.0)
    eng = mel._engine(n_t, np.float32, 0)
    got = eng.run_host_resampled(B.get_resampler(orig, 16000, 0), raw, np.array(lens, np.int32))
    assert got.shape == (len(lens), 40, 101)
    for i, n in enumerate(lens):
        y = B.resample(raw[i, :n], orig, 16000)              # one clip per launch
        clip = np.zeros(n_t, np.float32)
        clip[:min(n_t, len(y))] = y[:n_t]
        one = mel.extract_batch(clip[None])[0]
        assert np.array_equal(got[i], one), (i, float(np.abs(got[i] - one).max()))
        yr = L.resample_restated(L.pcm16_to_float(raw[i, :n]) if dtype == np.int16 else raw[i, :n], orig, 16000)
        ref = L.audio_mel_spec(yr, 16000, 40, 512, 160, 1.0)
        assert np.abs(got[i] - ref).max() <= 1e-4
    mel.close()


@pytest.mark.gpu
def test_mixed_format_dataset_through_the_native_front_end(tmp_path, caplog):
    """One folder with mono PCM16 at the target rate, 44.1 kHz PCM16, stereo 22.05 kHz, float32 at the target
    rate, 24-bit at 48 kHz and a broken file: every decodable file yields the features the per-file
    extract() path yields (librosa.load semantics: channel mean, resample, pad / trim), in loader order,
    and the broken one is skipped with a warning — one sample, not its whole batch."""
    import logging
    import scipy.io.wavfile as wf
    import audio_edge_ml_pipeline_b200 as P
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
    rng = np.random.default_rng(17)
    d = tmp_path / "ds" / "c0"
    d.mkdir(parents=True)

    def sig(n, sr):
        return 0.2 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 600 * np.arange(n) / sr)
    for k in range(5):
        wavio.write_wav_pcm16(d / f"a{k}_m16.wav", synth.to_pcm16(sig(14000 + 900 * k, 16000)), 16000)
        wavio.write_wav_pcm16(d / f"b{k}_44k.wav", synth.to_pcm16(sig(40000 + 2500 * k, 44100)), 44100)
        wf.write(d / f"c{k}_stereo22k.wav", 22050, np.stack([synth.to_pcm16(sig(25000, 22050)),
                                                            synth.to_pcm16(sig(25000, 22050))], axis=1))
        wf.write(d / f"d{k}_f32.wav", 16000, sig(17000, 16000).astype(np.float32))
        wf.write(d / f"e{k}_i32_48k.wav", 48000, (sig(50000, 48000) * 2**30).astype(np.int32))
    (d / "b9_broken.wav").write_bytes(b"RIFF\x00\x00\x00\x00WAVEfmt ")
    ex = P.AudioMelSpectrogram(duration=1.0)
    loader = AudioFolderLoader(tmp_path / "ds")
    with caplog.at_level(logging.WARNING):
        fs = ex.extract_dataset(loader)
    assert fs.n_samples == 25 and fs.features.shape == (25, 40, 101)
    assert sum("Skipping" in r.message for r in caplog.records) == 1
    names = [m["filename"] for m in fs.metadata]
    assert names == sorted(names) and "b9_broken.wav" not in names
    for row, meta in zip(fs.features, fs.metadata):
        one = ex.extract(d / meta["filename"])                # Python decode + single-clip resampler
        assert np.abs(row - one).max() <= 2e-6, (meta["filename"], float(np.abs(row - one).max()))
    ex.close()

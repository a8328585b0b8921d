"""BASELINE configs 1-3 at full size (27 x 75 = 2025 synthetic clips) against the oracle — the
committed tests use subsets so that the suite stays fast; this is the one-off full run whose
result DESIGN.md quotes.   python tools/full_parity.py [n_clips]"""
import json, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2025


def _mel(c): return L.audio_mel_spec(L.pcm16_to_float(c), duration=5.0)
def _mfcc(c):
    """z-scored MFCC with each row's standard deviation (MFCC units) appended as a last column."""
    z = L.audio_mfcc_seq(L.pcm16_to_float(c), 16000, 13, 512, 160, 5.0, n_mels=40)
    y = L.prepare_audio(L.pcm16_to_float(c), 16000, 5.0, min_samples=512)
    sd = L.mfcc(y, sr=16000, n_mfcc=13, n_fft=512, hop_length=160, n_mels=40).std(axis=1)
    return np.concatenate([z, sd[:, None].astype(np.float32)], axis=1)
def _cqt(c): return L.audio_cqt(L.pcm16_to_float(c), duration=5.0)


def run(name, kind, sr, n, fn, n_clips, **kw):
    pcm = synth.make_suite(n_clips, sr, n, seed=1234)
    cfg = B.default_config(kind); cfg.n_samples = n
    for k, v in kw.items(): setattr(cfg, k, v)
    t0 = time.time()
    with B.Engine(cfg, 0) as e:
        got = e.run_host(pcm)
    t1 = time.time()
    with ProcessPoolExecutor() as ex:
        ref = np.stack(list(ex.map(fn, pcm, chunksize=8)))
    extra = {}
    if ref.shape[-1] == got.shape[-1] + 1:        # mfcc: per-row sd rides along (tests/test_gpu_parity.py tolerance)
        sd, ref = ref[..., -1], ref[..., :-1]
        row = np.abs(got - ref).max(axis=2)       # (clips, rows) z-score error
        extra = dict(z_err_by_min_sd={str(t): float(np.where(sd >= t, row, 0).max()) for t in (0.5, 1, 2, 4, 8, 16)},
                     mfcc_err_by_max_sd={str(t): float(np.where(sd < t, row * sd, 0).max()) for t in (0.5, 1, 2, 4, 8, 16)},
                     worst_over_tolerance=float((row / np.maximum(1.0, 1.0 / np.maximum(sd, 1e-12))).max() / 1e-3))
    err = np.abs(got - ref).reshape(n_clips, -1).max(axis=1)
    fam = [float(err[f::synth.N_FAMILIES].max()) for f in range(synth.N_FAMILIES)]
    print(json.dumps(dict(config=name, clips=n_clips, shape=list(got.shape[1:]), max_abs=float(err.max()),
                          p99=float(np.percentile(err, 99)), per_family_max=fam, gpu_s=round(t1 - t0, 3),
                          oracle_s=round(time.time() - t1, 1), **extra)), flush=True)


run("1 audio_mel_spec", B.KIND_MEL, 16000, 80000, _mel, N)
run("2 audio_mfcc_seq", B.KIND_MFCC, 16000, 80000, _mfcc, N, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=13)
run("3 audio_cqt", B.KIND_CQT, 22050, 110250, _cqt, min(N, 405))


# the reference-default shapes (n_fft 1024 kernel, round 2): audio_mfcc_seq defaults and audio_mel_spec at 1024 / 256 / 64
def _mfcc_def(c):
    y = L.pcm16_to_float(c)
    z = L.audio_mfcc_seq(y, 22050, 40, 1024, 512, 5.0, n_mels=128)
    yy = L.prepare_audio(y, 22050, 5.0, min_samples=1024)
    sd = L.mfcc(yy, sr=22050, n_mfcc=40, n_fft=1024, hop_length=512, n_mels=128).std(axis=1)
    return np.concatenate([z, sd[:, None].astype(np.float32)], axis=1)
def _mel1024(c): return L.audio_mel_spec(L.pcm16_to_float(c), 16000, 64, 1024, 256, 5.0)


if len(sys.argv) > 2 and sys.argv[2] == "defaults":
    run("audio_mfcc_seq reference defaults (22050/1024/512/128->40)", B.KIND_MFCC, 22050, 110250, _mfcc_def, min(N, 810))
    run("audio_mel_spec 16000/1024/256/64", B.KIND_MEL, 16000, 80000, _mel1024, min(N, 810), n_fft=1024, hop_length=256, n_mels=64)

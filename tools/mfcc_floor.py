"""Where the mfcc z-score error comes from: the headline mfcc shape over the
2025-clip suite (or a subset) through whichever library B2A_LIBRARY names, per-family worst error in
z units and in MFCC units, and the same for the rows with the smallest standard deviation.

    B2A_LIBRARY=.../build/libb2a_<tag>.so python tools/mfcc_floor.py [n_clips] [tag]
"""
import json, os, sys
from concurrent.futures import ProcessPoolExecutor
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2025
TAG = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("B2A_LIBRARY", "product")


def _ref(c):
    z = L.audio_mfcc_seq(L.pcm16_to_float(c), 16000, 13, 512, 160, 5.0, n_mels=40)
    y = L.prepare_audio(L.pcm16_to_float(c), 16000, 5.0, min_samples=512)
    sd = L.mfcc(y, sr=16000, n_mfcc=13, n_fft=512, hop_length=160, n_mels=40).std(axis=1)
    return np.concatenate([z, sd[:, None].astype(np.float32)], axis=1)


def main():
    pcm = synth.make_suite(N, 16000, 80000, seed=1234)
    cache = f"/tmp/mfcc_floor_ref_{N}.npy"
    if os.path.exists(cache):
        ref = np.load(cache)
    else:
        with ProcessPoolExecutor() as ex:
            ref = np.stack(list(ex.map(_ref, pcm, chunksize=8)))
        np.save(cache, ref)
    sd, ref = ref[..., -1], ref[..., :-1]
    cfg = B.default_config(B.KIND_MFCC)
    cfg.n_samples, cfg.sample_rate, cfg.n_fft, cfg.hop_length, cfg.n_mels, cfg.n_mfcc = 80000, 16000, 512, 160, 40, 13
    with B.Engine(cfg, 0) as e:
        got = e.run_host(pcm)
    row = np.abs(got - ref).max(axis=2)                    # (clips, 13) z error
    silent = ~pcm.any(axis=1)
    row[silent] = 0
    fam = [float(row[f::synth.N_FAMILIES].max()) for f in range(synth.N_FAMILIES)]
    i, k = np.unravel_index(np.argmax(row), row.shape)
    print(json.dumps(dict(tag=TAG, clips=N, max_z=float(row.max()), p99_z=float(np.percentile(row.max(axis=1), 99)),
                          per_family_max_z=fam, worst=dict(clip=int(i), coef=int(k), sd=float(sd[i, k]),
                                                           mfcc_err=float(row[i, k] * sd[i, k])),
                          max_z_sd_lt_1=float(np.where(sd < 1, row, 0).max()),
                          max_z_sd_ge_1=float(np.where(sd >= 1, row, 0).max()),
                          n_rows_over_1e3=int((row > 1e-3).sum()))), flush=True)


if __name__ == "__main__":
    main()

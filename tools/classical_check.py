"""audio_classical on the GPU against oracle/classical_restated.py, group by group.
    python tools/classical_check.py [n_clips] [sr n_fft hop secs]
Prints one JSON line: per feature group the worst error relative to the group's scale, the clips whose tuning
estimate differs (the arg-max of a histogram: the one discontinuous step), and the device rate."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from audio_edge_ml_pipeline_b200 import _lib as B, synth
from oracle import classical_restated as C, librosa_restated as L

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 21
sr, n_fft, hop, secs = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])) if len(sys.argv) > 5 else (22050, 1024, 512, 5.0)
n = int(sr * secs)
pcm = synth.make_suite(n_clips, sr, n, seed=4321)
cfg = B.default_config(B.KIND_CLASSICAL)
cfg.n_samples, cfg.sample_rate, cfg.n_fft, cfg.hop_length = n, sr, n_fft, hop
GROUPS = [("mfcc", 40), ("delta_mfcc", 40), ("delta2_mfcc", 40), ("spectral_centroid", 1), ("spectral_rolloff", 1),
          ("spectral_bandwidth", 1), ("spectral_contrast", 7), ("spectral_flatness", 1), ("chroma", 12), ("zcr", 1),
          ("rms", 1), ("tonnetz", 6)]
# the oracle first, on every host core (forked workers: before this process touches CUDA)
def _ref(c):
    y = L.pcm16_to_float(c)
    return C.audio_classical(y, sr=sr, n_fft=n_fft, hop=hop), C.estimate_tuning(np.abs(L.stft(y, n_fft=n_fft, hop_length=hop)) ** 2.0, sr, n_fft)
import os
from concurrent.futures import ProcessPoolExecutor
with ProcessPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as pool:
    _both = list(pool.map(_ref, list(pcm), chunksize=2))
ref = np.stack([b[0] for b in _both])
rtun = np.array([b[1] for b in _both])
with B.Engine(cfg, 0) as e:
    t0 = time.time(); got = e.run_host(pcm)[:, :, 0]; dt = time.time() - t0
    t0 = time.time(); got2 = e.run_host(pcm)[:, :, 0]; dt2 = time.time() - t0
    tun = e.classical_tunings(n_clips)
assert np.array_equal(got, got2), "not deterministic"
same = np.abs(tun - rtun) < 1e-6
out = {"clips": n_clips, "cfg": [sr, n_fft, hop, secs], "finite": bool(np.isfinite(got).all()), "tuning_mismatch": np.flatnonzero(~same).tolist(),
       "tunings_differing": [[int(i), round(float(tun[i]), 2), round(float(rtun[i]), 2)] for i in np.flatnonzero(~same)], "first_call_s": dt, "second_call_s": dt2, "groups": {}}
pos = 0
for name, dim in GROUPS:
    for agg in ("mean", "std"):
        g, r = got[:, pos:pos + dim], ref[:, pos:pos + dim]
        scale = np.maximum(np.abs(r).max(axis=1, keepdims=True), 1e-3 if name not in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth") else 1.0)
        err = np.abs(g - r) / scale
        rows = same if name in ("chroma", "tonnetz") else np.ones(n_clips, bool)
        out["groups"][f"{name}.{agg}"] = {"max_rel": float(err[rows].max()) if rows.any() else None, "worst_clip": int(err.max(axis=1).argmax()),
                                          "max_abs": float(np.abs(g - r)[rows].max()) if rows.any() else None}
        pos += dim
print(json.dumps(out))

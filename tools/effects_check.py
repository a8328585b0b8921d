"""time_stretch / pitch_shift on the GPU against oracle/effects_restated.py.
    python tools/effects_check.py [n_clips]
One JSON line per effect: worst and rms error relative to the clip's peak, per clip."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from audio_edge_ml_pipeline_b200 import _lib as B, synth
from oracle import effects_restated as E, librosa_restated as L

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sr = 16000
rng = np.random.default_rng(11)
lens = [80000 if k % 3 else int(rng.integers(6000, 50000)) for k in range(n_clips)]
clips = [L.pcm16_to_float(synth.make_suite(1, sr, n, seed=500 + k)[0]) for k, n in enumerate(lens)]
rates = rng.uniform(0.85, 1.15, n_clips); rates[0] = 1.0
steps = rng.uniform(-3.0, 3.0, n_clips)

def _st(a): return E.time_stretch(a[0], a[1])
def _ps(a): return E.pitch_shift(a[0], sr, a[1])
import os
from concurrent.futures import ProcessPoolExecutor
with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
    ref_st = list(pool.map(_st, list(zip(clips, rates))))
    ref_ps = list(pool.map(_ps, list(zip(clips, steps))))

def report(name, got, ref, extra):
    rows = []
    for g, r in zip(got, ref):
        assert g.shape == r.shape, (g.shape, r.shape)
        peak = max(float(np.abs(r).max()), 1e-9)
        d = g.astype(np.float64) - r
        rows.append([float(np.abs(d).max() / peak), float(np.sqrt(np.mean(d * d)) / peak)])
    rows = np.array(rows)
    print(json.dumps({"effect": name, "clips": len(got), "worst_max_rel": float(rows[:, 0].max()), "median_max_rel": float(np.median(rows[:, 0])),
                      "worst_rms_rel": float(rows[:, 1].max()), "per_clip_max_rel": [round(x, 6) for x in rows[:, 0]], **extra}))

t0 = time.time(); got = B.time_stretch_rows(clips, rates); t1 = time.time(); got = B.time_stretch_rows(clips, rates); t2 = time.time()
report("time_stretch", got, ref_st, {"first_call_s": t1 - t0, "second_call_s": t2 - t1, "finite": bool(all(np.isfinite(g).all() for g in got))})
t0 = time.time(); got = B.pitch_shift_rows(clips, sr, steps); t1 = time.time()
report("pitch_shift", got, ref_ps, {"call_s": t1 - t0, "finite": bool(all(np.isfinite(g).all() for g in got))})

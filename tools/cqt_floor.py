"""Where the cqt error comes from: config 3 over the suite through whichever library B2A_LIBRARY
names; worst error per family and per octave (rows 12 o .. 12 o + 11 come from octave 6 - o's signal).

    B2A_LIBRARY=.../build/libb2a_<tag>.so python tools/cqt_floor.py [n_clips] [tag]
"""
import json, os, sys
from concurrent.futures import ProcessPoolExecutor
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 405
TAG = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("B2A_LIBRARY", "product")


def _ref(c):
    return L.audio_cqt(L.pcm16_to_float(c), duration=5.0)


def main():
    pcm = synth.make_suite(N, 22050, 110250, seed=1234)
    cache = f"/tmp/cqt_floor_ref_{N}.npy"
    if os.path.exists(cache):
        ref = np.load(cache)
    else:
        with ProcessPoolExecutor() as ex:
            ref = np.stack(list(ex.map(_ref, pcm, chunksize=4)))
        np.save(cache, ref)
    cfg = B.default_config(B.KIND_CQT)
    cfg.n_samples = 110250
    with B.Engine(cfg, 0) as e:
        got = e.run_host(pcm)
    err = np.abs(got - ref)                               # (clips, 84, 216)
    clip = err.reshape(N, -1).max(axis=1)
    fam = [float(clip[f::synth.N_FAMILIES].max()) for f in range(synth.N_FAMILIES)]
    octv = [float(err[:, 12 * o:12 * o + 12].max()) for o in range(7)]
    print(json.dumps(dict(tag=TAG, clips=N, max_abs=float(clip.max()), p99=float(np.percentile(clip, 99)),
                          per_family_max=fam, per_octave_rows_max=octv)), flush=True)


if __name__ == "__main__":
    main()

"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth


def run(kind, n, n_clips, **kw):
    cfg = B.default_config(kind)
    cfg.n_samples = n
    for k, v in kw.items():
        setattr(cfg, k, v)
    pcm = synth.make_suite(n_clips, cfg.sample_rate, n, seed=3)
    with B.Engine(cfg, 0) as e:
        out = e.run_host(pcm)
    print(kind, kw, out.shape, float(out.min()), float(out.max()), flush=True)


run(B.KIND_MEL, 8000, 320)                                                     # warp-specialised kernel, 2-3 clips per CTA
run(B.KIND_MFCC, 8000, 320, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=13)
run(B.KIND_MFCC, 8000, 40, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=20)
run(B.KIND_MEL, 8000, 40, n_fft=1024, hop_length=256, n_mels=64)               # generic kernel
run(B.KIND_CQT, 22050, 8)

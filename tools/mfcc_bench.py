"""audio_mfcc_seq (config 2: per-clip DCT; 12 coefficients: per-tile DCT) and audio_mel_spec
device-resident throughput, with a 42-clip parity check — profiling aid.   python tools/mfcc_bench.py"""
import sys, json
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L


def run(kind, n_clips, **kw):
    cfg = B.default_config(kind)
    for k, v in kw.items(): setattr(cfg, k, v)
    with B.Engine(cfg, 0) as e:
        x = (torch.randn((n_clips, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
        out = torch.empty((n_clips, e.rows, e.frames), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        err = None
        if kind == B.KIND_MFCC:
            pcm = synth.make_suite(42, 16000, 80000, seed=5)
            got = e.run_host(pcm)
            ref = np.stack([L.audio_mfcc_seq(L.pcm16_to_float(c), 16000, kw["n_mfcc"], 512, 160, 5.0, n_mels=40) for c in pcm])
            err = float(np.abs(got - ref).max())
    print(json.dumps(dict(kind=kind, cfg=kw, clips_per_s=n_clips / ms * 1e3,
                          ms=ms, max_abs_42=err)), flush=True)


run(B.KIND_MFCC, 40000, n_samples=80000, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=13)
run(B.KIND_MFCC, 40000, n_samples=80000, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=12)
run(B.KIND_MEL, 40000, n_samples=80000)

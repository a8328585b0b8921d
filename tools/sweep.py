"""Throughput sweep over configurations (device-resident, CUDA events) — profiling aid."""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B

def run(kind, n_clips, **kw):
    cfg = B.default_config(kind)
    for k, v in kw.items(): setattr(cfg, k, v)
    e = B.Engine(cfg, 0)
    x = (torch.randn((n_clips, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
    if cfg.input_dtype == B.IN_F32:
        x = x.to(torch.float32) / 32768.0
    out = torch.empty((n_clips, e.rows, e.frames), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    secs = cfg.n_samples / cfg.sample_rate
    print(json.dumps(dict(kind=kind, cfg=kw, rows=e.rows, frames=e.frames, clips_per_s=n_clips / ms * 1e3,
                          audio_s_per_s=n_clips * secs / ms * 1e3, ms=ms)), flush=True)
    e.close()

run(B.KIND_MEL, 20000, n_samples=80000)
run(B.KIND_MEL, 20000, n_samples=80000, input_dtype=B.IN_F32)
run(B.KIND_MEL, 20000, n_samples=80000, n_fft=1024, hop_length=256, n_mels=64)
run(B.KIND_MEL, 10000, n_samples=110250, sample_rate=22050, n_fft=2048, hop_length=512, n_mels=128)
run(B.KIND_MEL, 20000, n_samples=80000, n_fft=256, hop_length=128, n_mels=40)
run(B.KIND_MFCC, 20000, n_samples=80000, sample_rate=16000, n_fft=512, hop_length=160, n_mels=40, n_mfcc=13)
run(B.KIND_MFCC, 10000, n_samples=110250)          # reference defaults 22050/1024/512/128 -> 40
run(B.KIND_CQT, 4096, n_samples=110250)


def run_resampler(orig, target, seconds=5.0, reps=200):
    n = int(seconds * orig)
    x = (torch.randn(n, device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
    with B.Resampler(orig, target) as r:
        out = torch.empty(r.out_len(n), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(5): r.run_device(x.data_ptr(), B.IN_I16, n, out.data_ptr(), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): r.run_device(x.data_ptr(), B.IN_I16, n, out.data_ptr(), st)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        up, down, half, poly = B.resampler_design(orig, target)
    macs = r.out_len(n) * poly.shape[1] if False else int(np.ceil(n * target / orig)) * poly.shape[1]
    print(json.dumps(dict(resampler=f"{orig}->{target}", clip_s=seconds, us_per_clip=ms * 1e3, clips_per_s=1e3 / ms,
                          taps_per_output=int(poly.shape[1]), gmac_per_s=macs / ms / 1e6)), flush=True)

run_resampler(44100, 16000)
run_resampler(48000, 16000)
run_resampler(22050, 16000)
run_resampler(44100, 22050)

"""audio_mel_spec at n_fft 1024 / hop 256 / 64 mels (aligned 5 s 16 kHz clips) on the warp-specialised 1024 kernel —
a short run for ncu:  ncu --set full -k regex:logmel1024 -s 1 -c 1 python tools/prof_1024.py [mfcc]"""
import sys; sys.path.insert(0, ".")
import torch
from audio_edge_ml_pipeline_b200 import _lib as B
mf = len(sys.argv) > 1 and sys.argv[1] == "mfcc"
cfg = B.default_config(B.KIND_MFCC if mf else B.KIND_MEL)
if mf:
    cfg.n_samples = 110256            # reference defaults, clip length rounded up to a 16-byte multiple
else:
    cfg.n_samples = 80000; cfg.n_fft = 1024; cfg.hop_length = 256; cfg.n_mels = 64
e = B.Engine(cfg, 0); n = 4000
x = (torch.randn((n, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
out = torch.empty((n, e.rows, e.frames), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): e.run_device(x.data_ptr(), n, out.data_ptr(), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); e.run_device(x.data_ptr(), n, out.data_ptr(), st); b.record(); torch.cuda.synchronize()
print("ok", e.rows, e.frames, "clips/s", n / a.elapsed_time(b) * 1e3)

import sys; sys.path.insert(0, ".")
import torch
from audio_edge_ml_pipeline_b200 import _lib as B
cfg = B.default_config(B.KIND_MEL); cfg.n_samples=80000; cfg.n_fft=1024; cfg.hop_length=256; cfg.n_mels=64
e = B.Engine(cfg, 0); n=8000
x = (torch.randn((n, 80000), device="cuda")*3276.8).round().clamp(-32768,32767).to(torch.int16)
out = torch.empty((n, e.rows, e.frames), dtype=torch.float32, device="cuda")
for _ in range(3): e.run_device(x.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("ok")

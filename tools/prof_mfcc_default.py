"""audio_mfcc_seq at the reference defaults (22 050 Hz, n_fft 1024, hop 512, 128 mels -> 40) on the generic
kernel — a short run for ncu:  ncu --set full -k regex:front_kernel -s 1 -c 1 python tools/prof_mfcc_default.py"""
import sys; sys.path.insert(0, ".")
import torch
from audio_edge_ml_pipeline_b200 import _lib as B
cfg = B.default_config(B.KIND_MFCC); cfg.n_samples = 110250
e = B.Engine(cfg, 0); n = 4000
x = (torch.randn((n, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
out = torch.empty((n, e.rows, e.frames), dtype=torch.float32, device="cuda")
for _ in range(3): e.run_device(x.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("ok", e.rows, e.frames)

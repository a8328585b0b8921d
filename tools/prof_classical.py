"""audio_classical at the reference defaults, device-resident — a short run for timing and ncu:
   ncu --set full -k regex:classical -s 1 -c 1 python tools/prof_classical.py"""
import sys; sys.path.insert(0, ".")
import torch
from audio_edge_ml_pipeline_b200 import _lib as B
cfg = B.default_config(B.KIND_CLASSICAL); cfg.n_samples = 110250
e = B.Engine(cfg, 0); n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
x = (torch.randn((n, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
out = torch.empty((n, e.rows), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): e.run_device(x.data_ptr(), n, out.data_ptr(), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); e.run_device(x.data_ptr(), n, out.data_ptr(), st); b.record(); torch.cuda.synchronize()
print("ok", e.rows, "clips/s", n / a.elapsed_time(b) * 1e3, "ms", a.elapsed_time(b))

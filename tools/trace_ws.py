"""Hand-off timeline of CTA 0 of the warp-specialised logmel512 kernel (debug build only):
    B2A_NVCC_EXTRA=-DB2A_TRACE python -m audio_edge_ml_pipeline_b200.build --force && python tools/trace_ws.py
slots: 0 TMA issue, 1/2 FFT warp 0 raw wait start/end, 3/4 FFT warp 15, 5/6 mel warp 0 pow_full wait start/end."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B

cfg = B.default_config(B.KIND_MEL); cfg.n_samples = 80000
e = B.Engine(cfg, 0)
n = 148 * 12
x = (torch.randn((n, 80000), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
out = torch.empty((n, e.rows, e.frames), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): e.run_device(x.data_ptr(), n, out.data_ptr(), st)
torch.cuda.synchronize()
lib = B.load_library()
buf = (ctypes.c_longlong * (8 * 512))()
lib.b2a_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.b2a_debug_trace_read(buf, 8 * 512))
t = np.array(buf, dtype=np.int64).reshape(8, 512)
t0 = t[1, 0]
print("tile | issue  | w0 wait start/end (dur) | w15 wait start/end (dur) | mel0 pow_full start/end (dur)")
for k in range(60, 110):
    r = t[:, k] - t0
    print(f"{k:4d} | {r[0]:7d} | {r[1]:7d} {r[2]:7d} ({r[2]-r[1]:5d}) | {r[3]:7d} {r[4]:7d} ({r[4]-r[3]:5d}) | {r[5]:7d} {r[6]:7d} ({r[6]-r[5]:5d})")
d = t[:, 16:192]
print("mean tile period", np.diff(d[1]).mean(), "w0 raw wait", (d[2]-d[1]).mean(), "w15 raw wait", (d[4]-d[3]).mean(),
      "mel0 wait", (d[6]-d[5]).mean(), "issue lead (issue k -> w0 needs k)", (d[1]-d[0]).mean())

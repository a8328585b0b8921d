"""audio_cqt (config 3) device-resident throughput + parity on a few clips — profiling aid.
   python tools/cqt_bench.py [n_clips]"""
import sys, json
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from audio_edge_ml_pipeline_b200 import _lib as B
from audio_edge_ml_pipeline_b200 import synth
from oracle import librosa_restated as L

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = B.default_config(B.KIND_CQT); cfg.n_samples = 110250
with B.Engine(cfg, 0) as e:
    pcm = synth.make_suite(28, 22050, 110250, seed=77)
    got = e.run_host(pcm)
    ref = np.stack([L.audio_cqt(L.pcm16_to_float(c), duration=5.0) for c in pcm])
    err = float(np.abs(got - ref).max())
    x = (torch.randn((n_clips, cfg.n_samples), device="cuda") * 3276.8).round().clamp(-32768, 32767).to(torch.int16)
    out = torch.empty((n_clips, e.rows, e.frames), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): e.run_device(x.data_ptr(), n_clips, out.data_ptr(), st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
print(json.dumps(dict(cqt_clips_per_s=n_clips / ms * 1e3, ms=ms, n_clips=n_clips, max_abs_vs_oracle_28=err)), flush=True)

"""cProfile of one config-5 run (loader scan + extract_dataset + save) after a warm-up pass."""
import sys, shutil, cProfile, pstats, io; sys.path.insert(0, ".")
from pathlib import Path
import numpy as np
import audio_edge_ml_pipeline_b200 as P
from audio_edge_ml_pipeline_b200 import synth, wavio
from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader
root = Path("/dev/shm/b2a_prof"); shutil.rmtree(root, ignore_errors=True)
rng = np.random.default_rng(1)
pool = [synth.pad_or_trim_pcm(synth.to_pcm16(synth.make_clip(rng, k % 5, 16000, 80000)), 80000) for k in range(16)]
for c in range(27):
    d = root / "ds" / f"class_{c:02d}"; d.mkdir(parents=True)
    for i in range(260): wavio.write_wav_pcm16(d / f"clip_{i:04d}.wav", np.roll(pool[(c + i) % 16], 37 * i), 16000)
ext = P.get("audio_mel_spec")(duration=5.0, n_mels=40, sample_rate=16000, n_fft=512, hop_length=160, devices=[0])
def once():
    loader = AudioFolderLoader(root / "ds")
    fs = P.FeaturePipeline(loader, ext).run(output_dir=root / "out")
    P.FeaturePipeline.save(fs, root / "out")
once(); once()
pr = cProfile.Profile(); pr.enable(); once(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
ext.close(); shutil.rmtree(root, ignore_errors=True)

"""Where config 5's extract time goes: native decode of one window vs the host-buffer engine call on it."""
import sys, time, shutil; sys.path.insert(0, ".")
from pathlib import Path
import numpy as np
import audio_edge_ml_pipeline_b200 as P
from audio_edge_ml_pipeline_b200 import _lib as B, synth, wavio, extractors
root = Path("/dev/shm/b2a_breakdown"); shutil.rmtree(root, ignore_errors=True); root.mkdir(parents=True)
rng = np.random.default_rng(1)
pool = [synth.pad_or_trim_pcm(synth.to_pcm16(synth.make_clip(rng, k % 5, 16000, 80000)), 80000) for k in range(16)]
paths = []
for i in range(4096):
    p = root / f"c{i:05d}.wav"; wavio.write_wav_pcm16(p, np.roll(pool[i % 16], 13 * i), 16000); paths.append(str(p))
ext = P.get("audio_mel_spec")(duration=5.0, n_mels=40, sample_rate=16000, n_fft=512, hop_length=160, devices=[0])
a_in = ext._stage(("i16", extractors.HOST_BATCH_CLIPS, 80000, "i2"), (80000,), np.int16)
out = np.empty((4096, 40, 501), np.float32)
for rep in range(3):
    t0 = time.perf_counter()
    st = B.decode_wav_pcm16_batch(paths, 16000, 80000, a_in, None, None, extractors.DECODE_WORKERS)
    t1 = time.perf_counter()
    ext.extract_batch(a_in[:4096], out)
    t2 = time.perf_counter()
    print(f"rep {rep}: decode 4096 files {t1 - t0:.4f} s ({4096 * 160e3 / (t1 - t0) / 1e9:.1f} GB/s), engine {t2 - t1:.4f} s, workers {extractors.DECODE_WORKERS}")
ext.close(); shutil.rmtree(root, ignore_errors=True)

#!/usr/bin/env python
"""End-to-end Stage-2 run (BASELINE config 5): an fsc22-shaped, augmentation-expanded set of
PCM16 WAV files -> loader scan -> decode -> pinned H2D -> log-mel on every visible GPU -> D2H ->
features.npy, through the reference-facing API (registered extractor + FeaturePipeline).

    python bench_stage2.py [--classes 27 --per-class 260] [--devices all] [--dir /dev/shm/b2a_stage2]

27 x 260 = 7020 clips = the reference's train split (52/class) x (1 + n_augments=4)
(config/augmentation.yaml:18-19,25).  The WAVs are synthetic (the reference ships no audio).
Prints one JSON line; host-bound by design (decode + np.save), reported beside the kernel bench.
"""
import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def run(classes: int = 27, per_class: int = 260, devices: str = "all", dir: str = "/dev/shm/b2a_stage2",
        repeat: int = 3, file_rate: int = 16000) -> dict:
    """file_rate != 16000 writes the same clips as files recorded at another rate (5 s each): Stage 2 then
    decodes them natively at that rate and resamples on the device (deep.py:44-50 -> librosa.load)."""
    args = argparse.Namespace(classes=classes, per_class=per_class, devices=devices, dir=dir, repeat=repeat)
    import audio_edge_ml_pipeline_b200 as P
    from audio_edge_ml_pipeline_b200 import _lib as B
    from audio_edge_ml_pipeline_b200 import synth, wavio
    from audio_edge_ml_pipeline_b200.loaders import AudioFolderLoader

    root = Path(args.dir)
    if root.exists():
        shutil.rmtree(root)
    ds = root / "fsc22_device_augmented"
    rng = np.random.default_rng(2026)
    n_file = 5 * file_rate
    pool = [synth.pad_or_trim_pcm(synth.to_pcm16(synth.make_clip(rng, k % 5, file_rate, n_file)), n_file) for k in range(40)]
    t0 = time.perf_counter()
    n = 0
    for c in range(args.classes):
        d = ds / f"class_{c:02d}"
        d.mkdir(parents=True)
        for i in range(args.per_class):
            wavio.write_wav_pcm16(d / f"clip_{i:04d}.wav", np.roll(pool[(c + i) % len(pool)], 37 * i), file_rate)
            n += 1
    gen_s = time.perf_counter() - t0

    ext = P.get("audio_mel_spec")(duration=5.0, n_mels=40, sample_rate=16000, n_fft=512, hop_length=160,
                                  devices=args.devices)
    n_dev = len(ext.devices)
    times = []
    for r in range(args.repeat + 1):            # first pass = warm-up (engine creation, page cache)
        t0 = time.perf_counter()
        loader = AudioFolderLoader(ds)
        t1 = time.perf_counter()
        fs = P.FeaturePipeline(loader, ext).run(output_dir=root / "out")
        t2 = time.perf_counter()
        P.FeaturePipeline.save(fs, root / "out")
        t3 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
    best = min(times[1:], key=lambda x: x[3])
    assert fs.features.shape == (n, 40, 501)
    line = {
        "metric": "Stage-2 end-to-end clips/sec (WAV files -> features.npy), audio_mel_spec",
        "value": n / best[3], "unit": "clips/s", "n_gpus": n_dev, "clips": n,
        "seconds": {"loader_scan": best[0], "decode+h2d+kernel+d2h": best[1], "save (features.npy rows are written in place during the run; labels, json)": best[2], "total": best[3]},
        "extract_only_clips_per_s": n / best[1], "features_bytes": int(fs.features.nbytes),
        "decode_workers": P.extractors.DECODE_WORKERS, "host_cores": os.cpu_count(),
        "dataset": f"{args.classes} classes x {args.per_class} PCM16 5 s {file_rate} Hz WAVs on {root}", "file_rate": file_rate,
        "dataset_write_s": gen_s, "device_count_visible": B.device_count(),
    }
    ext.close()
    shutil.rmtree(root, ignore_errors=True)
    return line


def run_device_augmented(classes: int = 27, per_class: int = 52, n_augments: int = 4, devices: str = "0",
                         out_dir: str = "/dev/shm/b2a_stage2_aug", full_chain: bool = False) -> dict:
    """Config 5 without the WAV round trip: the 27 x 52 originals are augmented on the device (Stage 1b,
    augment.py's chain minus the two librosa-backed steps, host-drawn reference RNG), quantised like the
    PCM16 files the reference writes, and go straight into audio_mel_spec -> features.npy."""
    import audio_edge_ml_pipeline_b200 as P
    from audio_edge_ml_pipeline_b200 import augment as G
    from audio_edge_ml_pipeline_b200 import synth
    rng = np.random.default_rng(2026)
    pool = np.stack([synth.pad_or_trim_pcm(synth.to_pcm16(synth.make_clip(rng, k % 5, 16000, 80000)), 80000)
                     for k in range(40)])
    originals = np.stack([np.roll(pool[(c + i) % len(pool)], 37 * i) for c in range(classes) for i in range(per_class)])
    chain = [{"type": "volume_scale", "min_gain": 0.7, "max_gain": 1.3},
             {"type": "gaussian_noise", "min_amplitude": 0.001, "max_amplitude": 0.004},
             {"type": "time_shift", "max_fraction": 0.2}]
    if full_chain:                                   # config/augmentation.yaml:38-60, the reference's default chain
        chain = [{"type": "volume_scale", "min_gain": 0.7, "max_gain": 1.3},
                 {"type": "pdm_hiss", "min_amplitude": 0.01, "max_amplitude": 0.04},
                 {"type": "gaussian_noise", "min_amplitude": 0.001, "max_amplitude": 0.004},
                 {"type": "time_stretch", "min_rate": 0.85, "max_rate": 1.15},
                 {"type": "pitch_shift", "min_steps": -3.0, "max_steps": 3.0},
                 {"type": "time_shift", "max_fraction": 0.2}]
    ext = P.get("audio_mel_spec")(duration=5.0, n_mels=40, sample_rate=16000, n_fft=512, hop_length=160, devices=devices)
    dev0 = ext.devices[0]
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        if full_chain:                               # staged: draws and device calls interleave (augment._run_staged)
            aug = G.augment_batch(originals, chain, n_augments, 42, out_dtype=np.int16, device=dev0, sample_rate=16000)
            t1 = t0
            rows = len(aug)
        else:
            src_clip, steps, noise, max_steps = G.plan([80000] * len(originals), [chain] * len(originals), n_augments, 42)
            t1 = time.perf_counter()
            rows = len(src_clip)
            aug = np.empty((rows, 80000), dtype=np.int16)
            G._run_host(dev0, originals.reshape(-1), np.ascontiguousarray(src_clip * 80000), np.full(rows, 80000, np.int32),
                        np.arange(rows, dtype=np.int64) * 80000, steps, max_steps, noise, aug.reshape(-1))
        t2 = time.perf_counter()
        feats = ext.extract_batch(aug)
        t3 = time.perf_counter()
        Path(out_dir).mkdir(parents=True, exist_ok=True)
        np.save(Path(out_dir) / "features.npy", feats)
        t4 = time.perf_counter()
        cur = (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)
        best = cur if best is None or cur[4] < best[4] else best
    ext.close()
    shutil.rmtree(out_dir, ignore_errors=True)
    return {"metric": "Stage 1b + 2 without the WAV round trip: clips/sec (originals in memory -> features.npy)",
            "value": rows / best[4], "unit": "clips/s", "clips": rows, "originals": len(originals),
            "seconds": ({"augment (host draws in the reference's order interleaved with the staged device calls)": best[0] + best[1]}
                        if full_chain else
                        {"host_rng_plan (numpy default_rng, the reference's sequence)": best[0],
                         "device_augment (H2D + kernel + D2H int16)": best[1]}) |
                       {"log-mel (host buffers)": best[2], "np.save": best[3], "total": best[4]},
            "chain": [c["type"] for c in chain], "features_shape": list(feats.shape)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=27)
    ap.add_argument("--per-class", type=int, default=260)
    ap.add_argument("--devices", default="all")
    ap.add_argument("--dir", default="/dev/shm/b2a_stage2")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--file-rate", type=int, default=16000, help="rate the WAV files are written at (!= 16000: device resampling)")
    ap.add_argument("--device-augmented", action="store_true", help="config 5 without the WAV round trip (Stage 1b on the device)")
    ap.add_argument("--full-chain", action="store_true", help="with --device-augmented: the reference's default chain incl. time_stretch / pitch_shift")
    a = ap.parse_args()
    if a.device_augmented:
        print(json.dumps(run_device_augmented(a.classes, devices=a.devices if a.devices != "all" else "0", full_chain=a.full_chain)), flush=True)
        return
    print(json.dumps(run(a.classes, a.per_class, a.devices, a.dir, a.repeat, a.file_rate)), flush=True)


if __name__ == "__main__":
    main()

"""Stage 1b — audio augmentation on the GPU.

Host-side mirror of the reference's ``src/preprocessing/augment.py``: the same YAML keys and defaults
(``load_config`` :219-240), the same per-class / per-file / per-copy loop and — because every random
number is drawn HERE, with ``np.random.default_rng(seed)`` consumed in the reference's order
(augment.py:325, 88-132) — the same augmented waveforms, bit for bit.  The arithmetic (gain, noise add +
clip, cyclic shift, polarity, level match) runs in one CUDA kernel per batch (``csrc/augment.cu``); there
is no CPU path: without the library or a CUDA device :func:`augment_batch` raises.

``pdm_hiss`` (augment.py:135-167) is a noise SOURCE like ``gaussian_noise``: its pink, notched, unit-RMS
noise row is synthesised on the host next to the draw that feeds it (the same ``np.fft`` calls on the same
``rng.standard_normal(n)``, hence the same float32 row as the reference), and the device mixes it into the
clip exactly like white noise (``clip(y + row * amplitude)``).

``time_stretch`` / ``pitch_shift`` (augment.py:105-118 -> librosa.effects) change the clip as a whole — and the
first one its length — so a chain that names them runs in stages: the element-wise steps before, between and after
them stay fused in the augmentation kernel, each vocoder step is one batched call into ``csrc/effects.cu`` for every
row that reaches it, and the lengths that later draws depend on (``standard_normal(len(y))``, the roll of
``time_shift``) are tracked at draw time (``int(round(n / rate))``, as librosa computes it), so the random stream
is still consumed exactly like the reference consumes it.

Two ways out of Stage 1b:
  * :func:`run` writes the class-per-folder WAV tree the reference writes (PCM16 via soundfile there;
    the same quantisation rule here), for Stage 2's ``audio_folder`` loader;
  * :func:`augment_batch` with ``out_dtype=np.int16`` hands Stage 2 the samples it would have read back
    from those WAVs, skipping the file round trip (BASELINE config 5's input set).
"""
from __future__ import annotations

import ctypes as C
import json
import logging
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import _lib as B
from . import wavio

logger = logging.getLogger(__name__)

AUG_END, AUG_GAIN, AUG_NOISE, AUG_ROLL, AUG_POLARITY = -1, 0, 1, 2, 3
ELEMENTWISE = ("volume_scale", "gaussian_noise", "time_shift", "polarity_inversion", "pdm_hiss")
VOCODER = ("time_stretch", "pitch_shift")          # librosa.effects: csrc/effects.cu, one batched call per step
SUPPORTED = ELEMENTWISE + VOCODER
NOT_BUILT = ()
VALID_TYPES = sorted(SUPPORTED)


class AugStep(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_float), ("shift", C.c_int32), ("reserved", C.c_int32),
                ("noise_off", C.c_int64)]


STEP_DTYPE = np.dtype([("op", "<i4"), ("a", "<f4"), ("shift", "<i4"), ("reserved", "<i4"), ("noise_off", "<i8")])
assert STEP_DTYPE.itemsize == C.sizeof(AugStep) == 24


def load_config(path) -> dict:
    """augment.py:219-240 — same required key and defaults."""
    import yaml
    with Path(path).open() as fh:
        cfg = yaml.safe_load(fh) or {}
    for key in ("output_dir",):
        if key not in cfg:
            raise ValueError(f"augmentation.yaml must include '{key}'.")
    cfg.setdefault("n_augments", 4)
    cfg.setdefault("preserve_length", True)
    cfg.setdefault("seed", 42)
    cfg.setdefault("sample_rate", None)
    cfg.setdefault("augmentations", [])
    cfg.setdefault("class_overrides", {})
    cfg.setdefault("loader", "audio_folder")
    cfg.setdefault("split", "train")
    cfg.setdefault("level_match_db", 0.0)
    return cfg


def check_specs(aug_specs: Sequence[dict]) -> None:
    for spec in aug_specs:
        t = spec["type"]
        if t not in VALID_TYPES:
            raise ValueError(f"Unknown augmentation type '{t}'. Valid types: {VALID_TYPES}")    # augment.py:196-200


def _pink_row(white: np.ndarray, sr: int, notch_freq: float) -> np.ndarray:
    """The noise row of pdm_hiss (augment.py:147-165): white -> 1/sqrt(f) shaping -> +-2-bin notch -> unit RMS,
    float32.  Noise synthesis stays with the generator on the host (numpy's own rfft / irfft, the calls the
    reference makes), so the row is the reference's bit for bit; the device only mixes it in."""
    n = len(white)
    fft = np.fft.rfft(white)
    freqs = np.fft.rfftfreq(n, d=1.0 / sr)
    freqs[0] = 1.0
    fft /= np.sqrt(freqs)
    pink = np.fft.irfft(fft, n=n).astype(np.float32)
    fft2 = np.fft.rfft(pink)
    fft2[np.abs(np.fft.rfftfreq(n, d=1.0 / sr) - notch_freq) < (sr / n * 2)] = 0.0
    pink = np.fft.irfft(fft2, n=n).astype(np.float32)
    pink /= np.sqrt(np.mean(pink ** 2)) + 1e-9
    return pink


def plan(lengths: Sequence[int], specs_per_clip: Sequence[Sequence[dict]], n_augments: int, seed: int,
         level_match_db: float = 0.0, include_originals: bool = True, rng: Optional[np.random.Generator] = None,
         sample_rate: int = 16000):
    """Draw every random parameter in the reference's order and lay out the device work.

    Clips are taken in the order given (the reference walks classes sorted by name, files in loader
    order, copies 1..n_augments, steps in specification order: augment.py:341-364).  Returns
    ``(src_clip, steps, noise, max_steps)``: per output row the index of its source clip, the
    ``(rows, max_steps)`` step table, and the concatenated float32 noise rows."""
    rng = np.random.default_rng(seed) if rng is None else rng
    scale = 10.0 ** (float(level_match_db) / 20.0)                       # augment.py:318
    lead = 1 if scale != 1.0 else 0
    max_steps = lead + max((len(s) for s in specs_per_clip), default=0)
    max_steps = max(max_steps, 1)
    rows_per = n_augments + (1 if include_originals else 0)
    steps = np.zeros((len(lengths) * rows_per, max_steps), dtype=STEP_DTYPE)
    steps["op"] = AUG_END
    src_clip = np.repeat(np.arange(len(lengths), dtype=np.int64), rows_per)
    noise_rows: list = []
    noise_pos = 0
    r = 0
    for n, specs in zip(lengths, specs_per_clip):
        check_specs(specs)
        if any(sp["type"] in VOCODER for sp in specs):
            raise ValueError("plan() lays out element-wise chains only; chains with time_stretch / pitch_shift go "
                             "through augment_ragged / augment_batch (staged execution)")
        n = int(n)
        for copy in range(rows_per):
            k = 0
            if lead:
                steps[r, k] = (AUG_GAIN, np.float32(scale), 0, 0, 0)     # y = (y * scale).astype(float32)
                k += 1
            if not (include_originals and copy == 0):
                for spec in specs:
                    t = spec["type"]
                    if t == "volume_scale":
                        gain = rng.uniform(spec.get("min_gain", 0.7), spec.get("max_gain", 1.3))
                        steps[r, k] = (AUG_GAIN, np.float32(gain), 0, 0, 0)
                    elif t == "gaussian_noise":
                        amp = rng.uniform(spec.get("min_amplitude", 0.001), spec.get("max_amplitude", 0.008))
                        noise_rows.append(rng.standard_normal(n).astype(np.float32))
                        steps[r, k] = (AUG_NOISE, np.float32(amp), 0, 0, noise_pos)
                        noise_pos += n
                    elif t == "pdm_hiss":
                        white = rng.standard_normal(n)                   # drawn BEFORE the amplitude (augment.py:146, 165)
                        noise_rows.append(_Pink(white, sample_rate, spec.get("notch_freq", 4000.0)))
                        amp = rng.uniform(spec.get("min_amplitude", 0.02), spec.get("max_amplitude", 0.08))
                        steps[r, k] = (AUG_NOISE, np.float32(amp), 0, 0, noise_pos)
                        noise_pos += n
                    elif t == "time_shift":
                        f = spec.get("max_fraction", 0.2)
                        steps[r, k] = (AUG_ROLL, 0.0, int(rng.uniform(-f, f) * n), 0, 0)
                    elif t == "polarity_inversion":
                        steps[r, k] = (AUG_POLARITY, 0.0, 0, 0, 0)
                    k += 1
            r += 1
    noise = np.concatenate(_materialise(noise_rows)) if noise_rows else np.zeros(0, np.float32)
    return src_clip, steps, noise, max_steps


class _Pink:
    """A pdm_hiss noise row still to be synthesised from its white draw: the draw has to happen in sequence (one
    generator), the four FFTs behind it do not — :func:`_materialise` runs them on a thread pool."""
    __slots__ = ("white", "sr", "notch")

    def __init__(self, white, sr, notch):
        self.white, self.sr, self.notch = white, sr, notch

    def __len__(self):
        return len(self.white)


def _materialise(rows: list) -> list:
    """Replace every pending :class:`_Pink` in ``rows`` by its float32 noise row (numpy's FFT releases the GIL)."""
    todo = [i for i, r in enumerate(rows) if isinstance(r, _Pink)]
    if not todo:
        return rows
    from concurrent.futures import ThreadPoolExecutor
    import os
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        done = list(pool.map(lambda i: _pink_row(rows[i].white, rows[i].sr, rows[i].notch), todo))
    for i, d in zip(todo, done):
        rows[i] = d
    return rows


def _draw_ops(rng, n: int, specs, sample_rate: int) -> list:
    """One augmented copy's chain with every random parameter drawn, in specification order (augment.py:186-203):
    a list of ("gain", g) / ("noise", amp, row) / ("roll", shift) / ("pol",) / ("stretch", rate) /
    ("pitch", n_steps).  ``n`` follows the chain: after a time_stretch the clip has int(round(n / rate)) samples."""
    ops = []
    for spec in specs:
        t = spec["type"]
        if t == "volume_scale":
            ops.append(("gain", np.float32(rng.uniform(spec.get("min_gain", 0.7), spec.get("max_gain", 1.3)))))
        elif t == "gaussian_noise":
            amp = rng.uniform(spec.get("min_amplitude", 0.001), spec.get("max_amplitude", 0.008))
            ops.append(("noise", np.float32(amp), rng.standard_normal(n).astype(np.float32)))
        elif t == "pdm_hiss":
            white = rng.standard_normal(n)                   # drawn BEFORE the amplitude (augment.py:146, 165)
            amp = rng.uniform(spec.get("min_amplitude", 0.02), spec.get("max_amplitude", 0.08))
            ops.append(("noise", np.float32(amp), _Pink(white, sample_rate, spec.get("notch_freq", 4000.0))))
        elif t == "time_shift":
            f = spec.get("max_fraction", 0.2)
            ops.append(("roll", int(rng.uniform(-f, f) * n)))
        elif t == "polarity_inversion":
            ops.append(("pol",))
        elif t == "time_stretch":
            rate = rng.uniform(spec.get("min_rate", 0.85), spec.get("max_rate", 1.15))
            ops.append(("stretch", float(rate)))
            n = int(round(n / rate))
        elif t == "pitch_shift":
            ops.append(("pitch", float(rng.uniform(spec.get("min_steps", -3.0), spec.get("max_steps", 3.0)))))
    return ops


def _run_staged(clips, specs_per_clip, n_augments, seed, level_match_db, include_originals, out_dtype, device, rng,
                sample_rate, preserve_length) -> list:
    """Chains that contain vocoder steps: rows advance in rounds — their next run of element-wise steps in one
    augmentation-kernel call, then their next vocoder step in one batched effects call — until every row is done."""
    rng = np.random.default_rng(seed) if rng is None else rng
    scale = 10.0 ** (float(level_match_db) / 20.0)
    rows_per = n_augments + (1 if include_originals else 0)
    cur, todo, orig_len = [], [], []
    for clip, specs in zip(clips, specs_per_clip):
        check_specs(specs)
        y = np.asarray(clip)
        y = y.astype(np.float32) / np.float32(32768.0) if y.dtype == np.int16 else y.astype(np.float32, copy=False)
        for copy in range(rows_per):
            ops = [("gain", np.float32(scale))] if scale != 1.0 else []
            if not (include_originals and copy == 0):
                ops += _draw_ops(rng, len(y), specs, sample_rate)
            cur.append(y)
            todo.append(ops)
            orig_len.append(len(y))
    while any(todo):
        # (1) the element-wise prefix of every row that has one
        idx = [i for i, ops in enumerate(todo) if ops and ops[0][0] not in ("stretch", "pitch")]
        if idx:
            segs = []
            for i in idx:
                k = 0
                while k < len(todo[i]) and todo[i][k][0] not in ("stretch", "pitch"):
                    k += 1
                segs.append(todo[i][:k])
                todo[i] = todo[i][k:]
            max_steps = max(len(sg) for sg in segs)
            steps = np.zeros((len(idx), max_steps), dtype=STEP_DTYPE)
            steps["op"] = AUG_END
            noise_rows, pos = [], 0
            for r, sg in enumerate(segs):
                for k, op in enumerate(sg):
                    if op[0] == "gain":
                        steps[r, k] = (AUG_GAIN, op[1], 0, 0, 0)
                    elif op[0] == "noise":
                        steps[r, k] = (AUG_NOISE, op[1], 0, 0, pos)
                        noise_rows.append(op[2])
                        pos += len(op[2])
                    elif op[0] == "roll":
                        steps[r, k] = (AUG_ROLL, 0.0, op[1], 0, 0)
                    else:
                        steps[r, k] = (AUG_POLARITY, 0.0, 0, 0, 0)
            lens = np.array([len(cur[i]) for i in idx], dtype=np.int32)
            off = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.int64)
            src = np.concatenate([cur[i] for i in idx])
            out = np.empty(src.size, dtype=np.float32)
            noise = np.concatenate(_materialise(noise_rows)) if noise_rows else np.zeros(0, np.float32)
            _run_host(device, src, off, lens, off, steps, max_steps, noise, out)
            for r, i in enumerate(idx):
                cur[i] = out[off[r]:off[r] + lens[r]]
        # (2) the vocoder step of every row that is waiting at one
        st = [i for i, ops in enumerate(todo) if ops and ops[0][0] == "stretch"]
        if st:
            got = B.time_stretch_rows([cur[i] for i in st], [todo[i][0][1] for i in st], device)
            for i, g in zip(st, got):
                cur[i], todo[i] = g, todo[i][1:]
        ps = [i for i, ops in enumerate(todo) if ops and ops[0][0] == "pitch"]
        if ps:
            got = B.pitch_shift_rows([cur[i] for i in ps], sample_rate, [todo[i][0][1] for i in ps], device)
            for i, g in zip(ps, got):
                cur[i], todo[i] = g, todo[i][1:]
    if preserve_length:                                                  # augment.py:206-212
        cur = [y[:n] if len(y) >= n else np.pad(y, (0, n - len(y))) for y, n in zip(cur, orig_len)]
    if np.dtype(out_dtype) == np.int16:
        cur = [quantize_pcm16(y) for y in cur]
    return [cur[i * rows_per:(i + 1) * rows_per] for i in range(len(clips))]


def _run_host(device, src, src_off, lengths, out_off, steps, max_steps, noise, out) -> None:
    lib = B.load_library()
    fn = lib.b2a_augment_host
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    fn.argtypes = [i32, vp, i32, i64, vp, vp, vp, i64, vp, i32, vp, i64, vp, i32, i64]
    fn.restype = C.c_int
    B._check(fn(int(device), src.ctypes.data, B.IN_I16 if src.dtype == np.int16 else B.IN_F32, src.size,
                src_off.ctypes.data, lengths.ctypes.data, out_off.ctypes.data, len(lengths),
                steps.ctypes.data, int(max_steps), noise.ctypes.data if noise.size else None, noise.size,
                out.ctypes.data, B.IN_I16 if out.dtype == np.int16 else B.IN_F32, out.size))


def augment_ragged(clips: Sequence[np.ndarray], specs_per_clip, n_augments: int = 4, seed: int = 42,
                   level_match_db: float = 0.0, include_originals: bool = True, out_dtype=np.float32,
                   device: int = 0, rng: Optional[np.random.Generator] = None, sample_rate: int = 16000,
                   preserve_length: bool = True) -> list:
    """Clips of any lengths (1-D int16 or float32, one dtype) -> per clip the list
    ``[original (level-matched), copy 1, ..., copy n_augments]`` (originals only when asked for)."""
    if len(clips) == 0:
        return []
    if any(sp["type"] in VOCODER for specs in specs_per_clip for sp in specs):
        return _run_staged(clips, specs_per_clip, n_augments, seed, level_match_db, include_originals, out_dtype,
                           device, rng, sample_rate, preserve_length)
    dt = np.int16 if clips[0].dtype == np.int16 else np.float32
    lens = np.array([len(c) for c in clips], dtype=np.int64)
    src_clip, steps, noise, max_steps = plan(lens, specs_per_clip, n_augments, seed, level_match_db,
                                             include_originals, rng, sample_rate)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    src = np.concatenate([np.asarray(c, dtype=dt) for c in clips]) if lens.sum() else np.zeros(0, dt)
    row_len = lens[src_clip].astype(np.int32)
    out_off = np.concatenate([[0], np.cumsum(row_len.astype(np.int64))[:-1]]).astype(np.int64)
    out = np.empty(int(row_len.sum()), dtype=out_dtype)
    _run_host(device, src, np.ascontiguousarray(starts[src_clip]), row_len, out_off, steps, max_steps, noise, out)
    rows_per = n_augments + (1 if include_originals else 0)
    rows = [out[o:o + n] for o, n in zip(out_off, row_len)]
    return [rows[i * rows_per:(i + 1) * rows_per] for i in range(len(clips))]


def augment_batch(clips: np.ndarray, aug_specs, n_augments: int = 4, seed: int = 42, level_match_db: float = 0.0,
                  include_originals: bool = True, out_dtype=np.float32, device: int = 0,
                  rng: Optional[np.random.Generator] = None, sample_rate: int = 16000) -> np.ndarray:
    """Equal-length clips ``(N, n)`` int16 / float32 -> ``(N * (1 + n_augments), n)`` in the reference's file
    order: each original (level-matched) followed by its copies.  ``aug_specs`` is one chain for every clip or
    a per-clip list of chains (class overrides, augment.py:337-339).  ``out_dtype=np.int16`` quantises like the
    PCM16 WAV the reference writes between the stages."""
    clips = np.ascontiguousarray(clips)
    if clips.dtype != np.int16:
        clips = clips.astype(np.float32, copy=False)
    if clips.ndim != 2:
        raise ValueError("clips must be (N, n_samples)")
    n_clips, n = clips.shape
    per_clip = aug_specs if (len(aug_specs) and isinstance(aug_specs[0], (list, tuple))) else [aug_specs] * n_clips
    if len(per_clip) != n_clips:
        raise ValueError("one augmentation chain per clip expected")
    if any(sp["type"] in VOCODER for specs in per_clip for sp in specs):
        groups = _run_staged(list(clips), per_clip, n_augments, seed, level_match_db, include_originals, out_dtype,
                             device, rng, sample_rate, True)
        return np.stack([row for grp in groups for row in grp])
    src_clip, steps, noise, max_steps = plan([n] * n_clips, per_clip, n_augments, seed, level_match_db,
                                             include_originals, rng, sample_rate)
    rows = len(src_clip)
    out = np.empty((rows, n), dtype=out_dtype)
    _run_host(device, clips.reshape(-1), np.ascontiguousarray(src_clip * n), np.full(rows, n, np.int32),
              np.arange(rows, dtype=np.int64) * n, steps, max_steps, noise, out.reshape(-1))
    return out


def quantize_pcm16(y: np.ndarray) -> np.ndarray:
    """float32 in [-1, 1] -> int16 as soundfile writes subtype PCM_16 (libsndfile with clipping on:
    lrintf(x * 32768) saturated to [-32768, 32767]); soundfile is absent here, so this rule is unpinned."""
    s = np.asarray(y, dtype=np.float32) * np.float32(32768.0)
    return np.clip(np.rint(s), -32768, 32767).astype(np.int16)


def _iter_audio_folder(cfg: dict):
    """augment.py:268-303 — (path, class) pairs of a class-per-subfolder tree, optional manifest filter."""
    audio_folder = cfg.get("audio_folder") or cfg.get("dataset")
    if not audio_folder:
        raise ValueError("augmentation.yaml must include 'audio_folder' when loader=audio_folder.")
    root = Path(audio_folder)
    extensions = {".wav", ".flac", ".mp3", ".ogg", ".aiff"}
    allowed = None
    if cfg.get("manifest"):
        manifest = json.loads(Path(cfg["manifest"]).read_text())
        allowed = set(manifest.get(cfg.get("split", "train"), []))
    for class_dir in sorted(root.iterdir()):
        if not class_dir.is_dir():
            continue
        for f in sorted(class_dir.iterdir()):
            if f.suffix.lower() not in extensions:
                continue
            if allowed is not None and f"{class_dir.name}/{f.name}" not in allowed:
                continue
            yield f, class_dir.name


def run(cfg: dict, device: int = 0) -> int:
    """augment.py:308-389 — scan, group by class (sorted), level-match, copy the original, write
    ``n_augments`` augmented copies per file as ``<stem>_aug%03d.wav``.  Returns the number of augmented
    files written.  Files are decoded at ``sample_rate`` (None = their own rate)."""
    if cfg.get("loader", "audio_folder") != "audio_folder":
        raise ValueError(f"Unknown loader '{cfg.get('loader')}'. Valid: ['audio_folder'] (the fsc22 loader is the reference's own)")
    output_dir = Path(cfg["output_dir"])
    n_aug, seed = int(cfg["n_augments"]), int(cfg["seed"])
    output_dir.mkdir(parents=True, exist_ok=True)
    by_class: dict = {}
    for path, cname in _iter_audio_folder(cfg):
        by_class.setdefault(cname, []).append(path)
    rng = np.random.default_rng(seed)                                     # ONE generator for the whole run
    written = 0
    for cname, paths in sorted(by_class.items()):
        (output_dir / cname).mkdir(exist_ok=True)
        specs = cfg["class_overrides"].get(cname, {}).get("augmentations", cfg["augmentations"])
        clips, rates = [], []
        for p_ in paths:
            y, sr = wavio.decode_wav(p_)
            y = y.astype(np.float32) / np.float32(32768.0) if y.dtype == np.int16 else y
            if cfg["sample_rate"] and sr != cfg["sample_rate"]:
                y, sr = wavio.resample_audio(y, sr, int(cfg["sample_rate"]), device), int(cfg["sample_rate"])
            clips.append(np.ascontiguousarray(y, dtype=np.float32))
            rates.append(sr)
        if len(set(rates)) > 1:
            raise ValueError(f"class {cname}: files at different rates {sorted(set(rates))}; set sample_rate in the config")
        groups = augment_ragged(clips, [specs] * len(clips), n_aug, seed, float(cfg["level_match_db"]),
                                include_originals=True, out_dtype=np.float32, device=device, rng=rng,
                                sample_rate=rates[0] if rates else 16000, preserve_length=bool(cfg["preserve_length"]))
        for p_, sr, grp in zip(paths, rates, groups):
            dest = output_dir / cname / p_.name
            if not dest.exists():
                wavio.write_wav_pcm16(dest, quantize_pcm16(grp[0]), sr)
            for i, ya in enumerate(grp[1:], start=1):
                wavio.write_wav_pcm16(output_dir / cname / f"{p_.stem}_aug{i:03d}.wav", quantize_pcm16(ya), sr)
                written += 1
        logger.info("  %-20s  %d orig -> %d total (%d augmented)", cname, len(paths), len(paths) * (1 + n_aug),
                    len(paths) * n_aug)
    return written


def main(argv=None) -> None:
    import argparse
    ap = argparse.ArgumentParser(prog="python -m audio_edge_ml_pipeline_b200.augment",
                                 description="Stage 1b — audio data augmentation (GPU)")
    ap.add_argument("--config", metavar="YAML", required=True)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    cfg = load_config(Path(args.config))
    print(f"wrote {run(cfg, args.device)} augmented files to {cfg['output_dir']}")


if __name__ == "__main__":
    main()

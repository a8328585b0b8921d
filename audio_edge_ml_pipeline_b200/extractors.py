"""The three registered extractors, B200-backed.

Drop-in for ``AudioMelSpectrogram`` / ``AudioCQT`` / ``AudioMFCCSequence`` of the reference
(``src/preprocessing/feature_extraction/audio/deep.py:75-134, 196-260, 268-328``): same ``name``,
``feature_type``, ``modality``, constructor keywords and defaults, ``extract`` signature (unknown
loader metadata swallowed), output dtype/shape/layout and skip-on-error behaviour.  The librosa
calls are replaced by hand-written sm_100a kernels reached through the C ABI (``_lib.Engine``);
there is no CPU path: without the library or a CUDA device every extract call raises.

``extract_dataset`` is overridden to batch clips through the GPU(s) while reproducing the
reference loop's observable behaviour (base.py:176-234): iteration order, ``max_samples`` by
enumeration index, per-sample skip with a warning, label index by first *successful*
occurrence, int32 labels, ``RuntimeError`` when nothing was extracted.

Documented extensions (defaults preserve the reference): ``n_mels`` on ``audio_mfcc_seq``
(librosa's 128), ``pad_mode`` on mel/mfcc (``"constant"`` = librosa >= 0.10), ``devices``.
"""

from __future__ import annotations

import logging
import os
import threading
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import _lib as B
from . import wavio
from .base import BaseFeatureExtractor, FeatureSet, assemble_feature_set
from .registry import register

logger = logging.getLogger(__name__)

HOST_BATCH_CLIPS = 4096          # clips decoded and shipped per GPU round in extract_dataset
DECODE_WORKERS = min(16, os.cpu_count() or 1)   # file reads release the GIL; order is preserved
NATIVE_DECODE = True             # threaded C decoder for mono PCM16 WAVs (fixed-duration windows)
# extract_dataset: feature rows via page-locked staging + helper-thread copies instead of D2H straight into the (pageable)
# result array.  Opt-in (B2A_OUT_STAGING=1): measured on config 5 it takes the extract step from 0.152 s to 0.138 s once
# the two 330 MB staging buffers exist, but allocating them costs more than that on a one-shot run.
OUT_STAGING = os.environ.get("B2A_OUT_STAGING", "0") == "1"


def _make_engine(cfg: B.B2AConfig, device: int):
    """Engine factory (a seam: host-logic tests substitute an oracle-backed stand-in)."""
    return B.Engine(cfg, device)


def _alloc_staging(shape, dtype):
    """Page-locked staging (true async H2D/D2H); plain memory when no CUDA runtime is usable
    (the engine call that follows then fails loudly — this is not a compute fallback)."""
    try:
        pa = B.PinnedArray(shape, dtype)
        return pa.array, pa
    except Exception:  # noqa: BLE001
        return np.empty(shape, dtype), None


def _resolve_devices(devices) -> list:
    if devices is None:
        env = os.environ.get("B2A_DEVICES", "").strip()
        if not env:
            return [0]
        devices = env
    if isinstance(devices, str):
        if devices.lower() == "all":
            n = B.device_count()
            if n <= 0:
                raise RuntimeError("no CUDA device visible: audio_edge_ml_pipeline_b200 has no CPU path")
            return list(range(n))
        return [int(x) for x in devices.split(",") if x.strip() != ""]
    return [int(d) for d in devices]


def _populate_async(arr: np.ndarray, n_threads: int = 4, block: int = 16 << 20) -> None:
    """Fault the pages of a freshly created file mapping in from helper threads (madvise MADV_POPULATE_WRITE, front
    to back in 16 MiB blocks) while the first windows are still being decoded: left to the D2H copies, the
    137 000 first-touch faults of a 560 MB features.npy cost as much as the np.save they replace.  Best effort —
    where the kernel does not know the advice, the copies fault the pages themselves."""
    import ctypes
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        madvise = libc.madvise
        madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        madvise.restype = ctypes.c_int
    except (OSError, AttributeError):
        return
    page = os.sysconf("SC_PAGE_SIZE") if hasattr(os, "sysconf") else 4096
    lo = arr.ctypes.data & ~(page - 1)
    hi = arr.ctypes.data + arr.nbytes
    blocks = [(a, min(block, hi - a)) for a in range(lo, hi, block)]

    def work(k):
        for a, ln in blocks[k::n_threads]:
            if madvise(a, ln, 23) != 0:          # MADV_POPULATE_WRITE (Linux >= 5.14)
                return

    for k in range(min(n_threads, len(blocks))):
        threading.Thread(target=work, args=(k,), daemon=True).start()


class _GpuAudioExtractor(BaseFeatureExtractor):
    feature_type = "deep"
    modality = "audio"
    _kind: int = -1

    sample_rate: int
    duration: Optional[float]

    def _init_common(self, devices) -> None:
        self._devices_arg = devices
        self._engines: dict = {}
        self._staging: dict = {}      # (n_samples, dtype) -> pinned (in, out) batch buffers, reused across calls
        self._lock = threading.Lock()
        self._out_slot = 0
        self._pending: dict = {}      # staging slot -> futures of the copies still moving its rows to the result array
        self._copier = None

    # ---- per-extractor hooks -----------------------------------------------------------
    def _min_samples(self) -> int:
        raise NotImplementedError

    def _fill_config(self, cfg: B.B2AConfig) -> None:
        raise NotImplementedError

    # ---- engines -----------------------------------------------------------------------
    @property
    def devices(self) -> list:
        return _resolve_devices(self._devices_arg)

    def _engine(self, n_samples: int, in_dtype, device: int):
        key = (int(n_samples), np.dtype(in_dtype).str, int(device))
        with self._lock:
            eng = self._engines.get(key)
            if eng is None:
                cfg = B.default_config(self._kind)
                cfg.n_samples = int(n_samples)
                cfg.input_dtype = B.IN_I16 if np.dtype(in_dtype) == np.int16 else B.IN_F32
                cfg.sample_rate = int(self.sample_rate)
                self._fill_config(cfg)
                eng = _make_engine(cfg, device)
                self._engines[key] = eng
            return eng

    def close(self) -> None:
        self._wait_copies()
        if self._copier is not None:
            self._copier.shutdown(wait=True)
            self._copier = None
        for e in self._engines.values():
            e.close()
        self._engines.clear()
        for st in self._staging.values():
            for hnd in st[2:]:
                if hnd is not None:
                    hnd.close()
        self._staging.clear()

    # ---- host-side front end (deep.py:30-61) ---------------------------------------------
    def _prepare(self, sample_path, start_time, end_time) -> np.ndarray:
        audio = wavio.load_segment(sample_path, self.sample_rate, start_time, end_time,
                                   min_samples=self._min_samples())
        if self.duration is not None:
            audio = wavio.pad_or_trim(audio, int(self.duration * self.sample_rate))
        return audio

    # ---- compute ---------------------------------------------------------------------------
    def extract_batch(self, clips: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """(N, n_samples) int16/float32 clips, already padded/trimmed -> (N, rows, T) float32.
        Clips are sharded in contiguous blocks over ``devices`` (no collective)."""
        clips = np.asarray(clips)
        if clips.dtype != np.int16:
            clips = clips.astype(np.float32, copy=False)
        if clips.ndim != 2:
            raise ValueError("clips must be (N, n_samples)")
        n, ns = clips.shape
        devs = self.devices
        engines = [self._engine(ns, clips.dtype, d) for d in devs[:max(1, min(len(devs), n))]]
        rows, frames = engines[0].rows, engines[0].frames
        if out is None:
            out = np.empty((n, rows, frames), dtype=np.float32)
        if len(engines) == 1:
            engines[0].run_host(clips, out)
            return out
        bounds = [n * g // len(engines) for g in range(len(engines) + 1)]
        errs: list = []

        def work(g):
            try:
                a, b = bounds[g], bounds[g + 1]
                if b > a:
                    engines[g].run_host(clips[a:b], out[a:b])
            except Exception as exc:  # noqa: BLE001
                errs.append(exc)

        threads = [threading.Thread(target=work, args=(g,)) for g in range(len(engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out

    def extract_array(self, audio: np.ndarray) -> np.ndarray:
        """One decoded clip (int16 or float32, any length) -> (rows, T): the body of
        ``extract`` after ``librosa.load`` (pad to min_samples, pad/trim to duration)."""
        audio = np.asarray(audio)
        if audio.dtype != np.int16:
            audio = audio.astype(np.float32, copy=False)
        if len(audio) < self._min_samples():
            audio = np.pad(audio, (0, self._min_samples() - len(audio)))
        if self.duration is not None:
            audio = wavio.pad_or_trim(audio, int(self.duration * self.sample_rate))
        return self._extract_one(audio)

    def extract(self, sample_path: Path, start_time: Optional[float] = None,
                end_time: Optional[float] = None, **_kwargs) -> np.ndarray:
        return self._extract_one(self._prepare(sample_path, start_time, end_time))

    def _extract_one(self, audio: np.ndarray) -> np.ndarray:
        """One prepared clip.  With ``duration=None`` every file has its own length: the clip goes through the
        ragged entry point of an engine sized to the next power of two, so a loop over files (the reference's
        ``extract_dataset``) touches a handful of engines instead of one per distinct length."""
        if self.duration is None and self._kind != B.KIND_CQT:
            cap = 1 << int(np.ceil(np.log2(max(len(audio), 2))))
            return self._engine(cap, audio.dtype, self.devices[0]).run_host_ragged([audio])[0]
        return self.extract_batch(audio[None, :])[0]

    # ---- dataset extraction ------------------------------------------------------------------
    def _run_sharded(self, n: int, fn) -> None:
        """fn(device_slot, device, a, b): clips [a, b) of a batch of n on one device; contiguous blocks,
        one host thread per device (no collective, the gather is each device's D2H)."""
        devs = self.devices[:max(1, min(len(self.devices), n))]
        if len(devs) == 1:
            fn(0, devs[0], 0, n)
            return
        bounds = [n * g // len(devs) for g in range(len(devs) + 1)]
        errs: list = []

        def work(g):
            try:
                if bounds[g + 1] > bounds[g]:
                    fn(g, devs[g], bounds[g], bounds[g + 1])
            except Exception as exc:  # noqa: BLE001
                errs.append(exc)

        threads = [threading.Thread(target=work, args=(g,)) for g in range(len(devs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]

    def _copy_async(self, slot: int, dst: np.ndarray, src: np.ndarray, parts: int = 4) -> None:
        from concurrent.futures import ThreadPoolExecutor
        if self._copier is None:
            self._copier = ThreadPoolExecutor(max_workers=parts)
        n = len(src)
        bounds = [n * k // parts for k in range(parts + 1)]
        self._pending[slot] = [self._copier.submit(np.copyto, dst[a:b], src[a:b])
                               for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

    def _wait_copies(self, slot: Optional[int] = None) -> None:
        for k in ([slot] if slot is not None else list(self._pending)):
            for fut in self._pending.pop(k, []):
                fut.result()

    def _stage(self, key, shape_tail, dtype):
        """Pinned staging rows, reused across windows and calls: key -> array (HOST_BATCH_CLIPS or fewer rows)."""
        ent = self._staging.get(key)
        if ent is None:
            arr, hnd = _alloc_staging((key[1],) + tuple(shape_tail), dtype)
            ent = self._staging[key] = (arr, None, hnd, None)
        return ent[0]

    def _resample_group(self, paths, offs, durs, rate: int, n: int, mono16: bool, dst: np.ndarray) -> None:
        """Files at `rate` != sample_rate (deep.py:44-50: librosa.load resamples them): native decode at the
        file rate -> pinned staging -> device resampler + pad / trim to n samples + kernels -> dst rows."""
        up, down, half, _poly = B.resampler_design(rate, self.sample_rate)
        need_in = ((n - 1) * down + half) // up + 1          # input frames the first n outputs touch
        dt = np.int16 if mono16 else np.float32
        per = max(1, min(HOST_BATCH_CLIPS, (256 << 20) // (need_in * np.dtype(dt).itemsize)))
        raw = self._stage(("raw", per, need_in, np.dtype(dt).str), (need_in,), dt)
        for a in range(0, len(paths), per):
            b = min(len(paths), a + per)
            _r, n_out, status = B.decode_wav_batch(paths[a:b], need_in, raw, offs[a:b], durs[a:b], DECODE_WORKERS)
            if status.any():
                raise RuntimeError(f"decode failed for {paths[a + int(np.flatnonzero(status)[0])]} (status {status.max()})")

            def one(slot, dev, lo, hi):
                eng = self._engine(n, np.float32, dev)
                eng.run_host_resampled(B.get_resampler(rate, self.sample_rate, dev), raw[lo:hi], n_out[lo:hi],
                                       dst[a + lo:a + hi])
            self._run_sharded(b - a, one)

    def _float_group(self, paths, offs, durs, n: int, dst: np.ndarray) -> None:
        """Files at the right rate that are not mono PCM16 (stereo, 8/24/32-bit, float): native decode to
        float32 with the channel mean, then the float32-input engine."""
        per = max(1, min(HOST_BATCH_CLIPS, (256 << 20) // (n * 4)))
        raw = self._stage(("f32", per, n, "f4"), (n,), np.float32)
        for a in range(0, len(paths), per):
            b = min(len(paths), a + per)
            _r, _n, status = B.decode_wav_batch(paths[a:b], n, raw, offs[a:b], durs[a:b], DECODE_WORKERS)
            if status.any():
                raise RuntimeError(f"decode failed for {paths[a + int(np.flatnonzero(status)[0])]} (status {status.max()})")
            self._run_sharded(b - a, lambda slot, dev, lo, hi: self._engine(n, np.float32, dev)
                              .run_host(raw[lo:hi], dst[a + lo:a + hi]))

    def _window_native(self, items, dst: np.ndarray) -> list:
        """One window of loader items, fixed duration, through the native front end.  Features of item i go
        to dst[i]; returns a per-item list: None = done, an Exception = skip this sample, "py" = not a
        file the native decoder covers (the Python path decides: other containers raise there)."""
        n = int(self.duration * self.sample_rate)
        paths = [str(p_) for p_, _l, _m in items]
        metas = [m for _p, _l, m in items]
        offs = np.array([float(m.get("start_time") or 0.0) for m in metas])
        durs = np.array([max(float(m["end_time"]) - o, 0.1) if m.get("end_time") is not None else -1.0
                         for m, o in zip(metas, offs)])
        info = B.probe_wav_batch(paths, DECODE_WORKERS)
        res: list = [None] * len(items)
        groups: dict = {}
        for i in range(len(items)):
            if info["status"][i] != B.DEC_OK:
                res[i] = "py"
                continue
            mono16 = bool(info["format_tag"][i] == 1 and info["bits"][i] == 16 and info["channels"][i] == 1)
            groups.setdefault((int(info["rate"][i]), mono16), []).append(i)

        def run(key, idxs):
            """A group, or on failure its halves, down to single clips (skip granularity = one sample)."""
            rate, mono16 = key
            sub = lambda arr: [arr[i] for i in idxs] if isinstance(arr, list) else arr[idxs]  # noqa: E731
            contiguous = idxs[-1] - idxs[0] + 1 == len(idxs)
            out = dst[idxs[0]:idxs[-1] + 1] if contiguous else np.empty((len(idxs),) + dst.shape[1:], np.float32)
            try:
                if rate == self.sample_rate and mono16:
                    a_in = self._stage(("i16", HOST_BATCH_CLIPS, n, "i2"), (n,), np.int16)
                    status = B.decode_wav_pcm16_batch(sub(paths), self.sample_rate, n, a_in, sub(offs), sub(durs),
                                                      DECODE_WORKERS)
                    if status.any():
                        raise RuntimeError(f"decode failed (status {int(status.max())})")
                    if contiguous and OUT_STAGING:
                        # the engine's D2H lands in page-locked staging at link speed (into the pageable result
                        # array it runs at the driver's single-threaded bounce-buffer rate) and helper threads move
                        # it on while the next window is decoded and shipped; two staging buffers alternate
                        self._out_slot ^= 1
                        o_pin = self._stage(("o32", HOST_BATCH_CLIPS, self._out_slot) + tuple(out.shape[1:]), out.shape[1:],
                                            np.float32)[:len(idxs)]
                        self._wait_copies(self._out_slot)
                        self.extract_batch(a_in[:len(idxs)], o_pin)
                        self._copy_async(self._out_slot, out, o_pin)
                    else:
                        self.extract_batch(a_in[:len(idxs)], out)
                elif rate == self.sample_rate:
                    self._float_group(sub(paths), sub(offs), sub(durs), n, out)
                else:
                    self._resample_group(sub(paths), sub(offs), sub(durs), rate, n, mono16, out)
                if not contiguous:
                    dst[idxs] = out
            except Exception as exc:  # noqa: BLE001
                if len(idxs) == 1:
                    res[idxs[0]] = exc
                else:
                    h = len(idxs) // 2
                    run(key, idxs[:h])
                    run(key, idxs[h:])

        for key, idxs in groups.items():
            run(key, idxs)
        return res

    def extract_dataset(self, loader, max_samples: Optional[int] = None, features_out=None) -> FeatureSet:
        """``features_out``: path of the ``features.npy`` this run will be saved as (FeaturePipeline passes it).
        With a fixed duration and a sized loader the feature rows are then written straight into that file
        (a memory-mapped NPY v1 array, the same bytes ``np.save`` produces) instead of into an anonymous
        array that ``save`` copies once more; if any sample is skipped the run falls back to the in-memory
        array and ``save`` writes the file as usual."""
        feats: list = []             # arrays (k, rows, T) (or per-clip (1, rows, T_i) when ragged), loader order
        labels: list = []
        metas: list = []
        label_to_idx: dict = {}
        ragged = self.duration is None and self._kind != B.KIND_CQT
        n_fixed = int(self.duration * self.sample_rate) if self.duration is not None else 0
        native = NATIVE_DECODE and self.duration is not None and n_fixed >= self._min_samples()
        # fixed duration + a sized loader: windows land in one preallocated array (no concatenation) for as
        # long as every window goes through the native front end; whatever comes after is appended (`tail`)
        final = {"arr": None, "pos": 0, "open": True}
        tail: list = []

        def book(item) -> None:
            _p, label, meta = item
            metas.append(meta)
            if label is not None:
                if label not in label_to_idx:
                    label_to_idx[label] = len(label_to_idx)
                labels.append(label_to_idx[label])

        def window():
            buf = []
            for i, item in enumerate(loader):
                if max_samples is not None and i >= max_samples:
                    break
                buf.append(item)
                if len(buf) >= HOST_BATCH_CLIPS:
                    yield buf
                    buf = []
            if buf:
                yield buf

        def decode(item):
            sample_path, _label, meta = item
            try:
                return self._prepare(sample_path, meta.get("start_time"), meta.get("end_time"))
            except Exception as exc:  # noqa: BLE001 — reference semantics: warn and skip (base.py:204-206)
                return exc

        def python_path(items, pool) -> list:
            """Python decoder + one engine call per (length, dtype) group (one ragged call per dtype when
            duration is None).  Returns per item a (rows, T) array or an Exception."""
            decoded = list(pool.map(decode, items)) if DECODE_WORKERS > 1 else [decode(it) for it in items]
            out: list = list(decoded)
            groups: dict = {}
            for i, a in enumerate(decoded):
                if not isinstance(a, Exception):
                    # ragged: clips are bucketed by the next power of two of their length, so one very long file
                    # sizes only its own bucket's engine (and scratch), not the whole window's
                    key = int(np.ceil(np.log2(max(len(a), 2)))) if ragged else len(a)
                    groups.setdefault((key, a.dtype.str), []).append(i)

            def run(idxs):
                try:
                    clips = [decoded[i] for i in idxs]
                    if ragged:
                        cap = 1 << int(np.ceil(np.log2(max(len(c) for c in clips))))   # few engines: power-of-two maxima
                        got = self._engine(cap, clips[0].dtype, self.devices[0]).run_host_ragged(clips)
                    else:
                        got = self.extract_batch(np.stack(clips))
                    for k, i in enumerate(idxs):
                        out[i] = got[k]
                except Exception as exc:  # noqa: BLE001 — a failing batch is retried in halves: the skip
                    if len(idxs) == 1:    # granularity stays one sample, as in the reference loop
                        out[idxs[0]] = exc
                    else:
                        run(idxs[:len(idxs) // 2])
                        run(idxs[len(idxs) // 2:])

            for idxs in groups.values():
                run(idxs)
            return out

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=DECODE_WORKERS) as pool:
            for items in window():
                res: list = ["py"] * len(items)
                dst, in_place = None, False
                if native:
                    try:
                        eng = self._engine(n_fixed, np.int16, self.devices[0])
                        shape = (eng.rows, eng.frames)
                        if final["arr"] is None and final["open"] and hasattr(loader, "__len__"):
                            cap = len(loader) if max_samples is None else min(len(loader), max_samples)
                            if features_out is not None and cap > 0:
                                Path(features_out).parent.mkdir(parents=True, exist_ok=True)
                                final["arr"] = np.lib.format.open_memmap(str(features_out), mode="w+", dtype=np.float32,
                                                                         shape=(cap,) + shape)
                                _populate_async(final["arr"])
                            else:
                                final["arr"] = np.empty((cap,) + shape, dtype=np.float32)
                        pos = final["pos"]
                        in_place = final["open"] and final["arr"] is not None and pos + len(items) <= len(final["arr"])
                        if in_place:
                            dst = final["arr"][pos:pos + len(items)]
                        else:
                            final["open"] = False
                            dst = np.empty((len(items),) + shape, dtype=np.float32)
                        res = self._window_native(items, dst)
                    except Exception as exc:  # noqa: BLE001 — e.g. engine creation failed: per-sample policy below
                        logger.debug("native front end unavailable: %s", exc)
                        res, dst, in_place = ["py"] * len(items), None, False
                if any(r is not None for r in res):
                    self._wait_copies()                  # rows are about to be patched / compacted below
                todo = [i for i, r in enumerate(res) if isinstance(r, str)]
                if todo:
                    got = python_path([items[i] for i in todo], pool)
                    for i, g in zip(todo, got):
                        res[i] = g
                # commit in loader order
                ok = [i for i, r in enumerate(res) if not isinstance(r, Exception)]
                for i, r in enumerate(res):
                    if isinstance(r, Exception):
                        logger.warning("Skipping %s: %s", items[i][0], r)
                if dst is not None:
                    for i in ok:
                        if res[i] is not None:
                            dst[i] = res[i]                      # rows the Python path produced join the window's array
                    if len(ok) != len(items):
                        dst[:len(ok)] = dst[ok]                  # close the gaps the skipped samples left
                    if in_place:
                        final["pos"] += len(ok)                  # the next window starts right behind these rows
                    elif ok:
                        tail.append(dst[:len(ok)])
                else:
                    final["open"] = False                        # rows from here on are appended, not placed
                    tail.extend(np.asarray(res[i])[None] for i in ok)
                for i in ok:
                    book(items[i])
        self._wait_copies()                                  # every staged row has reached the result array
        if not metas:
            raise RuntimeError("No features were successfully extracted.")
        mapped = isinstance(final["arr"], np.memmap)
        if mapped and not tail and final["pos"] == len(final["arr"]):
            fs = assemble_feature_set(self, final["arr"], labels, metas, label_to_idx)
            fs.features_file = Path(features_out)      # save() finds the rows already in place
            return fs
        parts = ([final["arr"][:final["pos"]]] if final["pos"] else []) + tail
        features = np.array(parts[0]) if (len(parts) == 1 and mapped) else \
            (parts[0] if len(parts) == 1 else np.concatenate(parts))            # ragged shapes -> ValueError
        if mapped:                                     # samples were skipped: the file's shape is wrong, drop it
            final["arr"] = None
            Path(features_out).unlink(missing_ok=True)
        return assemble_feature_set(self, features, labels, metas, label_to_idx)


@register
class AudioMelSpectrogram(_GpuAudioExtractor):
    """Log-mel spectrogram in [0, 1], shape ``(n_mels, T)`` — deep.py:75-134."""

    name = "audio_mel_spec"
    _kind = B.KIND_MEL

    def __init__(self, sample_rate: int = 16000, n_mels: int = 40, n_fft: int = 512,
                 hop_length: int = 160, duration: Optional[float] = None, *,
                 pad_mode: str = "constant", devices: Optional[Sequence[int]] = None) -> None:
        self.sample_rate = sample_rate
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.duration = duration
        self.pad_mode = pad_mode
        self._init_common(devices)

    def _min_samples(self) -> int:
        return self.n_fft                      # deep.py:119-120

    def _fill_config(self, cfg) -> None:
        cfg.n_fft, cfg.hop_length, cfg.n_mels = int(self.n_fft), int(self.hop_length), int(self.n_mels)
        cfg.pad_mode = {"constant": B.PAD_CONSTANT, "reflect": B.PAD_REFLECT}[self.pad_mode]


@register
class AudioCQT(_GpuAudioExtractor):
    """Log-magnitude constant-Q transform in [0, 1], shape ``(n_bins, T)`` — deep.py:196-260."""

    name = "audio_cqt"
    _kind = B.KIND_CQT

    def __init__(self, sample_rate: int = 22050, hop_length: int = 512, n_bins: int = 84,
                 bins_per_octave: int = 12, fmin: Optional[float] = None,
                 duration: Optional[float] = None, *, devices: Optional[Sequence[int]] = None) -> None:
        self.sample_rate = sample_rate
        self.hop_length = hop_length
        self.n_bins = n_bins
        self.bins_per_octave = bins_per_octave
        self.fmin = fmin
        self.duration = duration
        self._init_common(devices)

    def _min_samples(self) -> int:
        return self.hop_length * 2             # deep.py:242-243

    def _fill_config(self, cfg) -> None:
        cfg.hop_length, cfg.n_bins = int(self.hop_length), int(self.n_bins)
        cfg.bins_per_octave = int(self.bins_per_octave)
        cfg.fmin = float(self.fmin) if self.fmin is not None else 0.0


@register
class AudioMFCCSequence(_GpuAudioExtractor):
    """Per-coefficient z-scored MFCC sequence, shape ``(n_mfcc, T)`` — deep.py:268-328."""

    name = "audio_mfcc_seq"
    _kind = B.KIND_MFCC

    def __init__(self, sample_rate: int = 22050, n_mfcc: int = 40, n_fft: int = 1024,
                 hop_length: int = 512, duration: Optional[float] = None, *, n_mels: int = 128,
                 pad_mode: str = "constant", devices: Optional[Sequence[int]] = None) -> None:
        self.sample_rate = sample_rate
        self.n_mfcc = n_mfcc
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.duration = duration
        self.n_mels = n_mels
        self.pad_mode = pad_mode
        self._init_common(devices)

    def _min_samples(self) -> int:
        return self.n_fft                      # deep.py:311-312

    def _fill_config(self, cfg) -> None:
        cfg.n_fft, cfg.hop_length = int(self.n_fft), int(self.hop_length)
        cfg.n_mels, cfg.n_mfcc = int(self.n_mels), int(self.n_mfcc)
        cfg.pad_mode = {"constant": B.PAD_CONSTANT, "reflect": B.PAD_REFLECT}[self.pad_mode]


_CLASSICAL_FEATURES = ["mfcc", "delta_mfcc", "delta2_mfcc", "spectral_centroid", "spectral_rolloff",
                       "spectral_bandwidth", "spectral_contrast", "spectral_flatness", "chroma", "zcr", "rms",
                       "tonnetz"]                                     # classical.py:61-74
_CLASSICAL_RAW_DIMS = {"spectral_centroid": 1, "spectral_rolloff": 1, "spectral_bandwidth": 1,
                       "spectral_contrast": 7, "spectral_flatness": 1, "chroma": 12, "zcr": 1, "rms": 1,
                       "tonnetz": 6}                                   # classical.py:78-88
_CLASSICAL_AGGREGATIONS = ["mean", "std"]


@register
class AudioClassicalExtractor(_GpuAudioExtractor):
    """Flat classical feature vector (MFCC + deltas, spectral shape, contrast, chroma, zcr, rms, tonnetz; mean / std
    over time) — classical.py:96-355.  The device computes every group with both aggregations in one pass over
    the clip (``B2A_KIND_CLASSICAL``); ``features`` / ``aggregations`` select columns of that vector in the
    reference's canonical order.  ``duration`` is an addition (the reference extractor always takes the whole
    segment): with it set, clips are padded / trimmed to a fixed length so that a dataset runs as whole batches."""

    name = "audio_classical"
    feature_type = "classical"
    _kind = B.KIND_CLASSICAL

    def __init__(self, sample_rate: int = 22050, n_mfcc: int = 40, n_mels: int = 128, n_fft: int = 1024,
                 hop_length: int = 512, min_duration: float = 0.1, features: Optional[list] = None,
                 aggregations: Optional[list] = None, *, duration: Optional[float] = None,
                 devices: Optional[Sequence[int]] = None) -> None:
        self.sample_rate = sample_rate
        self.n_mfcc = n_mfcc
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.min_duration = min_duration
        self.duration = duration
        if features is None:
            self.features = list(_CLASSICAL_FEATURES)
        else:
            unknown = set(features) - set(_CLASSICAL_FEATURES)
            if unknown:
                raise ValueError(f"Unknown feature group(s): {sorted(unknown)}. Valid keys: {_CLASSICAL_FEATURES}")
            self.features = [k for k in _CLASSICAL_FEATURES if k in set(features)]     # canonical order
        if aggregations is None:
            self.aggregations = list(_CLASSICAL_AGGREGATIONS)
        else:
            unknown = set(aggregations) - set(_CLASSICAL_AGGREGATIONS)
            if unknown:
                raise ValueError(f"Unknown aggregation(s): {sorted(unknown)}. Valid values: {_CLASSICAL_AGGREGATIONS}")
            if not aggregations:
                raise ValueError("aggregations must contain at least one value.")
            self.aggregations = [a for a in _CLASSICAL_AGGREGATIONS if a in set(aggregations)]
        self._columns = self._column_index()
        self._init_common(devices)

    @property
    def feature_dim(self) -> int:
        """classical.py:197-207."""
        return len(self._columns)

    def _column_index(self) -> np.ndarray:
        """Columns of the device vector (every group: [means..., stds...]) that this configuration keeps."""
        cols, pos = [], 0
        for key in _CLASSICAL_FEATURES:
            dim = self.n_mfcc if key in ("mfcc", "delta_mfcc", "delta2_mfcc") else _CLASSICAL_RAW_DIMS[key]
            if key in self.features:
                for a, agg in enumerate(_CLASSICAL_AGGREGATIONS):
                    if agg in self.aggregations:
                        cols.extend(range(pos + a * dim, pos + (a + 1) * dim))
            pos += 2 * dim
        return np.asarray(cols, dtype=np.int64)

    def _min_samples(self) -> int:
        # classical.py:262-270: one STFT frame and nine MFCC frames for the width-9 delta
        return max(int(self.min_duration * self.sample_rate), self.n_fft, 8 * self.hop_length)

    def _fill_config(self, cfg) -> None:
        cfg.n_fft, cfg.hop_length = int(self.n_fft), int(self.hop_length)
        cfg.n_mels, cfg.n_mfcc = int(self.n_mels), int(self.n_mfcc)

    def _select(self, full: np.ndarray) -> np.ndarray:
        full = full.reshape(full.shape[0], -1)
        if len(self._columns) == full.shape[1]:
            return full
        return np.ascontiguousarray(full[:, self._columns])

    def extract_batch(self, clips: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """(N, n_samples) equal-length clips -> (N, feature_dim) float32."""
        if getattr(self, "_raw_rows", False):       # inside extract_dataset: rows stay device vectors until the end
            return super().extract_batch(clips, out)
        got = self._select(super().extract_batch(clips))
        if out is not None:
            out[...] = got
            return out
        return got

    def _extract_one(self, audio: np.ndarray) -> np.ndarray:
        if self.duration is None:                   # ragged engine: the raw (rows, 1) device vector comes back
            return self._select(super()._extract_one(audio)[None])[0]
        return super()._extract_one(audio)          # fixed length: extract_batch above has already selected

    def extract_dataset(self, loader, max_samples: Optional[int] = None, features_out=None) -> FeatureSet:
        # (features_out is not used: the rows written during the run are full device vectors, the file holds the
        #  selected columns; save() writes it)
        self._raw_rows = True
        try:
            fs = super().extract_dataset(loader, max_samples)
        finally:
            self._raw_rows = False
        fs.features = self._select(fs.features)    # (N, rows, 1) device vectors -> (N, feature_dim)
        return fs

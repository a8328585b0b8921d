"""ctypes binding of ``libb2a.so`` (C ABI in ``include/b2a.h``).

This is the ONLY compute path of the package: there is no CPU fallback.  Importing this module
does not need a GPU (so symbol/ABI checks run anywhere); creating an :class:`Engine` does, and
fails loudly when the library or a CUDA device is missing.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

import numpy as np

PKG = Path(__file__).resolve().parent
# B2A_LIBRARY: another build of the same library (build.build_variant: numerics A/B experiments)
LIB_PATH = Path(os.environ["B2A_LIBRARY"]) if os.environ.get("B2A_LIBRARY") else PKG / "libb2a.so"

KIND_MEL, KIND_MFCC, KIND_CQT, KIND_CLASSICAL = 0, 1, 2, 3
IN_I16, IN_F32 = 0, 1
PAD_CONSTANT, PAD_REFLECT = 0, 1
TABLE_WINDOW, TABLE_MEL_DENSE, TABLE_DCT, TABLE_DECIM_TAPS, TABLE_CQT_LENGTHS, TABLE_CQT_BASIS = range(6)
TABLE_CHROMA, TABLE_TONNETZ, TABLE_CONTRAST_BANDS = 6, 7, 8

EXPORTED_SYMBOLS = [
    "b2a_default_config", "b2a_create", "b2a_destroy", "b2a_out_shape", "b2a_run_device",
    "b2a_run_host", "b2a_run_host_copy_only", "b2a_run_host_ragged", "b2a_run_device_ragged", "b2a_last_launch_count", "b2a_alloc_pinned", "b2a_free_pinned",
    "b2a_decode_wav_pcm16_batch", "b2a_probe_wav_batch", "b2a_decode_wav_batch", "b2a_get_table", "b2a_cqt_geometry", "b2a_last_error", "b2a_abi_version", "b2a_device_count",
    "b2a_resampler_create", "b2a_resampler_destroy", "b2a_resampler_out_len", "b2a_resampler_geometry",
    "b2a_resampler_design", "b2a_resampler_run_host", "b2a_resampler_run_device", "b2a_resampler_last_error",
    "b2a_resampler_run_device_batch", "b2a_resampler_rates", "b2a_run_host_resampled",
    "b2a_augment_host", "b2a_augment_device", "b2a_classical_tunings",
    "b2a_time_stretch_host", "b2a_pitch_shift_host",
]


class B2AError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb2a error {code}: {msg}")
        self.code = code


class B2AConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("input_dtype", C.c_int32), ("sample_rate", C.c_int32),
        ("n_samples", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32),
        ("n_mels", C.c_int32), ("n_mfcc", C.c_int32), ("n_bins", C.c_int32),
        ("bins_per_octave", C.c_int32), ("fmin", C.c_double), ("pad_mode", C.c_int32),
        ("top_db", C.c_float), ("reserved", C.c_int32 * 8),
    ]


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen libb2a.so and declare prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m audio_edge_ml_pipeline_b200.build` "
            "(there is no CPU fallback for this package)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.b2a_default_config.argtypes = [i32, C.POINTER(B2AConfig)]
    lib.b2a_create.argtypes = [C.POINTER(B2AConfig), i32, C.POINTER(vp)]
    lib.b2a_destroy.argtypes = [vp]
    lib.b2a_out_shape.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    lib.b2a_run_device.argtypes = [vp, vp, i64, vp, vp]
    lib.b2a_run_host.argtypes = [vp, vp, i64, vp]
    lib.b2a_run_host_copy_only.argtypes = [vp, vp, i64, vp]
    lib.b2a_run_host_copy_only.restype = C.c_int
    lib.b2a_run_host_ragged.argtypes = [vp, vp, i64, vp, vp, vp, i64, vp, i64]
    lib.b2a_run_device_ragged.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    lib.b2a_decode_wav_pcm16_batch.argtypes = [vp, i64, i32, vp, vp, i32, vp, vp, i32]
    lib.b2a_decode_wav_pcm16_batch.restype = C.c_int
    lib.b2a_probe_wav_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i32]
    lib.b2a_probe_wav_batch.restype = C.c_int
    lib.b2a_decode_wav_batch.argtypes = [vp, i64, vp, vp, i64, i32, vp, i64, vp, vp, vp, i32]
    lib.b2a_decode_wav_batch.restype = C.c_int
    lib.b2a_resampler_run_device_batch.argtypes = [vp, vp, i32, i64, i64, vp, vp, i64, i64, vp]
    lib.b2a_resampler_run_device_batch.restype = C.c_int
    lib.b2a_resampler_rates.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.b2a_resampler_rates.restype = C.c_int
    lib.b2a_run_host_resampled.argtypes = [vp, vp, vp, i32, i64, vp, i64, vp]
    lib.b2a_run_host_resampled.restype = C.c_int
    lib.b2a_last_launch_count.argtypes = [vp]
    lib.b2a_last_launch_count.restype = i64
    lib.b2a_classical_tunings.argtypes = [vp, C.POINTER(C.c_float), i64]
    lib.b2a_classical_tunings.restype = C.c_int
    lib.b2a_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.b2a_free_pinned.argtypes = [vp]
    lib.b2a_get_table.argtypes = [vp, i32, vp, C.POINTER(i64)]
    lib.b2a_cqt_geometry.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp, vp, vp]
    lib.b2a_last_error.restype = C.c_char_p
    lib.b2a_resampler_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    lib.b2a_resampler_destroy.argtypes = [vp]
    lib.b2a_resampler_out_len.argtypes = [vp, i64]
    lib.b2a_resampler_out_len.restype = i64
    lib.b2a_resampler_geometry.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp]
    lib.b2a_resampler_design.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp, i64]
    lib.b2a_resampler_run_host.argtypes = [vp, vp, i32, i64, vp]
    lib.b2a_resampler_run_device.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.b2a_resampler_last_error.restype = C.c_char_p
    for name in ("b2a_resampler_create", "b2a_resampler_destroy", "b2a_resampler_geometry", "b2a_resampler_design",
                 "b2a_resampler_run_host", "b2a_resampler_run_device"):
        getattr(lib, name).restype = C.c_int
    for name in ("b2a_default_config", "b2a_create", "b2a_destroy", "b2a_out_shape", "b2a_run_device",
                 "b2a_run_host", "b2a_run_host_ragged", "b2a_run_device_ragged", "b2a_alloc_pinned", "b2a_free_pinned", "b2a_get_table",
                 "b2a_cqt_geometry", "b2a_abi_version", "b2a_device_count"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise B2AError(rc, load_library().b2a_last_error().decode(errors="replace"))


def default_config(kind: int) -> B2AConfig:
    cfg = B2AConfig()
    _check(load_library().b2a_default_config(kind, C.byref(cfg)))
    return cfg


def device_count() -> int:
    return int(load_library().b2a_device_count())


DEC_OK, DEC_EIO, DEC_EFORMAT, DEC_EUNSUPPORTED, DEC_ERATE = range(5)


def decode_wav_pcm16_batch(paths, sample_rate: int, n_samples: int, out: np.ndarray, offsets=None,
                           durations=None, n_threads: int = 0) -> np.ndarray:
    """Native threaded decode of mono PCM16 WAV files into ``out[:len(paths)]`` (int16,
    (>=N, n_samples), C-contiguous).  Returns the per-file B2A_DEC_* status array."""
    lib = load_library()
    n = len(paths)
    status = np.zeros(n, dtype=np.int32)
    if n == 0:
        return status
    assert out.dtype == np.int16 and out.flags.c_contiguous and out.shape[1] == n_samples and out.shape[0] >= n
    arr = (C.c_char_p * n)(*[str(p).encode() for p in paths])
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    dur = None if durations is None else np.ascontiguousarray(durations, dtype=np.float64)
    _check(lib.b2a_decode_wav_pcm16_batch(C.cast(arr, C.c_void_p), n, int(sample_rate),
                                          None if off is None else off.ctypes.data,
                                          None if dur is None else dur.ctypes.data,
                                          int(n_samples), out.ctypes.data, status.ctypes.data, int(n_threads)))
    return status


def _path_array(paths):
    n = len(paths)
    return (C.c_char_p * n)(*[str(p).encode() for p in paths])


def probe_wav_batch(paths, n_threads: int = 0) -> dict:
    """Header probe of many RIFF/WAVE files (native, threaded): arrays rate / channels / bits / format_tag /
    n_frames / status (B2A_DEC_*; non-WAV files report DEC_EFORMAT)."""
    lib = load_library()
    n = len(paths)
    out = {k: np.zeros(n, dtype=np.int32) for k in ("rate", "channels", "bits", "format_tag", "status")}
    out["n_frames"] = np.zeros(n, dtype=np.int64)
    if n:
        _check(lib.b2a_probe_wav_batch(C.cast(_path_array(paths), C.c_void_p), n, out["rate"].ctypes.data,
                                       out["channels"].ctypes.data, out["bits"].ctypes.data,
                                       out["format_tag"].ctypes.data, out["n_frames"].ctypes.data,
                                       out["status"].ctypes.data, int(n_threads)))
    return out


def _pack_rows(rows):
    lens = np.array([len(r) for r in rows], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.int64) if len(rows) else np.zeros(0, np.int64)
    src = np.concatenate([np.asarray(r, dtype=np.float32) for r in rows]) if len(rows) else np.zeros(0, np.float32)
    return src, off, lens


def time_stretch_rows(rows, rates, device: int = 0) -> list:
    """librosa.effects.time_stretch on the GPU for a batch of float32 rows of any lengths (augment.py:105-110):
    row i comes back with int(round(len / rate_i)) samples.  No CPU path."""
    lib = load_library()
    if len(rows) == 0:
        return []
    src, off, lens = _pack_rows(rows)
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    if not (rates > 0).all():
        raise ValueError("rate must be a positive number")          # librosa.effects.time_stretch's ParameterError
    out_len = np.array([int(round(int(n) / float(r))) for n, r in zip(lens, rates)], dtype=np.int32)
    out_off = np.concatenate([[0], np.cumsum(out_len.astype(np.int64))[:-1]]).astype(np.int64)
    out = np.empty(int(out_len.sum()), dtype=np.float32)
    fn = lib.b2a_time_stretch_host
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    fn.argtypes = [i32, vp, i64, vp, vp, vp, i64, vp, i64, vp, vp]
    fn.restype = C.c_int
    _check(fn(int(device), src.ctypes.data, src.size, off.ctypes.data, lens.ctypes.data, rates.ctypes.data, len(rows),
              out.ctypes.data, out.size, out_off.ctypes.data, out_len.ctypes.data))
    return [out[o:o + n] for o, n in zip(out_off, out_len)]


def pitch_shift_rows(rows, sr: int, n_steps, device: int = 0) -> list:
    """librosa.effects.pitch_shift on the GPU (augment.py:113-118): time_stretch by 2 ** (-n_steps / 12), resample back
    by that ratio, crop / zero-pad to the input length.  No CPU path."""
    lib = load_library()
    if len(rows) == 0:
        return []
    src, off, lens = _pack_rows(rows)
    rates = np.array([2.0 ** (-float(s) / 12) for s in n_steps], dtype=np.float64)
    ratios = np.array([float(sr) / (float(sr) / r) for r in rates], dtype=np.float64)     # target_sr / orig_sr (librosa.resample)
    mid = np.array([int(round(int(n) / float(r))) for n, r in zip(lens, rates)], dtype=np.int32)
    out = np.empty(int(lens.sum()), dtype=np.float32)
    fn = lib.b2a_pitch_shift_host
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    fn.argtypes = [i32, vp, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp]
    fn.restype = C.c_int
    _check(fn(int(device), src.ctypes.data, src.size, off.ctypes.data, lens.ctypes.data, rates.ctypes.data,
              ratios.ctypes.data, mid.ctypes.data, len(rows), out.ctypes.data, out.size, off.ctypes.data))
    return [out[o:o + n] for o, n in zip(off, lens)]


def decode_wav_batch(paths, max_frames: int, out: np.ndarray, offsets=None, durations=None, n_threads: int = 0):
    """Native threaded decode of RIFF/WAVE files AT THEIR OWN RATE into ``out[:len(paths), :max_frames]``
    (int16: mono PCM16 only; float32: every supported PCM / float format, channel mean).  Rows are zero-filled
    past the decoded frames.  Returns (rate, n_out, status) arrays."""
    lib = load_library()
    n = len(paths)
    rate, n_out, status = (np.zeros(n, dtype=np.int32) for _ in range(3))
    if n == 0:
        return rate, n_out, status
    assert out.dtype in (np.int16, np.float32) and out.ndim == 2 and out.strides[1] == out.itemsize
    assert out.shape[0] >= n and out.shape[1] >= max_frames
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    dur = None if durations is None else np.ascontiguousarray(durations, dtype=np.float64)
    _check(lib.b2a_decode_wav_batch(C.cast(_path_array(paths), C.c_void_p), n,
                                    None if off is None else off.ctypes.data, None if dur is None else dur.ctypes.data,
                                    int(max_frames), IN_I16 if out.dtype == np.int16 else IN_F32, out.ctypes.data,
                                    out.strides[0] // out.itemsize, rate.ctypes.data, n_out.ctypes.data,
                                    status.ctypes.data, int(n_threads)))
    return rate, n_out, status


class PinnedArray:
    """numpy view over page-locked host memory owned by the library."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        _check(self._lib.b2a_alloc_pinned(nbytes, C.byref(p)))
        self._ptr = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self) -> None:
        if self._ptr is not None:
            self.array = None
            self._lib.b2a_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One (device, configuration) handle of libb2a."""

    def __init__(self, cfg: B2AConfig, device: int = 0):
        self._lib = load_library()
        self.cfg = cfg
        self.device = device
        h = C.c_void_p()
        _check(self._lib.b2a_create(C.byref(cfg), device, C.byref(h)))
        self._h = h
        rows, frames = C.c_int32(), C.c_int32()
        _check(self._lib.b2a_out_shape(h, C.byref(rows), C.byref(frames)))
        self.rows, self.frames = rows.value, frames.value
        self.in_dtype = np.int16 if cfg.input_dtype == IN_I16 else np.float32

    # -- lifecycle ------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.b2a_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- compute --------------------------------------------------------------------------
    def run_host(self, clips: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """clips: (N, n_samples) int16/float32 host array -> (N, rows, frames) float32."""
        clips = np.ascontiguousarray(clips, dtype=self.in_dtype)
        if clips.ndim != 2 or clips.shape[1] != self.cfg.n_samples:
            raise ValueError(f"clips must be (N, {self.cfg.n_samples}), got {clips.shape}")
        n = clips.shape[0]
        if out is None:
            out = np.empty((n, self.rows, self.frames), dtype=np.float32)
        elif out.shape != (n, self.rows, self.frames) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous float32 (N, rows, frames)")
        _check(self._lib.b2a_run_host(self._h, clips.ctypes.data, n, out.ctypes.data))
        return out

    def run_host_resampled(self, resampler: "Resampler", clips: np.ndarray, lengths: np.ndarray,
                           out: Optional[np.ndarray] = None) -> np.ndarray:
        """Clips at the files' rate, (N, in_stride) int16/float32 with ``lengths[i]`` valid samples each ->
        resampled on the device to this engine's rate, padded / trimmed to its n_samples, extracted:
        (N, rows, frames) float32.  The engine must take float32 input."""
        clips = np.ascontiguousarray(clips)
        if clips.dtype != np.int16:
            clips = clips.astype(np.float32, copy=False)
        if clips.ndim != 2:
            raise ValueError("clips must be (N, in_stride)")
        n = clips.shape[0]
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        if lengths.shape != (n,):
            raise ValueError("lengths must be (N,)")
        if out is None:
            out = np.empty((n, self.rows, self.frames), dtype=np.float32)
        elif out.shape != (n, self.rows, self.frames) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous float32 (N, rows, frames)")
        _check(self._lib.b2a_run_host_resampled(self._h, resampler._h, clips.ctypes.data,
                                                IN_I16 if clips.dtype == np.int16 else IN_F32, clips.shape[1],
                                                lengths.ctypes.data, n, out.ctypes.data))
        return out

    def run_host_copy_only(self, clips: np.ndarray, out: np.ndarray) -> None:
        """run_host's H2D / D2H schedule without the kernels (transfer ceiling; `out` is garbage)."""
        assert clips.flags.c_contiguous and out.flags.c_contiguous and clips.dtype == self.in_dtype
        _check(self._lib.b2a_run_host_copy_only(self._h, clips.ctypes.data, clips.shape[0], out.ctypes.data))

    def run_host_ragged(self, clips: list) -> list:
        """Variable-length clips (each 1-D, dtype = the engine's input dtype, n_fft <= len <=
        cfg.n_samples) -> list of (rows, 1 + len // hop) float32 arrays, one launch."""
        n = len(clips)
        if n == 0:
            return []
        lens = np.array([len(c) for c in clips], dtype=np.int32)
        starts = np.zeros(n, dtype=np.int64)
        pos = 0
        for i, L in enumerate(lens):            # 8-element alignment keeps the TMA staging path
            starts[i] = pos
            pos += (int(L) + 7) & ~7
        packed = np.zeros(pos, dtype=self.in_dtype)
        for c, s, L in zip(clips, starts, lens):
            packed[s:s + L] = c
        frames = 1 + lens // self.cfg.hop_length
        if self.cfg.kind == KIND_CLASSICAL:       # one aggregated vector per clip whatever its length
            frames = np.ones(n, dtype=np.int32)
        sizes = self.rows * frames.astype(np.int64)
        ooff = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        out = np.empty(int(sizes.sum()), dtype=np.float32)
        _check(self._lib.b2a_run_host_ragged(self._h, packed.ctypes.data, packed.size, starts.ctypes.data,
                                             lens.ctypes.data, ooff.ctypes.data, n, out.ctypes.data, out.size))
        return [out[o:o + sz].reshape(self.rows, f) for o, sz, f in zip(ooff, sizes, frames)]

    def run_device(self, d_clips: int, n_clips: int, d_out: int, stream: int = 0) -> None:
        """Raw device pointers (e.g. torch ``tensor.data_ptr()``); asynchronous on ``stream``."""
        _check(self._lib.b2a_run_device(self._h, C.c_void_p(d_clips), n_clips, C.c_void_p(d_out),
                                        C.c_void_p(stream)))

    def classical_tunings(self, n: int) -> np.ndarray:
        """Tuning estimate (fractions of a semitone) of clips [0, n) of the last device launch (classical handles)."""
        out = np.empty(n, dtype=np.float32)
        _check(self._lib.b2a_classical_tunings(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), n))
        return out

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.b2a_last_launch_count(self._h))

    # -- introspection ----------------------------------------------------------------------
    def table(self, which: int) -> np.ndarray:
        n = C.c_int64(0)
        _check(self._lib.b2a_get_table(self._h, which, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        _check(self._lib.b2a_get_table(self._h, which, out.ctypes.data, C.byref(n)))
        return out

    def cqt_geometry(self):
        no, nf = C.c_int32(), C.c_int32()
        _check(self._lib.b2a_cqt_geometry(self._h, C.byref(no), C.byref(nf), None, None, None))
        a = np.zeros(no.value, np.int32)
        b = np.zeros(no.value, np.int32)
        c = np.zeros(no.value, np.int32)
        _check(self._lib.b2a_cqt_geometry(self._h, C.byref(no), C.byref(nf), a.ctypes.data, b.ctypes.data,
                                          c.ctypes.data))
        return dict(n_octaves=no.value, n_filters=nf.value, n_fft=a, hop=b, sig_len=c)


# ---------------------------------------------------------------------------------------------
# Rational resampler (files whose rate differs from the extractor's sample_rate; deep.py:44-50)
# ---------------------------------------------------------------------------------------------

def _check_rs(rc: int) -> None:
    if rc != 0:
        raise B2AError(rc, load_library().b2a_resampler_last_error().decode(errors="replace"))


def resampler_design(orig_sr: int, target_sr: int):
    """(up, down, half_len, poly[up, K] float32) of the library's resampler — host only, no GPU."""
    lib = load_library()
    up, down, half, k = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _check_rs(lib.b2a_resampler_design(orig_sr, target_sr, C.byref(up), C.byref(down), C.byref(half), C.byref(k), None, 0))
    poly = np.empty((up.value, k.value), dtype=np.float32)
    _check_rs(lib.b2a_resampler_design(orig_sr, target_sr, None, None, None, None, poly.ctypes.data_as(C.c_void_p), poly.size))
    return up.value, down.value, half.value, poly


class Resampler:
    """One (orig_sr -> target_sr, device) resampler; GPU only (raises B2AError without a device)."""

    def __init__(self, orig_sr: int, target_sr: int, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check_rs(self._lib.b2a_resampler_create(int(orig_sr), int(target_sr), int(device), C.byref(self._h)))
        self.orig_sr, self.target_sr, self.device = int(orig_sr), int(target_sr), int(device)

    def out_len(self, n_in: int) -> int:
        return int(self._lib.b2a_resampler_out_len(self._h, int(n_in)))

    def run_host(self, y: np.ndarray) -> np.ndarray:
        """1-D int16 (PCM, scaled by 1/32768) or float32 samples -> float32 at the target rate."""
        if y.dtype != np.int16:
            y = y.astype(np.float32, copy=False)
        y = np.ascontiguousarray(y)
        if y.ndim != 1:
            raise ValueError("resampler input must be 1-D (mono)")
        out = np.empty(self.out_len(len(y)), dtype=np.float32)
        _check_rs(self._lib.b2a_resampler_run_host(self._h, y.ctypes.data_as(C.c_void_p),
                                                   IN_I16 if y.dtype == np.int16 else IN_F32, len(y),
                                                   out.ctypes.data_as(C.c_void_p)))
        return out

    def run_device(self, d_in: int, in_dtype: int, n_in: int, d_out: int, stream: int = 0) -> None:
        """Device pointers, asynchronous on `stream`; d_out must hold out_len(n_in) floats."""
        _check_rs(self._lib.b2a_resampler_run_device(self._h, C.c_void_p(d_in), in_dtype, n_in,
                                                     C.c_void_p(d_out), C.c_void_p(stream)))

    def close(self) -> None:
        if self._h:
            self._lib.b2a_resampler_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_resamplers: dict = {}
_resamplers_lock = threading.Lock()


def get_resampler(orig_sr: int, target_sr: int, device: int = 0) -> "Resampler":
    """The process-wide resampler of (orig, target, device); creation is serialised (decode workers
    race here), use is serialised inside the library (b2a_resampler_run_host holds the handle's mutex)."""
    key = (int(orig_sr), int(target_sr), int(device))
    with _resamplers_lock:
        r = _resamplers.get(key)
        if r is None:
            r = _resamplers[key] = Resampler(*key)
        return r


def resample(y: np.ndarray, orig_sr: int, target_sr: int, device: int = 0) -> np.ndarray:
    """librosa.resample(y, orig_sr, target_sr, res_type="soxr_hq") stand-in (see Resampler)."""
    return get_resampler(orig_sr, target_sr, device).run_host(y)

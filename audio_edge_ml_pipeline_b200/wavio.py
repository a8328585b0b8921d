"""WAV decode for the segment front end (reference: ``_load_segment`` deep.py:30-55 ->
``librosa.load(path, sr, offset, duration, mono=True)`` -> soundfile float32).

What is covered: RIFF/WAVE PCM 8/16/24/32-bit and IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE,
offset/duration slicing in native frames, channel mean.  Mono PCM16 is returned as int16 (the
GPU path applies the exact /32768); everything else as float32 scaled like libsndfile.
Files whose rate differs from ``sample_rate`` are resampled on the GPU (``resample_audio``; the
reference uses soxr_hq, see DESIGN.md for the stand-in's specification).  What is NOT covered
(SURVEY 8f N2): non-WAV containers — these raise, and the caller skips the sample the
way the reference skips any failing sample (base.py:204-206).
"""

from __future__ import annotations

import struct
from pathlib import Path
from typing import Optional, Tuple

import numpy as np


class AudioDecodeError(RuntimeError):
    pass


def _parse_header(buf: memoryview, total: Optional[int] = None):
    """fmt fields and (offset, size) of the data chunk; `total` = file size when `buf` is only its head."""
    if len(buf) < 12 or bytes(buf[0:4]) != b"RIFF" or bytes(buf[8:12]) != b"WAVE":
        raise AudioDecodeError("not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    n = len(buf)
    total = n if total is None else total
    while pos + 8 <= n:
        cid = bytes(buf[pos:pos + 4])
        size = struct.unpack_from("<I", buf, pos + 4)[0]
        body = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _br, align, bits = struct.unpack_from("<HHIIHH", buf, body)
            if tag == 0xFFFE and size >= 26:     # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                tag = struct.unpack_from("<H", buf, body + 24)[0]
            fmt = (tag, ch, sr, align, bits)
        elif cid == b"data":
            data = (body, min(size, total - body))
            break
        pos = body + size + (size & 1)
    if fmt is None or data is None:
        raise AudioDecodeError("missing fmt or data chunk")
    return fmt, data


def wav_info(path) -> dict:
    """duration / sample_rate / n_channels without decoding (audio_folder_loader.py:76-103)."""
    try:
        import os
        total = os.path.getsize(path)
        with open(path, "rb") as f:
            head = f.read(4096)                      # canonical headers are 44 bytes; LIST chunks rarely pass 4 KB
            try:
                (tag, ch, sr, align, bits), (off, size) = _parse_header(memoryview(head), total)
            except AudioDecodeError:
                head += f.read((1 << 16) - len(head))
                (tag, ch, sr, align, bits), (off, size) = _parse_header(memoryview(head), total)
        frames = size // max(align, 1)
        return {"duration": frames / sr if sr else 0.0, "sample_rate": int(sr), "n_channels": int(ch)}
    except Exception:
        return {"duration": 0.0, "sample_rate": 0, "n_channels": 0}


def decode_wav(path, offset: float = 0.0, duration: Optional[float] = None) -> Tuple[np.ndarray, int]:
    """-> (samples, native_rate); samples int16 (mono PCM16) or float32, 1-D (channel mean)."""
    raw = np.fromfile(str(path), dtype=np.uint8)
    (tag, ch, sr, align, bits), (off, size) = _parse_header(memoryview(raw))
    if ch < 1 or align < 1:
        raise AudioDecodeError("bad channel count / block align")
    n_frames = size // align
    start = min(int(offset * sr), n_frames)
    stop = n_frames if duration is None else min(n_frames, start + int(duration * sr))
    body = raw[off + start * align: off + stop * align]
    if tag == 1:
        if bits == 16:
            x = body.view("<i2").reshape(-1, ch)
            if ch == 1:
                return np.ascontiguousarray(x[:, 0]), int(sr)
            y = x.astype(np.float32) / np.float32(32768.0)
        elif bits == 8:
            y = (body.reshape(-1, ch).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
        elif bits == 24:
            b = body.reshape(-1, 3).astype(np.int32)
            v = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
            v = (v ^ 0x800000) - 0x800000
            y = (v.astype(np.float64) / 8388608.0).astype(np.float32).reshape(-1, ch)
        elif bits == 32:
            y = (body.view("<i4").astype(np.float64) / 2147483648.0).astype(np.float32).reshape(-1, ch)
        else:
            raise AudioDecodeError(f"unsupported PCM width {bits}")
    elif tag == 3:
        if bits == 32:
            y = body.view("<f4").reshape(-1, ch).astype(np.float32)
        elif bits == 64:
            y = body.view("<f8").reshape(-1, ch).astype(np.float32)
        else:
            raise AudioDecodeError(f"unsupported float width {bits}")
    else:
        raise AudioDecodeError(f"unsupported WAVE format tag {tag}")
    if ch > 1:
        y = y.mean(axis=1, dtype=np.float32)       # librosa.to_mono
    else:
        y = y[:, 0]
    return np.ascontiguousarray(y, dtype=np.float32), int(sr)


def load_segment(path, sample_rate: int, start_time, end_time, min_duration: float = 0.1,
                 min_samples: int = 1) -> np.ndarray:
    """deep.py:30-55 — decode, slice [start, end), mono, zero-pad to ``min_samples``."""
    offset = float(start_time) if start_time is not None else 0.0
    duration = None
    if end_time is not None:
        duration = max(float(end_time) - offset, min_duration)
    if Path(path).suffix.lower() not in (".wav", ".wave"):
        raise AudioDecodeError(f"unsupported container {Path(path).suffix!r} (WAV only; SURVEY 8f N2)")
    audio, sr = decode_wav(path, offset, duration)
    if sr != sample_rate:
        audio = resample_audio(audio, sr, sample_rate)      # librosa.load resamples here (soxr_hq)
    if len(audio) < min_samples:
        audio = np.pad(audio, (0, min_samples - len(audio)))
    return audio


def resample_audio(audio: np.ndarray, orig_sr: int, target_sr: int, device: int = 0) -> np.ndarray:
    """What ``librosa.load(..., sr=target_sr)`` does to a file recorded at another rate
    (deep.py:44-50): float32 mono at ``target_sr``, ``ceil(n * target / orig)`` samples.  Runs on the
    GPU (``b2a_resampler_*``; no CPU path: without a device this raises and the sample is skipped
    like any failing sample, base.py:204-206)."""
    from . import _lib
    return _lib.resample(audio, orig_sr, target_sr, device)


def pad_or_trim(audio: np.ndarray, target_len: int) -> np.ndarray:
    """deep.py:58-61."""
    if len(audio) >= target_len:
        return audio[:target_len]
    return np.pad(audio, (0, target_len - len(audio)))


def write_wav_pcm16(path, pcm: np.ndarray, sample_rate: int) -> None:
    """Minimal mono PCM16 writer (fixtures, end-to-end bench)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + pcm.nbytes, b"WAVE", b"fmt ", 16, 1, 1,
                      sample_rate, sample_rate * 2, 2, 16, b"data", pcm.nbytes)
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(pcm.tobytes())

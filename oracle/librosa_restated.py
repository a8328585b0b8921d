"""CPU oracle: numpy/scipy restatement of the reference's Stage-2 audio hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``audio_edge_ml_pipeline_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do,
and there only as the checker / reported baseline, never as the thing shipped.

What it restates
----------------
* The three ``extract`` bodies and helpers of the reference:
  ``src/preprocessing/feature_extraction/audio/deep.py:30-67`` (``_load_segment`` minus file
  decode, ``_pad_or_trim``, ``_normalize``), ``:112-134`` (``audio_mel_spec``), ``:235-260``
  (``audio_cqt``), ``:304-328`` (``audio_mfcc_seq``).
* The arithmetic those bodies delegate to **librosa==0.11.0** (``requirements.txt:55``), a
  third-party dependency that is NOT vendored under ``/root/reference`` and is NOT installed in
  this image: ``stft``, ``filters.mel``, ``feature.melspectrogram``, ``power_to_db``,
  ``amplitude_to_db``, ``feature.mfcc`` (via ``scipy.fft.dct``), ``cqt``/``vqt`` with
  ``filters.wavelet`` / ``wavelet_lengths`` / ``util.sparsify_rows``.  The published algorithms
  are restated here function by function, dtype by dtype (float64 window x float32 frames ->
  complex128 rfft -> complex64 storage; float32 filterbank; float32 dB arithmetic).

Pinning status
--------------
* log-mel: pinned against the reference's own second implementation, the C template in
  ``src/deployment/codegen/model_to_c.py:505-624`` (compiled into ``oracle/_ref`` by
  ``oracle/build_ref.py``), on inputs whose dynamic range stays inside 80 dB (that template
  omits ``top_db``), plus the written invariants in ``CLAUDE.md:88-92`` (501 frames,
  zero padding) and torch/torchaudio as independent witnesses (tests/test_oracle.py).
* mfcc: DCT pinned against the reference's ``_dct_matrix`` (``src/deployment/export_svm.py:
  69-79``) and ``scipy.fft.dct``; front end shared with log-mel.
* cqt: **parity unpinned** against real ``librosa.cqt``.  The octave recursion decimates
  with ``soxr_hq`` (libsoxr, absent here and on the GPU box).  This oracle *defines* the
  decimator (:func:`halfband_taps`) to soxr-HQ's published spec (pass band to 0.913 of the new
  Nyquist, stop band from 1.0, ~125 dB); the CUDA path uses the same taps.  Everything
  else in the CQT follows librosa 0.11.0 step by step.
"""

from __future__ import annotations

from functools import lru_cache
from typing import Optional

import numpy as np
import scipy.fft
import scipy.signal

# --------------------------------------------------------------------------------------
# deep.py helpers  (reference: src/preprocessing/feature_extraction/audio/deep.py)
# --------------------------------------------------------------------------------------


def pcm16_to_float(pcm: np.ndarray) -> np.ndarray:
    """soundfile/librosa.load PCM16 -> float32 in [-1, 1): exact ``x / 32768``
    (deep.py:44-50 via librosa.load; same scaling as model_to_c.py:577)."""
    return (np.asarray(pcm, dtype=np.int16).astype(np.float32)) / np.float32(32768.0)


def load_segment_tail(audio: np.ndarray, min_samples: int) -> np.ndarray:
    """deep.py:52-53 — right zero-pad to ``min_samples`` (decode itself is the caller's)."""
    audio = np.asarray(audio, dtype=np.float32)
    if len(audio) < min_samples:
        audio = np.pad(audio, (0, min_samples - len(audio)))
    return audio


def pad_or_trim(audio: np.ndarray, target_len: int) -> np.ndarray:
    """deep.py:58-61."""
    if len(audio) >= target_len:
        return audio[:target_len]
    return np.pad(audio, (0, target_len - len(audio)))


def normalize01(x: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """deep.py:64-67 — min-max to [0,1]; float32 throughout under NumPy-2 promotion."""
    lo, hi = x.min(), x.max()
    return (x - lo) / (hi - lo + eps)


def prepare_audio(audio: np.ndarray, sample_rate: int, duration: Optional[float],
                  min_samples: int) -> np.ndarray:
    """Everything deep.py does between decode and the librosa call (``:119-124`` etc.)."""
    audio = load_segment_tail(audio, min_samples)
    if duration is not None:
        audio = pad_or_trim(audio, int(duration * sample_rate))
    return np.ascontiguousarray(audio, dtype=np.float32)


# --------------------------------------------------------------------------------------
# librosa 0.11.0: stft / filters.mel / melspectrogram / power_to_db / amplitude_to_db
# --------------------------------------------------------------------------------------


def n_frames_centered(n_samples: int, hop_length: int) -> int:
    """``1 + n // hop`` — CLAUDE.md:90, model_to_c.py:567 (n_fft even)."""
    return 1 + n_samples // hop_length


def stft(y: np.ndarray, n_fft: int, hop_length: int, window: str = "hann",
         pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(center=True): zero (or reflect) pad n_fft//2 each side, periodic window
    in float64, rfft in double, result stored complex64, shape (1+n_fft//2, T)."""
    y = np.asarray(y, dtype=np.float32)
    if window == "hann":
        win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    elif window == "ones":
        win = np.ones(n_fft, dtype=np.float64)
    else:  # pragma: no cover
        raise ValueError(window)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode=pad_mode)
    n_frames = 1 + (len(ypad) - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        ypad, shape=(n_fft, n_frames),
        strides=(ypad.strides[0], ypad.strides[0] * hop_length), writeable=False)
    out = np.empty((1 + n_fft // 2, n_frames), dtype=np.complex64, order="F")
    # float64 window * float32 frames -> float64; rfft -> complex128; store -> complex64
    out[...] = scipy.fft.rfft(win[:, None] * frames, axis=0)
    return out


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


@lru_cache(maxsize=32)
def mel_filterbank(sr: int, n_fft: int, n_mels: int = 128, fmin: float = 0.0,
                   fmax: Optional[float] = None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> (n_mels, 1+n_fft//2)."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    weights.setflags(write=False)
    return weights


def melspectrogram(y: np.ndarray, sr: int, n_fft: int, hop_length: int, n_mels: int = 128,
                   pad_mode: str = "constant") -> np.ndarray:
    """librosa.feature.melspectrogram(power=2.0): |stft|**2 (float32) then einsum with the
    float32 filterbank."""
    D = stft(y, n_fft=n_fft, hop_length=hop_length, pad_mode=pad_mode)
    S = np.abs(D) ** 2.0
    basis = mel_filterbank(sr, n_fft, n_mels)
    return np.einsum("ft,mf->mt", S, basis, optimize=True)


def power_to_db(S: np.ndarray, ref=1.0, amin: float = 1e-10,
                top_db: Optional[float] = 80.0) -> np.ndarray:
    """librosa.power_to_db — float32 in, float32 out."""
    S = np.asarray(S)
    magnitude = S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S: np.ndarray, ref=1.0, amin: float = 1e-5,
                    top_db: Optional[float] = 80.0) -> np.ndarray:
    """librosa.amplitude_to_db: ref evaluated on |S| first, then squared."""
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude)
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


# --------------------------------------------------------------------------------------
# librosa 0.11.0: feature.mfcc
# --------------------------------------------------------------------------------------


def mfcc(y: np.ndarray, sr: int, n_mfcc: int, n_fft: int, hop_length: int,
         n_mels: int = 128) -> np.ndarray:
    """librosa.feature.mfcc(dct_type=2, norm='ortho', lifter=0):
    power_to_db(melspectrogram) with ref=1.0/top_db=80, then scipy DCT-II along the mel axis."""
    S = power_to_db(melspectrogram(y, sr=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels))
    return scipy.fft.dct(S, axis=-2, type=2, norm="ortho")[:n_mfcc, :]


def dct2_ortho_matrix(n_out: int, n_in: int) -> np.ndarray:
    """float64 orthonormal DCT-II matrix (same definition as export_svm.py:69-79)."""
    k = np.arange(n_out)[:, None]
    n = np.arange(n_in)[None, :]
    d = np.cos(np.pi / n_in * (n + 0.5) * k) * np.sqrt(2.0 / n_in)
    d[0] /= np.sqrt(2.0)
    return d


# --------------------------------------------------------------------------------------
# Decimator standing in for soxr_hq 2:1 (see module docstring: defined here, shared taps)
# --------------------------------------------------------------------------------------

HALFBAND_PASS = 0.913     # pass-band edge, fraction of the NEW Nyquist (soxr HQ spec)
HALFBAND_STOP = 1.0       # stop-band start, fraction of the NEW Nyquist
HALFBAND_ATTEN_DB = 125.0
HALFBAND_NUMTAPS = 383    # odd; measured: pass ripple 6e-7, stop band -124.6 dB from the new Nyquist


@lru_cache(maxsize=1)
def halfband_taps() -> np.ndarray:
    """Linear-phase Kaiser-windowed-sinc low-pass at the INPUT rate, float64, unity DC gain.

    Cut-off is the middle of the transition band [0.913, 1.0] x (fs_in/4); beta from the
    Kaiser formula for HALFBAND_ATTEN_DB.  The CUDA library evaluates the same closed form
    (sinc x I0 window) in double and rounds to float32 (csrc/tables.cpp)."""
    n = HALFBAND_NUMTAPS
    a = HALFBAND_ATTEN_DB
    beta = 0.1102 * (a - 8.7)
    fc = 0.5 * (HALFBAND_PASS + HALFBAND_STOP) * 0.25   # cycles/sample at the input rate
    m = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = 2.0 * fc * np.sinc(2.0 * fc * m)
    r = 2.0 * m / (n - 1)
    w = np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / np.i0(beta)
    h = h * w
    h /= h.sum()
    h.setflags(write=False)
    return h


def decimate2(y: np.ndarray) -> np.ndarray:
    """Stand-in for ``librosa.resample(y, orig_sr=2, target_sr=1, res_type='soxr_hq',
    scale=True)``: zero-phase FIR, zero-extended edges, output length ceil(n/2), x sqrt(2).
    Taps are float32-rounded (what the GPU holds); accumulation here is float64; output float32."""
    y = np.asarray(y, dtype=np.float32)
    h = halfband_taps().astype(np.float32).astype(np.float64)
    n_out = int(np.ceil(len(y) * 0.5))
    full = np.convolve(y.astype(np.float64), h)           # full[k] = sum_j h[j] y[k-j]
    c = (len(h) - 1) // 2
    out = full[c:c + 2 * n_out:2]
    if len(out) < n_out:  # pragma: no cover
        out = np.pad(out, (0, n_out - len(out)))
    return (out * np.sqrt(2.0)).astype(np.float32)


def resampler_prototype(orig_sr: int, target_sr: int):
    """(up, down, half_len, g): the rational resampler's prototype low-pass, float64, unity DC
    gain, running at up x orig_sr.  Same specification as halfband_taps() scaled to the ratio
    (2:1 reproduces it exactly); the CUDA library evaluates the same closed form
    (csrc/tables.cpp: design_resampler)."""
    from math import gcd
    g_ = gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g_, int(orig_sr) // g_
    lo = float(min(orig_sr, target_sr))
    fs_up = float(up) * float(orig_sr)
    half = int(np.floor(95.5 * fs_up / lo + 0.5))
    beta = 0.1102 * (HALFBAND_ATTEN_DB - 8.7)
    fc = 0.5 * (HALFBAND_PASS + HALFBAND_STOP) * 0.5 * lo / fs_up
    m = np.arange(-half, half + 1, dtype=np.float64)
    h = 2.0 * fc * np.sinc(2.0 * fc * m)
    w = np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - (m / half) ** 2))) / np.i0(beta)
    h = h * w
    h /= h.sum()
    return up, down, half, h


def resample_restated(y: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """Stand-in for ``librosa.resample(y, orig_sr=, target_sr=, res_type="soxr_hq")`` as
    ``librosa.load`` applies it (deep.py:44-50): zero-phase FIR, zero-extended edges, output length
    ceil(n * target / orig).  Parity against libsoxr itself is unpinned (absent offline).
    Taps are float32-rounded (what the GPU holds), arithmetic float64 through scipy's own
    polyphase engine (an independent check of the indexing), output float32."""
    import scipy.signal
    y = np.asarray(y, dtype=np.float32)
    if int(orig_sr) == int(target_sr):
        return y.copy()
    up, down, _half, g = resampler_prototype(orig_sr, target_sr)
    g32 = (g * up).astype(np.float32).astype(np.float64) / up       # resample_poly multiplies by `up` itself
    out = scipy.signal.resample_poly(y.astype(np.float64), up, down, window=g32, padtype="constant")
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# librosa 0.11.0: cqt = vqt(gamma=0, intervals='equal')
# --------------------------------------------------------------------------------------

C1_HZ = 32.70319566257483   # librosa.note_to_hz("C1")
HANN_BANDWIDTH = 1.50018310546875   # librosa.filters.window_bandwidth("hann")


def cqt_frequencies(n_bins: int, fmin: float, bins_per_octave: int) -> np.ndarray:
    return fmin * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave)


def relative_bandwidth(freqs: np.ndarray) -> np.ndarray:
    bpo = np.empty_like(freqs)
    logf = np.log2(freqs)
    bpo[0] = 1 / (logf[1] - logf[0])
    bpo[-1] = 1 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2 / (logf[2:] - logf[:-2])
    return (2.0 ** (2 / bpo) - 1) / (2.0 ** (2 / bpo) + 1)


def wavelet_lengths(freqs, sr, alpha, filter_scale: float = 1.0, gamma: float = 0.0):
    Q = float(filter_scale) / alpha
    f_cutoff = max(freqs * (1 + 0.5 * HANN_BANDWIDTH / Q) + 0.5 * gamma)
    lengths = Q * sr / (freqs + gamma / alpha)
    return lengths, f_cutoff


def _float_hann(n: int) -> np.ndarray:
    return scipy.signal.get_window("hann", n, fftbins=True)


def wavelet_basis(freqs, sr, alpha):
    """librosa.filters.wavelet(window='hann', norm=1, pad_fft=True, dtype=complex64)."""
    lengths, _ = wavelet_lengths(freqs, sr, alpha)
    filts = []
    for ilen, freq in zip(lengths, freqs):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(1j * (t * 2 * np.pi * freq / sr))          # util.phasor
        sig = sig * _float_hann(len(sig))
        sig = sig / np.sum(np.abs(sig))                          # util.normalize(norm=1)
        filts.append(sig)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    out = np.zeros((len(filts), max_len), dtype=np.complex64)
    for i, f in enumerate(filts):
        lpad = (max_len - len(f)) // 2                           # util.pad_center
        out[i, lpad:lpad + len(f)] = f
    return out, lengths


def sparsify_rows(x: np.ndarray, quantile: float = 0.01) -> np.ndarray:
    """librosa.util.sparsify_rows, returned dense (zeros where librosa's CSR has no entry)."""
    out = np.zeros_like(x)
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    for i, j in enumerate(threshold_idx):
        idx = np.where(mags[i] >= mag_sort[i, j])
        out[i, idx] = x[i, idx]
    return out


def vqt_filter_fft(sr, freqs, alpha):
    basis, lengths = wavelet_basis(freqs, sr, alpha)
    n_fft = basis.shape[1]
    basis *= (lengths[:, np.newaxis] / float(n_fft))
    fft_basis = scipy.fft.fft(basis, n=n_fft, axis=1)[:, :(n_fft // 2) + 1]
    return sparsify_rows(fft_basis, 0.01).astype(np.complex64), n_fft, lengths


def _num_two_factors(x: int) -> int:
    if x <= 0:
        return 0
    n = 0
    while x % 2 == 0:
        n += 1
        x //= 2
    return n


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves) -> int:
    c1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    c2 = max(0, _num_two_factors(hop_length) - n_octaves + 1)
    return min(c1, c2)


def cqt_plan(sr: float, hop_length: int, n_bins: int, bins_per_octave: int,
             fmin: Optional[float]):
    """Static part of librosa.cqt for one configuration: per-octave (basis, n_fft, hop, rate),
    early-downsample count, final lengths."""
    if fmin is None:
        fmin = C1_HZ
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
    alpha = relative_bandwidth(freqs)
    lengths, filter_cutoff = wavelet_lengths(freqs, sr, alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ValueError(f"Wavelet basis with max frequency={np.max(freqs)} would exceed the "
                         f"Nyquist frequency={nyquist}. Try reducing the number of frequency bins.")
    n_early = early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves)
    sr0 = sr / float(2 ** n_early)
    hop0 = hop_length // (2 ** n_early)
    octs = []
    my_sr, my_hop = sr0, hop0
    for i in range(n_octaves):
        sl = slice(-n_filters, None) if i == 0 else slice(-n_filters * (i + 1), -n_filters * i)
        fft_basis, n_fft, _ = vqt_filter_fft(my_sr, freqs[sl], alpha[sl])
        fft_basis = fft_basis * np.sqrt(sr0 / my_sr)
        decimate_after = (my_hop % 2 == 0)
        octs.append(dict(basis=fft_basis.astype(np.complex64), n_fft=n_fft, hop=my_hop,
                         sr=my_sr, decimate_after=decimate_after, rows=sl))
        if decimate_after:
            my_hop //= 2
            my_sr /= 2.0
    final_lengths, _ = wavelet_lengths(freqs, sr0, alpha)
    return dict(freqs=freqs, alpha=alpha, n_early=n_early, octaves=octs,
                lengths=final_lengths, n_octaves=n_octaves, n_filters=n_filters)


def cqt(y: np.ndarray, sr: int = 22050, hop_length: int = 512, n_bins: int = 84,
        bins_per_octave: int = 12, fmin: Optional[float] = None) -> np.ndarray:
    """librosa.cqt(...) with defaults tuning=0, filter_scale=1, norm=1, sparsity=0.01,
    window='hann', scale=True, pad_mode='constant', res_type -> :func:`decimate2`."""
    y = np.asarray(y, dtype=np.float32)
    plan = cqt_plan(float(sr), hop_length, n_bins, bins_per_octave, fmin)
    for _ in range(plan["n_early"]):
        y = decimate2(y)      # scale=True keeps sqrt(2) per stage == sqrt(factor) overall
    resp = []
    my_y = y
    for o in plan["octaves"]:
        D = stft(my_y, n_fft=o["n_fft"], hop_length=o["hop"], window="ones")
        resp.append((o["basis"] @ D).astype(np.complex64))
        if o["decimate_after"]:
            my_y = decimate2(my_y)
    max_col = min(r.shape[-1] for r in resp)
    V = np.empty((n_bins, max_col), dtype=np.complex64, order="F")
    end = n_bins
    for r in resp:
        n_oct = r.shape[0]
        if end < n_oct:
            V[:end, :] = r[-end:, :max_col]
        else:
            V[end - n_oct:end, :] = r[:, :max_col]
        end -= n_oct
    V /= np.sqrt(plan["lengths"])[:, None]
    return V


# --------------------------------------------------------------------------------------
# The three extractors, end to end from decoded float32 audio
# --------------------------------------------------------------------------------------


def audio_mel_spec(audio: np.ndarray, sample_rate: int = 16000, n_mels: int = 40,
                   n_fft: int = 512, hop_length: int = 160,
                   duration: Optional[float] = None, pad_mode: str = "constant") -> np.ndarray:
    """deep.py:112-134."""
    y = prepare_audio(audio, sample_rate, duration, min_samples=n_fft)
    mel = melspectrogram(y, sr=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                         pad_mode=pad_mode)
    log_mel = power_to_db(mel, ref=np.max)
    return normalize01(log_mel).astype(np.float32)


def audio_mfcc_seq(audio: np.ndarray, sample_rate: int = 22050, n_mfcc: int = 40,
                   n_fft: int = 1024, hop_length: int = 512,
                   duration: Optional[float] = None, n_mels: int = 128) -> np.ndarray:
    """deep.py:304-328 (``n_mels`` is this project's extension key; 128 = reference)."""
    y = prepare_audio(audio, sample_rate, duration, min_samples=n_fft)
    m = mfcc(y, sr=sample_rate, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)
    mean = m.mean(axis=1, keepdims=True)
    std = m.std(axis=1, keepdims=True) + 1e-8
    return ((m - mean) / std).astype(np.float32)


def audio_cqt(audio: np.ndarray, sample_rate: int = 22050, hop_length: int = 512,
              n_bins: int = 84, bins_per_octave: int = 12, fmin: Optional[float] = None,
              duration: Optional[float] = None) -> np.ndarray:
    """deep.py:235-260."""
    y = prepare_audio(audio, sample_rate, duration, min_samples=hop_length * 2)
    c = np.abs(cqt(y, sr=sample_rate, hop_length=hop_length, n_bins=n_bins,
                   bins_per_octave=bins_per_octave, fmin=fmin))
    log_cqt = amplitude_to_db(c, ref=np.max)
    return normalize01(log_cqt).astype(np.float32)

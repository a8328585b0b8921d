"""TEST INFRASTRUCTURE: an oracle-backed stand-in for ``_lib.Engine`` so the host logic
(batching, grouping, skip-on-error, label indexing, persistence, sharding) can be exercised on a
box without a GPU.  Installed by monkeypatching ``extractors._make_engine``; never shipped."""
import numpy as np

from audio_edge_ml_pipeline_b200 import _lib as B
from oracle import librosa_restated as L


class FakeEngine:
    calls = []          # (device, n_clips) per run_host, for sharding assertions

    def __init__(self, cfg, device):
        self.cfg, self.device = cfg, device
        n = cfg.n_samples
        if cfg.kind == B.KIND_CQT:
            plan = L.cqt_plan(float(cfg.sample_rate), cfg.hop_length, cfg.n_bins, cfg.bins_per_octave,
                              cfg.fmin if cfg.fmin > 0 else None)   # raises like librosa on a bad config
            self.rows, self.frames = cfg.n_bins, 1 + n // cfg.hop_length
            del plan
        else:
            self.rows = cfg.n_mfcc if cfg.kind == B.KIND_MFCC else cfg.n_mels
            self.frames = 1 + n // cfg.hop_length

    def close(self):
        pass

    def _one(self, x):
        c = self.cfg
        y = L.pcm16_to_float(x) if x.dtype == np.int16 else x.astype(np.float32)
        if c.kind == B.KIND_MEL:
            return L.audio_mel_spec(y, c.sample_rate, c.n_mels, c.n_fft, c.hop_length, None,
                                    "constant" if c.pad_mode == 0 else "reflect")
        if c.kind == B.KIND_MFCC:
            return L.audio_mfcc_seq(y, c.sample_rate, c.n_mfcc, c.n_fft, c.hop_length, None, n_mels=c.n_mels)
        return L.audio_cqt(y, c.sample_rate, c.hop_length, c.n_bins, c.bins_per_octave,
                           c.fmin if c.fmin > 0 else None, None)

    def run_host_ragged(self, clips):
        FakeEngine.calls.append((self.device, len(clips)))
        assert all(len(c) <= self.cfg.n_samples for c in clips)
        return [self._one(np.asarray(c)) for c in clips]

    def run_host(self, clips, out=None):
        FakeEngine.calls.append((self.device, len(clips)))
        res = np.stack([self._one(c) for c in clips]) if len(clips) else np.empty((0, self.rows, self.frames), np.float32)
        if out is None:
            return res
        out[...] = res
        return out

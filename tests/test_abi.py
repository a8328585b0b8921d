"""The C-ABI library loads here (no GPU) and exports every symbol include/b2a.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "b2a.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2a_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_documented_surface():
    names = _declared_symbols()
    for must in ("b2a_create", "b2a_destroy", "b2a_out_shape", "b2a_run_host", "b2a_run_device", "b2a_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(str(lib_built))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/b2a.h but not exported"
    from audio_edge_ml_pipeline_b200 import _lib as B
    assert sorted(B.EXPORTED_SYMBOLS) == _declared_symbols()
    assert lib.b2a_abi_version() == 1


def test_config_struct_layout_matches_header(lib_built):
    from audio_edge_ml_pipeline_b200 import _lib as B
    assert ctypes.sizeof(B.B2AConfig) == 4 * 10 + 8 + 4 + 4 + 32
    cfg = B.default_config(B.KIND_MEL)         # pure host code: reference defaults, deep.py:98-105
    assert (cfg.sample_rate, cfg.n_mels, cfg.n_fft, cfg.hop_length, cfg.top_db) == (16000, 40, 512, 160, 80.0)
    cfg = B.default_config(B.KIND_MFCC)        # deep.py:290-297 + librosa n_mels
    assert (cfg.sample_rate, cfg.n_mfcc, cfg.n_fft, cfg.hop_length, cfg.n_mels) == (22050, 40, 1024, 512, 128)
    cfg = B.default_config(B.KIND_CQT)         # deep.py:219-227
    assert (cfg.sample_rate, cfg.hop_length, cfg.n_bins, cfg.bins_per_octave) == (22050, 512, 84, 12)


def test_no_cpu_fallback_engine_fails_loudly_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audio_edge_ml_pipeline_b200 import _lib as B
    cfg = B.default_config(B.KIND_MEL)
    cfg.n_samples = 80000
    with pytest.raises(B.B2AError) as ei:
        B.Engine(cfg, 0)
    assert ei.value.code == -4 and "no CPU path" in str(ei.value)
    import numpy as np
    import audio_edge_ml_pipeline_b200 as P
    with pytest.raises(B.B2AError):
        P.AudioMelSpectrogram(duration=5.0).extract_array(np.zeros(80000, np.int16))


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "audio_edge_ml_pipeline_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in re.sub(r"#.*", "", src).replace("oracle-backed", ""), py
    for cu in (ROOT / "audio_edge_ml_pipeline_b200" / "csrc").rglob("*"):
        if cu.is_file():
            assert "#include \"../../oracle" not in cu.read_text()

"""Build ``oracle/_ref/`` — the reference's OWN second implementation of the log-mel path.

TEST INFRASTRUCTURE ONLY (see oracle/librosa_restated.py header).

The reference carries an in-tree C99 implementation of ``audio_mel_spec`` as string templates
inside ``/root/reference/src/deployment/codegen/model_to_c.py`` (``_FEATURES_H`` :476-503,
``_FEATURES_C`` :505-624) that it writes out for a microcontroller.  This recipe reads those two
string constants **where they lie** (AST parse, no import: the module's siblings need
tensorflow), instantiates the header exactly as ``_gen_features`` (:1330-1345) does, emits the
mel filterbank table exactly as ``_gen_feat_data`` (:1098-1136) does (``%.8g`` literals of the
float32 filterbank; librosa being absent, the table comes from the restated ``filters.mel``),
and compiles the result with plain ``gcc -O2`` into ``oracle/_ref/libfeatures_ref_<tag>.so``.

Outputs go ONLY under ``oracle/_ref/`` (git-ignored, not gpurun-ignored).  No reference source
is copied into the tracked tree.  ``/root/reference`` exists only in the authoring container; on
the GPU box the prebuilt ``.so`` files travel with the snapshot and this script is a no-op.
"""

from __future__ import annotations

import ast
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_ROOT = Path("/root/reference")
TEMPLATE_PY = REF_ROOT / "src/deployment/codegen/model_to_c.py"
OUT = HERE / "_ref"

# (tag, sample_rate, n_fft, hop, n_mels, n_samples)
VARIANTS = [
    ("16k_512_160_40_5s", 16000, 512, 160, 40, 80000),   # config/feature_extraction.yaml:60-70
    ("16k_512_160_40_1s", 16000, 512, 160, 40, 16000),   # short variant for small fixtures
]


def _template_strings() -> dict[str, str]:
    tree = ast.parse(TEMPLATE_PY.read_text())
    out: dict[str, str] = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id in ("_FEATURES_H", "_FEATURES_C"):
                out[t.id] = ast.literal_eval(node.value)
    if set(out) != {"_FEATURES_H", "_FEATURES_C"}:
        raise RuntimeError("reference C template not found in " + str(TEMPLATE_PY))
    return out


def _flt(v) -> str:
    # literal format of model_to_c.py:1108-1112
    s = f"{v:.8g}"
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s + "f"


def _feat_data(mel_fb: np.ndarray) -> tuple[str, str]:
    n_mels, n_bins = mel_fb.shape
    rows = ["    {" + ", ".join(_flt(v) for v in row) + "}" for row in mel_fb]
    header = ("#pragma once\n#include \"features.h\"\n\n"
              f"extern const float feat_mel_fb[{n_mels}][{n_bins}];\n")
    source = ("#include \"feat_data.h\"\n\n"
              f"const float feat_mel_fb[{n_mels}][{n_bins}] = {{\n" + ",\n".join(rows) + "\n};\n")
    return header, source


def lib_path(tag: str) -> Path:
    return OUT / f"libfeatures_ref_{tag}.so"


def build(force: bool = False) -> list[Path]:
    """Compile every variant; returns the built libraries.  No-op (returns what exists) when the
    reference tree is absent."""
    libs = [lib_path(v[0]) for v in VARIANTS]
    if not TEMPLATE_PY.exists():
        return [p for p in libs if p.exists()]
    if not force and all(p.exists() and p.stat().st_mtime >= Path(__file__).stat().st_mtime
                         for p in libs):
        return libs
    sys.path.insert(0, str(HERE.parent))
    from oracle import librosa_restated as L

    tpl = _template_strings()
    OUT.mkdir(exist_ok=True)
    for tag, sr, n_fft, hop, n_mels, n_samples in VARIANTS:
        gen = OUT / f"gen_{tag}"
        gen.mkdir(exist_ok=True)
        n_frames = 1 + n_samples // hop
        (gen / "features.h").write_text(tpl["_FEATURES_H"].format(
            sample_rate=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels,
            n_samples=(n_frames - 1) * hop, n_frames=n_frames))
        (gen / "features.c").write_text(tpl["_FEATURES_C"])
        h, c = _feat_data(L.mel_filterbank(sr, n_fft, n_mels).astype(np.float32))
        (gen / "feat_data.h").write_text(h)
        (gen / "feat_data.c").write_text(c)
        subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-iquote", str(gen),
                        str(gen / "features.c"), str(gen / "feat_data.c"),
                        "-o", str(lib_path(tag)), "-lm"], check=True)
    return libs


def features_extract(tag: str, pcm: np.ndarray) -> np.ndarray:
    """Call the compiled reference ``features_extract(const int16_t*, int, float*)``."""
    import ctypes

    v = {x[0]: x for x in VARIANTS}[tag]
    _, _sr, _n_fft, hop, n_mels, n_samples = v
    lib = ctypes.CDLL(str(lib_path(tag)))
    lib.features_extract.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.features_extract.restype = None
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    assert pcm.shape == (n_samples,)
    out = np.empty((n_mels, 1 + n_samples // hop), dtype=np.float32)
    lib.features_extract(pcm.ctypes.data, n_samples, out.ctypes.data)
    return out


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print(p)

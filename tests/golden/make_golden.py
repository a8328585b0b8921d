"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the authoring
container, where /root/reference exists; the fixtures travel, the reference does not).

  logmel_ref_c_1s.npz   inputs (int16, 1 s @ 16 kHz) and the outputs of the reference's own C
                        implementation of audio_mel_spec (model_to_c.py:505-624, compiled by
                        oracle/build_ref.py).  Inputs are noise-like so the template's missing
                        top_db clip never engages: max dynamic range < 80 dB.
  dct_ref.npz           the reference's ``_dct_matrix`` (src/deployment/export_svm.py:69-79),
                        executed from its source, for (13,40) and (40,128).
  helpers_ref.npz       the reference's ``_pad_or_trim`` / ``_normalize`` (deep.py:58-67),
                        executed from their source on seeded inputs.

    python tests/golden/make_golden.py
"""
import ast
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _func_from_source(py: Path, name: str, glb: dict):
    tree = ast.parse(py.read_text())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, str(py), "exec"), glb)
            return glb[name]
    raise KeyError(name)


def main():
    from oracle import build_ref
    from audio_edge_ml_pipeline_b200 import synth

    build_ref.build()
    rng = np.random.default_rng(20261018)
    n = 16000
    clips = []
    for fam, scale in ((0, 1.0), (0, 1.0), (1, 1.0), (1, 1.0), (0, 1.0), (1, 1.0)):
        y = synth.make_clip(rng, fam, 16000, n)
        clips.append(synth.to_pcm16(y))
    pcm = np.stack(clips)
    ref = np.stack([build_ref.features_extract("16k_512_160_40_1s", c) for c in pcm])
    np.savez_compressed(OUT / "logmel_ref_c_1s.npz", pcm=pcm, ref=ref)

    glb = {"np": np}
    dct = _func_from_source(REF / "src/deployment/export_svm.py", "_dct_matrix", glb)
    np.savez_compressed(OUT / "dct_ref.npz", d13_40=dct(13, 40), d40_128=dct(40, 128))

    glb = {"np": np}
    deep = REF / "src/preprocessing/feature_extraction/audio/deep.py"
    pad_or_trim = _func_from_source(deep, "_pad_or_trim", glb)
    normalize = _func_from_source(deep, "_normalize", glb)
    a = rng.standard_normal(1000).astype(np.float32)
    x = (rng.standard_normal((40, 101)) * 20 - 40).astype(np.float32)
    np.savez_compressed(OUT / "helpers_ref.npz", a=a, trim=pad_or_trim(a, 600), pad=pad_or_trim(a, 1500),
                        x=x, norm=normalize(x), norm_const=normalize(np.full((4, 5), -100.0, np.float32)))
    # Stage-1b augmentors, executed from the reference's own source (augment.py:88-212; the module itself
    # imports librosa/soundfile at import time, so its functions are lifted one by one like the helpers above)
    glb = {"np": np}
    aug = REF / "src/preprocessing/augment.py"
    names = {"volume_scale": "_volume_scale", "gaussian_noise": "_gaussian_noise", "time_shift": "_time_shift",
             "polarity_inversion": "_polarity_inversion", "pdm_hiss": "_pdm_hiss"}
    glb["_AUGMENTORS"] = {k: _func_from_source(aug, v, glb) for k, v in names.items()}
    apply_aug = _func_from_source(aug, "_apply_augmentations", glb)
    preserve = _func_from_source(aug, "_preserve_length", glb)
    specs = [
        {"type": "volume_scale", "min_gain": 0.7, "max_gain": 1.3},
        {"type": "pdm_hiss", "min_amplitude": 0.01, "max_amplitude": 0.04},
        {"type": "gaussian_noise", "min_amplitude": 0.001, "max_amplitude": 0.004},
        {"type": "time_shift", "max_fraction": 0.2},
        {"type": "polarity_inversion"},
    ]                                   # config/augmentation.yaml:38-60 minus the two librosa-backed steps
    arng = np.random.default_rng(42)    # augment.py:325 with the YAML's seed (augmentation.yaml:22)
    srng = np.random.default_rng(99)
    sr, n = 16000, 4000
    clips = [(0.3 * srng.standard_normal(n)).astype(np.float32), synth.make_clip(srng, 2, sr, n).astype(np.float32),
             np.clip(3.0 * srng.standard_normal(n), -1, 1).astype(np.float32)]
    outs = []
    for y in clips:                     # the per-file loop of run(): 4 copies each, one rng for the whole run
        for _ in range(4):
            outs.append(preserve(apply_aug(y, sr, specs, arng), len(y)))
    np.savez_compressed(OUT / "augment_ref.npz", clips=np.stack(clips), out=np.stack(outs), sr=sr, seed=42,
                        specs=np.array([repr(specs)]), pad=preserve(clips[0][:100], 150), trim=preserve(clips[0], 90))
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()

"""Thin batch driver with the reference's persistence contract.

``FeaturePipeline.run/save/load`` mirror ``src/preprocessing/pipeline.py:73-235``: the directory
written by ``save`` is byte-for-byte the layout the reference's consumers read — ``features.npy``
(NPY v1, float32, C-order, ``(N, rows, T)``), ``labels.npy`` (int32), ``label_names.json``,
``metadata.json`` (``default=str``), ``info.json``.  ``run_config`` accepts the reference's
``config/feature_extraction.yaml`` unchanged for the audio experiments (``audio_folder`` loader,
``audio_mel_spec`` / ``audio_mfcc_seq`` / ``audio_cqt`` extractors, ``extractor_params`` splatted
into the constructor exactly like pipeline.py:524 — an unknown key is a TypeError there and here).

    python -m audio_edge_ml_pipeline_b200.pipeline --config config/feature_extraction.yaml
"""

from __future__ import annotations

import argparse
import inspect
import json
import logging
import shutil
from pathlib import Path
from typing import Optional

import numpy as np

from . import extractors as _extractors  # noqa: F401  (registers the three extractors)
from .base import BaseDatasetLoader, BaseFeatureExtractor, FeatureSet
from .loaders import AudioFolderLoader
from .registry import get

logger = logging.getLogger(__name__)


class FeaturePipeline:
    def __init__(self, loader: BaseDatasetLoader, extractor: BaseFeatureExtractor) -> None:
        self.loader = loader
        self.extractor = extractor

    def run(self, max_samples: Optional[int] = None, output_dir=None) -> FeatureSet:
        """``output_dir`` (optional, an addition to the reference's signature): where ``save`` will put this run.
        Extractors that can then write their rows straight into ``output_dir/features.npy`` (no second copy)."""
        logger.info("Starting extraction: loader=%s (%d samples), extractor=%s",
                    type(self.loader).__name__, len(self.loader), self.extractor.name)
        kw = {}
        if output_dir is not None and "features_out" in inspect.signature(self.extractor.extract_dataset).parameters:
            kw["features_out"] = Path(output_dir) / "features.npy"
        fs = self.extractor.extract_dataset(self.loader, max_samples=max_samples, **kw)
        logger.info("Extraction complete: %s", fs)
        return fs

    @staticmethod
    def save(fs: FeatureSet, output_dir) -> None:
        output_dir = Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        placed = getattr(fs, "features_file", None)
        if placed is not None and Path(placed).resolve() == (output_dir / "features.npy").resolve() \
                and isinstance(fs.features, np.memmap):
            fs.features.flush()                        # extract_dataset wrote the rows into this very file
        else:
            np.save(output_dir / "features.npy", fs.features)
        if fs.labels is not None:
            np.save(output_dir / "labels.npy", fs.labels)
        if fs.label_names is not None:
            (output_dir / "label_names.json").write_text(json.dumps(fs.label_names, indent=2))
        if fs.cluster_assignments is not None:
            np.save(output_dir / "cluster_assignments.npy", fs.cluster_assignments)
        (output_dir / "metadata.json").write_text(json.dumps(fs.metadata, indent=2, default=str))
        info = {"feature_type": fs.feature_type, "modality": fs.modality, "n_samples": fs.n_samples,
                "feature_shape": list(fs.feature_shape), "n_classes": fs.n_classes,
                "is_supervised": fs.is_supervised}
        (output_dir / "info.json").write_text(json.dumps(info, indent=2))
        logger.info("FeatureSet saved to %s", output_dir)

    @staticmethod
    def load(output_dir) -> FeatureSet:
        output_dir = Path(output_dir)
        for p in (output_dir / "features.npy", output_dir / "info.json"):
            if not p.exists():
                raise FileNotFoundError(f"Expected file not found: {p}. "
                                        "Was this directory written by FeaturePipeline.save()?")
        info = json.loads((output_dir / "info.json").read_text())

        def opt_npy(name):
            return np.load(output_dir / name) if (output_dir / name).exists() else None

        def opt_json(name, default):
            return json.loads((output_dir / name).read_text()) if (output_dir / name).exists() else default

        return FeatureSet(features=np.load(output_dir / "features.npy"), feature_type=info["feature_type"],
                          modality=info["modality"], metadata=opt_json("metadata.json", []),
                          labels=opt_npy("labels.npy"), label_names=opt_json("label_names.json", None),
                          cluster_assignments=opt_npy("cluster_assignments.npy"))


_EXP_KEYS = ("extractor", "loader", "name", "dataset", "split", "output", "max_samples", "audio_folder",
             "extractor_params", "manifest", "manifest_split")


def resolve_experiments(cfg: dict) -> list:
    """Top-level keys are defaults, each experiment overrides them; unknown keys are ignored
    (config.py:200-261, 316-329)."""
    top = {k: cfg.get(k) for k in _EXP_KEYS}
    exps = cfg.get("experiments") or [dict()]
    out = []
    for i, e in enumerate(exps):
        m = {}
        for k in _EXP_KEYS:
            v = e.get(k)
            if k == "split":
                m[k] = v if ("split" in e and v is not None) else top[k]
            elif k == "extractor_params":
                m[k] = v if v else (top[k] or {})
            else:
                m[k] = v if v is not None else top[k]
        if not m["extractor"]:
            raise ValueError(f"Experiment #{i} is missing 'extractor'. Set it in the experiment or at the top level.")
        if not m["loader"]:
            raise ValueError(f"Experiment #{i} is missing 'loader'. Set it in the experiment or at the top level.")
        m["name"] = m["name"] or f"{m['loader']}_{m['extractor']}_{m['split']}"
        m["output"] = m["output"] or f"data/processed/{m['name']}"
        out.append(m)
    return out


def build_loader(exp: dict) -> BaseDatasetLoader:
    if exp["loader"] != "audio_folder":
        raise ValueError(f"Unknown loader: {exp['loader']!r}. This package ships 'audio_folder'; inside the "
                         "reference tree use its own loaders with these extractors (INTEGRATION.md).")
    root = exp.get("audio_folder") or exp["dataset"]
    manifest = exp.get("manifest")
    folder_split = None if (manifest or not exp.get("split")) else exp["split"]
    return AudioFolderLoader(root, split=folder_split, manifest=manifest, manifest_split=exp.get("manifest_split"))


def run_experiment(exp: dict, config_path: Optional[Path] = None) -> FeatureSet:
    loader = build_loader(exp)
    extractor = get(exp["extractor"])(**(exp.get("extractor_params") or {}))
    out = Path(exp["output"])
    fs = FeaturePipeline(loader, extractor).run(max_samples=exp.get("max_samples"), output_dir=out)
    FeaturePipeline.save(fs, out)
    if config_path is not None:
        shutil.copy2(config_path, out / "config.yaml")
    print(f"[{exp['name']}] {fs}")
    print(f"  -> {out}")
    return fs


def run_config(path) -> list:
    import yaml
    cfg = yaml.safe_load(Path(path).read_text()) or {}
    exps = resolve_experiments(cfg)
    print(f"Config: {path}  ({len(exps)} experiment(s))")
    res = []
    for e in exps:
        print(f"\nRunning: {e['name']} ...")
        res.append(run_experiment(e, Path(path)))
    print("\nAll experiments complete.")
    return res


def main() -> None:
    ap = argparse.ArgumentParser(description="B200 Stage-2 audio feature extraction")
    ap.add_argument("--config", required=True)
    logging.basicConfig(level=logging.INFO)
    run_config(ap.parse_args().config)


if __name__ == "__main__":
    main()

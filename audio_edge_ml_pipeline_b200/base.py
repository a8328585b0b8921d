"""Host-side mirror of the reference's plugin interface for this path.

Same names, fields, argument meaning and error behaviour as
``src/preprocessing/feature_extraction/base.py`` in the reference:
``FeatureSet`` (:27-134), ``BaseFeatureExtractor`` (:137-234), ``BaseDatasetLoader`` (:237-257).
Written independently; only what the Stage-2 audio path needs (no TensorFlow export).
"""

from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from dataclasses import dataclass
from pathlib import Path
from typing import Iterator, Optional

import numpy as np

logger = logging.getLogger(__name__)


@dataclass
class FeatureSet:
    """Container of extracted features: ``features`` is ``(N, *feature_dims)``; ``labels`` are
    int32 class indices (or None), ``label_names`` maps index -> name."""

    features: np.ndarray
    feature_type: str
    modality: str
    metadata: list
    labels: Optional[np.ndarray] = None
    label_names: Optional[list] = None
    cluster_assignments: Optional[np.ndarray] = None

    @property
    def n_samples(self) -> int:
        return len(self.features)

    @property
    def feature_shape(self) -> tuple:
        return self.features.shape[1:]

    @property
    def is_supervised(self) -> bool:
        return self.labels is not None

    @property
    def n_classes(self) -> Optional[int]:
        if self.label_names is not None:
            return len(self.label_names)
        if self.labels is not None:
            return int(self.labels.max()) + 1
        return None

    def to_sklearn(self):
        if self.labels is not None:
            return self.features, self.labels
        if self.cluster_assignments is not None:
            return self.features, self.cluster_assignments
        return self.features, None

    def __repr__(self) -> str:
        info = f"labels={self.n_classes} classes" if self.is_supervised else "unsupervised"
        return (f"FeatureSet(modality={self.modality!r}, feature_type={self.feature_type!r}, "
                f"n_samples={self.n_samples}, feature_shape={self.feature_shape}, {info})")


class BaseDatasetLoader(ABC):
    """Iterating yields ``(sample_path, label, metadata)``; metadata is splatted into
    ``extract(**metadata)`` (reference base.py:237-257)."""

    @abstractmethod
    def __iter__(self) -> Iterator[tuple]:
        ...

    @abstractmethod
    def __len__(self) -> int:
        ...


class BaseFeatureExtractor(ABC):
    """Class attributes ``name`` / ``feature_type`` / ``modality`` and ``extract``; the serial
    ``extract_dataset`` below is the reference's loop (base.py:176-234), kept as the semantic
    definition that the batched GPU override in extractors.py must reproduce."""

    name: str
    feature_type: str
    modality: str

    @abstractmethod
    def extract(self, sample_path: Optional[Path], **kwargs) -> np.ndarray:
        ...

    def extract_dataset(self, loader: BaseDatasetLoader, max_samples: Optional[int] = None) -> FeatureSet:
        feats, labels, metas, label_to_idx = [], [], [], {}
        for i, (sample_path, label, meta) in enumerate(loader):
            if max_samples is not None and i >= max_samples:
                break
            try:
                feat = self.extract(sample_path, **meta)
            except Exception as exc:  # noqa: BLE001 — reference semantics: warn and skip
                logger.warning("Skipping %s: %s", sample_path, exc)
                continue
            feats.append(feat)
            metas.append(meta)
            if label is not None:
                if label not in label_to_idx:
                    label_to_idx[label] = len(label_to_idx)
                labels.append(label_to_idx[label])
        return assemble_feature_set(self, feats, labels, metas, label_to_idx)


def assemble_feature_set(extractor, feats, labels, metas, label_to_idx) -> FeatureSet:
    """Tail of the reference's extract_dataset (base.py:216-234)."""
    if isinstance(feats, np.ndarray):
        features = feats
    else:
        if not feats:
            raise RuntimeError("No features were successfully extracted.")
        features = np.stack(feats)
    lab = np.array(labels, dtype=np.int32) if labels else None
    names = [k for k, _ in sorted(label_to_idx.items(), key=lambda x: x[1])] if label_to_idx else None
    return FeatureSet(features=features, feature_type=extractor.feature_type, modality=extractor.modality,
                      metadata=metas, labels=lab, label_names=names)

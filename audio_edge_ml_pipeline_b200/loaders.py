"""Class-per-folder audio loader with the iteration contract of the reference's
``AudioFolderLoader`` (``src/preprocessing/dataset_loaders/audio_folder_loader.py:106-232``):
sorted class folders, sorted clips, per-sample metadata ``filename / class_dir / duration /
sample_rate / n_channels`` (header probe, zeros on failure), optional split sub-directory and
``split_manifest.json`` filter.  Inside the reference tree the reference's own loaders are used
unchanged; this one lets the package run stand-alone."""

from __future__ import annotations

import json
import logging
import os
from pathlib import Path
from typing import Iterator, Optional

from .base import BaseDatasetLoader
from .wavio import wav_info

logger = logging.getLogger(__name__)


def _probe_all(paths: list) -> list:
    """``paths``: plain strings.  duration / sample_rate / n_channels of every file (audio_folder_loader.py:76-103: soundfile.info, zeros on
    failure).  RIFF/WAVE headers go through the library's threaded probe in one call (7 000 files: 0.03 s instead of
    0.09 s of Python open / read / parse); anything it does not recognise, or a missing library, takes the Python
    parser, which has the same zero fallback."""
    out = [None] * len(paths)
    try:
        from . import _lib
        wav = [i for i, p in enumerate(paths) if p[-4:].lower() == ".wav" or p[-5:].lower() == ".wave"]
        if wav:
            info = _lib.probe_wav_batch([paths[i] for i in wav])
            for k, i in enumerate(wav):
                if info["status"][k] == 0 and info["rate"][k] > 0:
                    out[i] = {"duration": int(info["n_frames"][k]) / int(info["rate"][k]),
                              "sample_rate": int(info["rate"][k]), "n_channels": int(info["channels"][k])}
    except Exception as exc:  # noqa: BLE001 — library not built / not loadable: the Python parser below
        logger.debug("native header probe unavailable: %s", exc)
    return [o if o is not None else wav_info(p) for o, p in zip(out, paths)]

_AUDIO_SUFFIXES = frozenset({".wav", ".flac", ".ogg", ".mp3", ".aac", ".m4a", ".opus", ".aiff", ".aif"})


class AudioFolderLoader(BaseDatasetLoader):
    def __init__(self, root, split: Optional[str] = None, extensions=None, class_names=None,
                 manifest=None, manifest_split: Optional[str] = None) -> None:
        eff = Path(root) / split if split else Path(root)
        if not eff.is_dir():
            raise NotADirectoryError(f"Dataset root not found: {eff}")
        exts = frozenset(e.lower() for e in extensions) if extensions is not None else _AUDIO_SUFFIXES
        if class_names is not None:
            self._class_names = list(class_names)
            class_dirs = [eff / c for c in class_names]
        else:
            class_dirs = sorted(p for p in eff.iterdir() if p.is_dir())
            self._class_names = [d.name for d in class_dirs]
        self._samples: list = []
        strs: list = []                                          # the same paths as plain strings, for the header probe
        for class_dir, label in zip(class_dirs, self._class_names):
            if not class_dir.is_dir():
                logger.warning("Class directory not found: %s (skipping)", class_dir)
                continue
            # os.scandir: the entry type comes with the directory read, no stat() per file (7 000 files: 10 ms)
            with os.scandir(class_dir) as it:
                names = [e.name for e in it if e.is_file() and os.path.splitext(e.name)[1].lower() in exts]
            names.sort()                                         # (plain names sort like the paths: same parent)
            clips = [class_dir / n for n in names]
            strs.extend(os.path.join(class_dir, n) for n in names)
            if not clips:
                logger.warning("No audio files found in: %s", class_dir)
            for clip in clips:
                self._samples.append((clip, label, {"filename": clip.name, "class_dir": class_dir.name}))
        for (clip, _label, meta), info in zip(self._samples, _probe_all(strs)):
            meta.update(info)
        if manifest is not None:
            if manifest_split is None:
                raise ValueError("manifest_split must be set when manifest is given")
            allowed = set(json.loads(Path(manifest).read_text()).get(manifest_split, []))
            rr = Path(root)
            self._samples = [s for s in self._samples if str(s[0].relative_to(rr)) in allowed]
        logger.info("AudioFolderLoader: %d clips across %d classes.", len(self._samples), len(self._class_names))

    def __len__(self) -> int:
        return len(self._samples)

    def __iter__(self) -> Iterator[tuple]:
        yield from self._samples

    @property
    def class_names(self) -> list:
        return list(self._class_names)

    @property
    def n_classes(self) -> int:
        return len(self._class_names)

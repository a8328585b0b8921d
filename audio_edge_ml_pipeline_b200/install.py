"""Adapter: put the B200 extractors behind the reference's own registry.

Inside the reference tree ``@register`` refuses duplicates (registry.py:59-64), so the three names
are taken over by replacing the registry entries after the reference package has been imported:

    import src.preprocessing.feature_extraction as fx          # the reference
    from audio_edge_ml_pipeline_b200.install import install_into_reference
    install_into_reference(fx.registry)                         # audio_mel_spec -> B200 class
    # python -m src.preprocessing.pipeline --config config/feature_extraction.yaml  now runs on the GPU
"""

from __future__ import annotations

from . import extractors

NAMES = ("audio_mel_spec", "audio_mfcc_seq", "audio_cqt")


def install_into_reference(registry_module) -> dict:
    """``registry_module`` is the reference's ``...feature_extraction.registry`` module.
    Returns the classes that were replaced (for restoring)."""
    reg = registry_module._REGISTRY
    old = {}
    for cls in (extractors.AudioMelSpectrogram, extractors.AudioMFCCSequence, extractors.AudioCQT,
                extractors.AudioClassicalExtractor):
        old[cls.name] = reg.get(cls.name)
        reg[cls.name] = cls
    return old


def restore(registry_module, old: dict) -> None:
    for name, cls in old.items():
        if cls is None:
            registry_module._REGISTRY.pop(name, None)
        else:
            registry_module._REGISTRY[name] = cls

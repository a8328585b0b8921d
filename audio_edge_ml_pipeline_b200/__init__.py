"""B200-native Stage-2 audio feature extraction (audio_mel_spec / audio_mfcc_seq / audio_cqt).

Importing the package needs neither the built library nor a GPU; computing anything does.
"""

from .base import BaseDatasetLoader, BaseFeatureExtractor, FeatureSet
from .extractors import AudioCQT, AudioMelSpectrogram, AudioMFCCSequence
from . import loaders
from .pipeline import FeaturePipeline
from .registry import get, list_extractors, register

__all__ = ["AudioCQT", "AudioMFCCSequence", "AudioMelSpectrogram", "BaseDatasetLoader", "BaseFeatureExtractor",
           "FeaturePipeline", "FeatureSet", "get", "list_extractors", "register"]

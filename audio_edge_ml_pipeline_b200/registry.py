"""Extractor registry with the reference's semantics
(``src/preprocessing/feature_extraction/registry.py:39-87``): ``register`` keys on ``cls.name``,
raises TypeError without a string name and ValueError on duplicates; ``get`` raises KeyError."""

from __future__ import annotations

from typing import Type

from .base import BaseFeatureExtractor

_REGISTRY: dict = {}


def register(cls: Type[BaseFeatureExtractor]) -> Type[BaseFeatureExtractor]:
    if not hasattr(cls, "name") or not isinstance(cls.name, str):
        raise TypeError(f"{cls.__qualname__} must define a string class attribute 'name'.")
    if cls.name in _REGISTRY:
        raise ValueError(f"An extractor named '{cls.name}' is already registered "
                         f"({_REGISTRY[cls.name].__qualname__}). Use a unique name or remove the duplicate.")
    _REGISTRY[cls.name] = cls
    return cls


def get(name: str) -> Type[BaseFeatureExtractor]:
    if name not in _REGISTRY:
        raise KeyError(f"No extractor named '{name!r}'. Available extractors: {list_extractors()}")
    return _REGISTRY[name]


def list_extractors() -> list:
    return sorted(_REGISTRY)

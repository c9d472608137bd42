"""Callers either side of the hot path: `FluteLinear` / `prepare_model_flute`."""
from .base import FluteLinear, prepare_model_flute   # noqa: F401

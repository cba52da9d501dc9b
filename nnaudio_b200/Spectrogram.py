"""Legacy import path, mirroring ``nnAudio/Spectrogram.py:1-8`` (which re-exports
``nnAudio.features`` and emits a deprecation ``Warning``)."""
import warnings

from .features import *  # noqa: F401,F403

warnings.warn(
    "importing from Spectrogram is deprecated; use `from nnaudio_b200 import features`",
    category=Warning,
)

"""nnAudio-compatible feature modules backed by ``libnnab.so``.

Public names mirror ``nnAudio.features`` for the hot-path classes of
SURVEY.md §8 (features/__init__.py:6-14 in the reference)."""
from .stft import STFT, iSTFT
from .mel import MelSpectrogram, MFCC
from .gammatone import Gammatonegram
from .cqt import CQT1992v2, CQT2010v2, CQT
from .vqt import VQT
from .cqt_v1 import CQT1992, CQT2010
from .griffin_lim import Griffin_Lim
from .cfp import Combined_Frequency_Periodicity, CFP

__all__ = ["STFT", "iSTFT", "MelSpectrogram", "MFCC", "Gammatonegram", "CQT1992v2", "CQT2010v2", "CQT", "VQT",
           "CQT1992", "CQT2010", "Griffin_Lim", "Combined_Frequency_Periodicity", "CFP"]

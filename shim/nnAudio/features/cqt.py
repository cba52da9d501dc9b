# nnAudio/features/cqt.py holds both generations of the CQT modules
from nnaudio_b200.features.cqt import CQT, CQT1992v2, CQT2010v2  # noqa: F401
from nnaudio_b200.features.cqt_v1 import CQT1992, CQT2010  # noqa: F401

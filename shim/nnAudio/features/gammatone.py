from nnaudio_b200.features.gammatone import Gammatonegram  # noqa: F401  (nnAudio/features/gammatone.py)

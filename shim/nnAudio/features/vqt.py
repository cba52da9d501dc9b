from nnaudio_b200.features.vqt import VQT  # noqa: F401  (nnAudio/features/vqt.py)

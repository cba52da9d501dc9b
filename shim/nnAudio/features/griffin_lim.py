from nnaudio_b200.features.griffin_lim import Griffin_Lim  # noqa: F401  (nnAudio/features/griffin_lim.py)

from nnaudio_b200.features.stft import STFT, iSTFT  # noqa: F401  (nnAudio/features/stft.py)

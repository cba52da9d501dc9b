from nnaudio_b200.features.mel import MFCC, MelSpectrogram  # noqa: F401  (nnAudio/features/mel.py)

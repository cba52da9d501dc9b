from nnaudio_b200.features.cfp import CFP, Combined_Frequency_Periodicity  # noqa: F401  (nnAudio/features/cfp.py)

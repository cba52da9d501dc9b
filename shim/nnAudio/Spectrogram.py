from nnaudio_b200.Spectrogram import *  # noqa: F401,F403  (emits the reference's deprecation Warning)

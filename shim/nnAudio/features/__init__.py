from nnaudio_b200.features import *  # noqa: F401,F403
from nnaudio_b200.features import __all__  # noqa: F401

"""Shim: keeps the reference import path `src.utils.KerasCallbacks` working."""
from cmr_landmark_detection_b200.utils.KerasCallbacks import *  # noqa: F401,F403
from cmr_landmark_detection_b200.utils.KerasCallbacks import get_callbacks  # noqa: F401,E402

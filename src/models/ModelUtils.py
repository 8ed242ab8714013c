from cmr_landmark_detection_b200.models.ModelUtils import *  # noqa: F401,F403

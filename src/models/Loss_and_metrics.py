from cmr_landmark_detection_b200.models.Loss_and_metrics import *  # noqa: F401,F403

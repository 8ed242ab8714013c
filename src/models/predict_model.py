from cmr_landmark_detection_b200.models.predict_model import *  # noqa: F401,F403

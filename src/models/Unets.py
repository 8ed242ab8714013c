from cmr_landmark_detection_b200.models.Unets import *  # noqa: F401,F403
from cmr_landmark_detection_b200.models.Unets import create_unet, get_model  # noqa: F401,E402

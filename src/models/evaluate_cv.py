from cmr_landmark_detection_b200.models.evaluate_cv import *  # noqa: F401,F403
from cmr_landmark_detection_b200.models.evaluate_cv import get_ip_from_rvip_mask_3d, get_mean_rvip_2d  # noqa: F401,E402

"""Shim: keeps the reference import path `src.models.predict_model` working (pred_fold on the B200 path)."""
from cmr_landmark_detection_b200.models.predict_model import pred_fold, predict_label_volume  # noqa: F401

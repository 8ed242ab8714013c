"""Shim: keeps the reference import path `src.data.Postprocess` working (clean_3d_prediction_2d_cc on the device)."""
from cmr_landmark_detection_b200.data.Postprocess import clean_3d_prediction_2d_cc  # noqa: F401

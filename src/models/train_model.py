"""Shim: keeps the reference import path `src.models.train_model` working (train_fold on the B200 path)."""
from cmr_landmark_detection_b200.models.train_model import train_fold  # noqa: F401

"""The compute steps of src/models/predict_model.py:pred_fold (model.predict :143, threshold -> label
map :149-156) on the device. pred_fold's file handling (SimpleITK .nrrd I/O, undo_generator_steps) is
out of scope and stays with the caller."""
from __future__ import annotations

import numpy as np
import torch

from ..extract import label_map_device


def predict_label_volume(model, x: np.ndarray, thr: float = 0.5, batch_size: int = 1) -> np.ndarray:
    """preds = model.predict(x); preds_flat[preds[...,0]>thr]=1; preds_flat[preds[...,1]>thr]=2 -> uint8 [N,H,W]
    (predict_model.py:143-156, :170-171). pred_fold predicts with BATCHSIZE=1 (:89)."""
    preds = model.predict(np.asarray(x, np.float32), batch_size=batch_size)
    heat = torch.from_numpy(preds).to(model.device)
    return label_map_device(heat, thr).cpu().numpy()

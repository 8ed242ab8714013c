"""Stub-import shim that lets the reference's numpy-only landmark functions be imported in the
BUILD container (tensorflow / SimpleITK / skimage / albumentations / matplotlib are absent).
Used ONLY by the golden generators in this directory; /root/reference does not exist on the GPU
box, so nothing under tests/ imports this at test time."""
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = '/root/reference'


def import_reference_eval():
    for m in ['SimpleITK', 'skimage', 'skimage.exposure', 'skimage.measure', 'skimage.transform',
              'albumentations', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.patches',
              'matplotlib.transforms', 'matplotlib.ticker', 'yaml', 'seaborn', 'tensorflow',
              'tensorflow.keras', 'tensorflow.keras.backend']:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.models.evaluate_cv as ev   # noqa
    return ev

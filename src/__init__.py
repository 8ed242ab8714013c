"""Import shim: `from src.models.Unets import create_unet` -- the reference's module path
(train_model.py:38, predict_model.py:12) -- resolves to the B200-native implementation."""

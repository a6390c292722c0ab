"""The upload-hook walker lives with the product's boundary (pflare_b200/upload.py); re-exported
here so the oracle and the CUDA path are fed by the very same code."""
from pflare_b200.upload import feed, AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE  # noqa: F401

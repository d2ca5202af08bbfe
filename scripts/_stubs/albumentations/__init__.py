"""Import stub: albumentations is not installed in this image and the synthetic dataset does not use it."""

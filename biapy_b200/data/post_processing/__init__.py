"""Post-processing steps on the device: test-time augmentation (``biapy/data/post_processing/``)."""

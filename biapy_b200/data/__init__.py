"""Host-side mirror of ``biapy.data`` for the hot path: patch bookkeeping, crop and overlap-add merge."""

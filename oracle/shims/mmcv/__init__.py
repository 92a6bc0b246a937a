"""Import shim (test infrastructure) for ``mmcv.cnn`` (core/segformer_head.py:11)."""

"""Import shim (test infrastructure): the subset of ``timm`` that the reference's
core/mix_transformer.py:11 imports.  Off the fusion hot path (SegFormer consumer)."""

"""Import shim (test infrastructure): test_original.py:666 imports matplotlib.pyplot."""

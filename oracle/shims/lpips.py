"""Import shim (test infrastructure): attack/attack.py:6 imports ``lpips``; only the
unused perceptual-loss attack variants touch it."""


class LPIPS:
    def __init__(self, *a, **k):
        raise NotImplementedError("lpips is off the evaluated path")

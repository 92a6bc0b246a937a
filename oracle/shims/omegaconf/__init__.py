"""Import shim (test infrastructure): ``OmegaConf.load`` as used at robust_test.py:48
(only ``cfg.exp.backbone`` is read, :262)."""
import yaml


class _Node(dict):
    def __getattr__(self, k):
        v = self[k]
        return _Node(v) if isinstance(v, dict) else v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _Node(yaml.safe_load(f))

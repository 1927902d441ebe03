"""Test-only stand-in: the reference only does isinstance(x, DictConfig) (detectron2/config.py:872)."""


class DictConfig(dict):
    pass

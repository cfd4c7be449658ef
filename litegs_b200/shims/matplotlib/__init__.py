"""``matplotlib`` stand-in: the reference imports ``matplotlib.pyplot`` at module import time (trainer.py:10) and never
calls it on the training path.  Any attribute resolves to a no-op so that stray debugging calls do not crash a run."""
__version__ = "0.0-litegs_b200-shim"


def use(*_a, **_k):
    return None

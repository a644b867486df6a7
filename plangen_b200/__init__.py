"""plangen_b200 - B200-native (sm_100a) implementation of PlanGen's CFG image-token decode path."""
from .config import Dims, JANUS_1P3B, JANUS_7B  # noqa: F401

__all__ = ["Dims", "JANUS_1P3B", "JANUS_7B", "FastJanus"]


def __getattr__(name):
    if name == "FastJanus":
        from .engine import FastJanus
        return FastJanus
    raise AttributeError(name)

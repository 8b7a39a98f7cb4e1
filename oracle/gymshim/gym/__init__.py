"""Minimal stand-in for OpenAI gym 0.21 -- TEST INFRASTRUCTURE ONLY.

The reference simulator (/root/reference/mate) imports ``gym`` which is not installed
in this image.  This package provides exactly the surface ``mate/`` touches so that the
UNMODIFIED reference can be imported by ``oracle/gen_golden.py`` to produce golden
fixtures.  It is never imported by the product package (``mate_b200``).
"""
import numpy as np

if not hasattr(np, 'bool8'):  # removed in NumPy 2; the reference uses it
    np.bool8 = np.bool_

__version__ = '0.21.0'

from gym import logger, spaces, utils  # noqa: E402
from gym.core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: E402

_registry = {}


class _Spec:
    def __init__(self, id):
        self.id = id


def register(id, entry_point=None, kwargs=None, **_unused):
    _registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    entry_point, defaults = _registry[id]
    merged = dict(defaults)
    merged.update(kwargs)
    env = entry_point(**merged)
    env.unwrapped.spec = _Spec(id)
    return env

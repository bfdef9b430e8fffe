"""TEST INFRASTRUCTURE -- minimal stand-in for the `gym` package (not installed in this
image) so that /root/reference/mrsgym imports verbatim.  Only what the reference's hot
path touches: gym.Env, gym.spaces.Box, gym.envs.registration.register, gym.make."""
import importlib
from . import spaces  # noqa: F401
from .envs import registration  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass


def make(env_id, **kwargs):
    entry = registration.REGISTRY[env_id]
    mod, cls = entry.split(':')
    return getattr(importlib.import_module(mod), cls)(**kwargs)

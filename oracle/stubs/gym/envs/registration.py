"""TEST INFRASTRUCTURE -- gym.envs.registration stub (see gym/__init__.py)."""
REGISTRY = {}


def register(id, entry_point, **kwargs):
    REGISTRY[id] = entry_point

"""RLlib-style adapters over the batched MRS.

Counterparts of /root/reference/mrsgym/MRSWrapper.py:11-55 without the `ray` dependency (the classes
are duck-typed: flat numpy observations, and a dict-per-agent view).  Single-env semantics like the
reference (N_ENVS = 1); for batched training use MRS directly.
"""
from __future__ import annotations

import numpy as np
import torch

from .env import MRS
from .spaces import Box


class MRS_RLlib(MRS):
    """Flat numpy observations (MRSWrapper.py:11-26): obs = X.reshape(-1).numpy(); config["action_fn"]
    transforms the given actions before the step."""

    def __init__(self, config=None):
        config = dict(config or {})
        self.action_fn = config.pop('action_fn', lambda action: action)
        self._constructing = True
        super().__init__(**config)
        self._constructing = False

    def step(self, actions):
        obs, reward, done, info = super().step(self.action_fn(actions))
        return obs.reshape(-1).cpu().numpy(), reward, done, info

    def reset(self, **kwargs):
        obs = super().reset(**kwargs)
        if getattr(self, '_constructing', False):
            return obs
        return obs.reshape(-1).cpu().numpy()


class MRS_RLlib_MultiAgent(MRS):
    """Dict-per-agent view (MRSWrapper.py:30-55): observations {"agent1": x_1 (D,), ...} of the newest
    slice, actions given as {"agentK": action}; agents without an entry get a zero action."""

    defaults = {}

    def __init__(self, config=None):
        params = dict(MRS_RLlib_MultiAgent.defaults)
        params.update(config or {})
        self.action_fn = params.pop('action_fn', lambda action: action)
        super().__init__(**params)
        # per-agent spaces (MRSWrapper.py:37-38 indexes the flat boxes of MRS as if they were 3-D and would raise;
        # the evident intent -- one agent's observation and action box -- is what is built here)
        self.observation_space = Box(np.full((self.STATE_SIZE,), -np.inf, dtype=np.float32),
                                     np.full((self.STATE_SIZE,), np.inf, dtype=np.float32))
        self.action_space = Box(self.action_space.low[:4][:max(self.swarm.action_dim, 1)],
                                self.action_space.high[:4][:max(self.swarm.action_dim, 1)])
        names = ['agent%d' % (i + 1) for i in range(self.N_AGENTS)]
        self.names_dict = {n: i for i, n in enumerate(names)}
        self.env.names_dict = self.names_dict

    def _obs_dict(self, Xk):
        x = Xk[0].cpu().numpy()            # newest slice (N, D)
        return {n: np.array(x[i]) for n, i in self.names_dict.items()}

    def step(self, actions):
        arr = np.zeros((self.N_AGENTS, self.swarm.action_dim), dtype=np.float32)
        for name, action in actions.items():
            arr[self.names_dict[name], :] = np.asarray(self.action_fn(action), dtype=np.float32)
        Xk, reward, done, info = super().step(torch.from_numpy(arr))
        return self._obs_dict(Xk), reward, done, info

    def reset(self, **kwargs):
        Xk = super().reset(**kwargs)
        if not hasattr(self, 'names_dict'):
            return Xk
        return self._obs_dict(Xk)

"""NormalizedBoxEnv for the batched environment (ast_sac/env_wrapper/normalized_box_env.py:14-61).

Same arithmetic as the reference wrapper, on torch tensors: the policy's action in [-1, 1] is mapped
to [lb, ub] in float32 (``lb + (a + 1) * 0.5 * (ub - lb)``, then clipped) and the reward is multiplied
by ``reward_scale``.  Attribute access falls through to the wrapped env like ProxyEnv
(ast_sac/env_wrapper/proxy_env.py:30-33).
"""
from __future__ import annotations

import numpy as np
import torch


class NormalizedBoxEnv:
    def __init__(self, env, reward_scale: float = 1.0):
        self._wrapped_env = env
        self._reward_scale = reward_scale
        self.action_space = _unit_box(env.action_space)
        self.observation_space = env.observation_space

    @property
    def wrapped_env(self):
        return self._wrapped_env

    def scale_action(self, action):
        """[-1, 1] -> [lb, ub], float32 like the reference (Box bounds are float32)."""
        space = self._wrapped_env.action_space
        if isinstance(action, torch.Tensor):
            lb = torch.as_tensor(space.low, dtype=torch.float32, device=action.device)
            ub = torch.as_tensor(space.high, dtype=torch.float32, device=action.device)
            a = action.to(torch.float32)
            scaled = lb + (a + 1.0) * 0.5 * (ub - lb)
            return torch.minimum(torch.maximum(scaled, lb), ub)
        lb, ub = space.low, space.high
        scaled = lb + (np.asarray(action, dtype=np.float32) + np.float32(1.0)) * np.float32(0.5) * (ub - lb)
        return np.clip(scaled, lb, ub)

    def step(self, action):
        next_obs, reward, done, info = self._wrapped_env.step(self.scale_action(action))
        return next_obs, reward * self._reward_scale, done, info

    def reset(self, *a, **k):
        return self._wrapped_env.reset(*a, **k)

    def __getattr__(self, attr):
        if attr == '_wrapped_env':
            raise AttributeError()
        return getattr(self._wrapped_env, attr)

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __str__(self):
        return "Normalized: %s" % self._wrapped_env


def _unit_box(space):
    ub = np.ones(space.shape, dtype=np.float32)
    return type(space)(-1 * ub, ub, dtype=np.float32)

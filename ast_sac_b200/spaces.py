"""``Box`` space: gymnasium's when it is installed, otherwise a minimal stand-in exposing what the
reference's consumers read (``low``, ``high``, ``shape``, ``dtype``, ``sample``;
ast_sac/env_wrapper/normalized_box_env.py:33,48-49, ast_sac/env_wrapper/env_utils.py:13)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not installed in the build image
    from gymnasium.spaces import Box  # type: ignore
except Exception:
    class Box:  # type: ignore
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.low = np.asarray(low, dtype=self.dtype)
            self.high = np.asarray(high, dtype=self.dtype)
            if shape is not None:
                self.low = np.broadcast_to(self.low, shape).copy()
                self.high = np.broadcast_to(self.high, shape).copy()
            self.shape = self.low.shape
            self._rng = np.random.default_rng()

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

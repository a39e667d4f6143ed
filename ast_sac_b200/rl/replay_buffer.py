"""GPU-resident replay buffer.

Mirror of SimpleReplayBuffer / EnvReplayBuffer (ast_sac/data_management/simple_replay_buffer.py:8-103,
env_replay_buffer.py:8-50): a ring of (observation, action, reward, terminal, next_observation) with
uniform sampling with replacement -- but the storage is torch tensors on the env's device, whole
batches of transitions are appended without a host synchronisation, and ``random_batch`` returns
device tensors, so neither the rollout nor the SAC update crosses PCIe (the reference copies every
batch host -> device in np_to_pytorch_batch, ast_sac/torch/core/module.py:67-75).

Semantics kept from the reference: ring order = insertion order, ``_top`` / ``_size`` bookkeeping,
terminals stored as uint8 (simple_replay_buffer.py:31), env_info keys dropped unless sizes are given,
``random_batch`` samples indices in [0, size) with replacement.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np
import torch


class GpuReplayBuffer:
    def __init__(self, max_replay_buffer_size: int, env=None, observation_dim: Optional[int] = None,
                 action_dim: Optional[int] = None, device=None, seed: Optional[int] = None,
                 dtype=torch.float32):
        if env is not None:
            observation_dim = int(np.prod(env.observation_space.shape)) if observation_dim is None else observation_dim
            action_dim = int(np.prod(env.action_space.shape)) if action_dim is None else action_dim
            if device is None:
                device = getattr(env, "_device", None)
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        N = int(max_replay_buffer_size)
        self._max_replay_buffer_size = N
        self._observation_dim, self._action_dim = observation_dim, action_dim
        # one spare row (index N) swallows the masked-out rows of add_batch
        self._observations = torch.zeros((N + 1, observation_dim), dtype=dtype, device=self.device)
        self._next_obs = torch.zeros((N + 1, observation_dim), dtype=dtype, device=self.device)
        self._actions = torch.zeros((N + 1, action_dim), dtype=dtype, device=self.device)
        self._rewards = torch.zeros((N + 1, 1), dtype=dtype, device=self.device)
        self._terminals = torch.zeros((N + 1, 1), dtype=torch.uint8, device=self.device)
        # ring position and fill level live on the device so that appends never synchronise
        self._top_dev = torch.zeros((), dtype=torch.int64, device=self.device)
        self._size_dev = torch.zeros((), dtype=torch.int64, device=self.device)
        # a private generator only when a seed is asked for; the default generator is the one CUDA graph capture
        # knows how to advance between replays
        self._gen = None
        if seed is not None:
            self._gen = torch.Generator(device=self.device)
            self._gen.manual_seed(seed)

    # -- reference API (single samples / paths; host values accepted) -------------------------------
    @property
    def _top(self) -> int:
        return int(self._top_dev.item())

    @property
    def _size(self) -> int:
        return int(self._size_dev.item())

    def add_sample(self, observation, action, reward, next_observation, terminal, env_info=None, **kwargs):
        t = lambda x, d: torch.as_tensor(np.asarray(x), device=self.device).to(d).reshape(1, -1)
        f = self._observations.dtype
        self.add_batch(t(observation, f), t(action, f), t(reward, f), t(next_observation, f),
                       t(terminal, torch.uint8))

    def terminate_episode(self):
        pass

    def add_path(self, path):
        """replay_buffer.py:34-73: every step of a path dict (rollout_functions.py:172-181), in order."""
        f = self._observations.dtype
        t = lambda x, d: torch.as_tensor(np.asarray(x), device=self.device).to(d)
        n = len(path["actions"])
        self.add_batch(t(path["observations"], f).reshape(n, -1), t(path["actions"], f).reshape(n, -1),
                       t(path["rewards"], f).reshape(n, 1), t(path["next_observations"], f).reshape(n, -1),
                       t(path["terminals"], torch.uint8).reshape(n, 1))
        self.terminate_episode()

    def add_paths(self, paths):
        for path in paths:
            self.add_path(path)

    def clear(self):
        self._top_dev.zero_()
        self._size_dev.zero_()

    def num_steps_can_sample(self) -> int:
        return self._size

    def get_diagnostics(self):
        return OrderedDict([('size', self._size)])

    def get_snapshot(self):
        return {}

    def end_epoch(self, epoch):
        return

    # -- batched, synchronisation-free append --------------------------------------------------------
    def add_batch(self, observations, actions, rewards, next_observations, terminals, mask=None):
        """Append the rows selected by ``mask`` (all if None), in row order, at the ring position.
        Everything stays on the device; no host synchronisation."""
        n_rows = observations.shape[0]
        N = self._max_replay_buffer_size
        if n_rows > N:
            raise ValueError("a single batch larger than the replay buffer would overwrite itself")
        if mask is None:
            pos = torch.arange(n_rows, device=self.device)
            count = torch.tensor(n_rows, dtype=torch.int64, device=self.device)
            dest = (self._top_dev + pos) % N
        else:
            m = mask.to(device=self.device, dtype=torch.bool).reshape(-1)
            pos = torch.cumsum(m.to(torch.int64), 0) - 1
            count = m.sum()
            dest = torch.where(m, (self._top_dev + pos) % N, torch.full_like(pos, N))
        f = self._observations.dtype
        self._observations.index_copy_(0, dest, observations.to(f).reshape(n_rows, -1))
        self._actions.index_copy_(0, dest, actions.to(f).reshape(n_rows, -1))
        self._rewards.index_copy_(0, dest, rewards.to(f).reshape(n_rows, 1))
        self._next_obs.index_copy_(0, dest, next_observations.to(f).reshape(n_rows, -1))
        self._terminals.index_copy_(0, dest, terminals.to(torch.uint8).reshape(n_rows, 1))
        self._top_dev.copy_((self._top_dev + count) % N)
        self._size_dev.copy_(torch.clamp(self._size_dev + count, max=N))

    def random_batch(self, batch_size: int):
        """simple_replay_buffer.py:70-84: uniform indices in [0, size) with replacement; tensors on the
        device, terminals as float (what np_to_pytorch_batch hands to SACTrainer.train_from_torch)."""
        u = torch.rand((batch_size,), device=self.device, generator=self._gen)
        idx = torch.clamp((u * self._size_dev.to(torch.float32)).to(torch.int64), max=self._max_replay_buffer_size - 1)
        idx = torch.minimum(idx, torch.clamp(self._size_dev - 1, min=0))
        return dict(
            observations=self._observations[idx],
            actions=self._actions[idx],
            rewards=self._rewards[idx],
            terminals=self._terminals[idx].to(self._observations.dtype),
            next_observations=self._next_obs[idx],
        )

"""Vectorised sampler.

Batched siblings of ``ast_sac_rollout`` (ast_sac/samplers/data_collector/rollout_functions.py:74-181) and
``MdpPathCollector`` (ast_sac/samplers/data_collector/path_collector.py:36-102).  The reference collects
one environment at a time with a host round trip per action (ast_sac/torch/sac/policies/base.py:24-37);
here one policy forward serves all environments, observations / actions / rewards never leave the
device, and the transitions go straight into the GPU replay buffer.

Per-environment semantics are the reference's: ``o = env.reset()``; up to ``max_path_length`` times
``a = policy(o)``, ``next_o, r, done, info = env.step(a)``, record (o, a, r, next_o, terminal =
info['terminal'], done), stop after ``done``.  ``paths()`` rebuilds the reference's list of path dicts.
"""
from __future__ import annotations

from collections import OrderedDict, deque
from typing import Optional

import numpy as np
import torch


class RolloutBatch:
    """Transitions of one batched rollout: tensors [T, B, ...] plus the validity mask [T, B]
    (environment b contributed a transition at step t iff it was not done before step t)."""

    def __init__(self, observations, actions, rewards, next_observations, terminals, dones, valid, events):
        self.observations, self.actions, self.rewards = observations, actions, rewards
        self.next_observations, self.terminals, self.dones = next_observations, terminals, dones
        self.valid, self.events = valid, events

    @property
    def num_steps(self) -> int:
        return int(self.valid.sum().item())

    def path_lengths(self) -> torch.Tensor:
        return self.valid.sum(dim=0)

    def returns(self) -> torch.Tensor:
        return (self.rewards.squeeze(-1) * self.valid).sum(dim=0)

    def paths(self, event_to_string=None):
        """List of per-environment path dicts in the reference's schema (rollout_functions.py:172-181)."""
        v = self.valid.cpu().numpy()
        o, a, r = self.observations.cpu().numpy(), self.actions.cpu().numpy(), self.rewards.cpu().numpy()
        no, t, d = self.next_observations.cpu().numpy(), self.terminals.cpu().numpy(), self.dones.cpu().numpy()
        ev = self.events.cpu().numpy()
        out = []
        for b in range(v.shape[1]):
            n = int(v[:, b].sum())
            if n == 0:
                continue
            infos = [dict(events=(event_to_string(int(ev[i, b])) if event_to_string else int(ev[i, b])),
                          terminal=bool(t[i, b, 0])) for i in range(n)]
            out.append(dict(observations=o[:n, b], actions=a[:n, b], rewards=r[:n, b].reshape(-1, 1),
                            next_observations=no[:n, b], terminals=t[:n, b].reshape(-1, 1).astype(bool),
                            dones=d[:n, b].reshape(-1, 1).astype(bool), agent_infos=[{} for _ in range(n)],
                            env_infos=infos))
        return out


@torch.no_grad()
def batched_ast_sac_rollout(env, agent, max_path_length: int, replay_buffer=None, active: Optional[torch.Tensor] = None,
                            deterministic: bool = False) -> RolloutBatch:
    """One episode of every (active) environment.  ``env`` is a batched env (or NormalizedBoxEnv around
    one) returning device tensors; ``agent.get_actions(obs[B, obs_dim]) -> actions[B, action_dim]``."""
    if hasattr(agent, "reset"):
        agent.reset()
    obs = env.reset().clone()                      # the env returns a view of its output buffer
    B = obs.shape[0]
    dev = obs.device
    alive = torch.ones(B, dtype=torch.bool, device=dev) if active is None else active.to(dev).clone()
    if active is not None and hasattr(env, "set_done"):
        env.set_done(~alive)                       # environments this wave does not need are parked, not simulated
    O, A, R, NO, T, D, V, E = [], [], [], [], [], [], [], []
    for _ in range(int(max_path_length)):
        if O and not bool(alive.any()):            # every active environment is done: no full-width steps for nobody
            break
        a = agent.get_actions(obs, deterministic=deterministic) if deterministic else agent.get_actions(obs)
        next_obs, r, done, info = env.step(a)
        next_obs = next_obs.clone()
        r = r.to(torch.float32).reshape(B, 1).clone()
        term = info['terminal'].reshape(B, 1).clone()
        done = done.reshape(B, 1).clone()
        if replay_buffer is not None:
            replay_buffer.add_batch(obs, a, r, next_obs, term, mask=alive)
        O.append(obs); A.append(a.reshape(B, -1).to(torch.float32)); R.append(r); NO.append(next_obs)
        T.append(term); D.append(done); V.append(alive.clone()); E.append(info['events'].clone())
        alive = alive & ~done.reshape(B)
        obs = next_obs
    st = torch.stack
    return RolloutBatch(st(O), st(A), st(R), st(NO), st(T), st(D), st(V), st(E))


class VectorizedPathCollector:
    """MdpPathCollector for the batched env.  ``collect_new_steps`` runs whole-batch episodes until at
    least ``num_steps`` transitions were collected (the last wave only activates as many environments
    as are still needed, so the overshoot is below one episode per environment).

    Differences from the reference's collector (path_collector.py:36-75), which steps one environment at a time:
    every path runs to its end -- the last one is not truncated to the remaining step budget
    (``max_path_length_this_loop``) and ``discard_incomplete_paths`` has nothing to discard -- so a loop can add up
    to one episode more than ``num_steps`` transitions to the replay buffer; and the transitions of a wave enter the
    buffer time-interleaved (step t of every environment, then step t + 1), not path after path."""

    def __init__(self, env, policy, replay_buffer=None, max_num_epoch_paths_saved=None, deterministic=False,
                 save_env_in_snapshot=True):
        self._env, self._policy = env, policy
        self._replay_buffer = replay_buffer
        self._deterministic = deterministic
        self._epoch_batches = deque(maxlen=max_num_epoch_paths_saved)
        self._num_steps_total = 0
        self._num_paths_total = 0
        self._save_env_in_snapshot = save_env_in_snapshot

    def collect_new_steps(self, max_path_length, num_steps, discard_incomplete_paths=False):
        B = self._env.num_envs
        dev = self._env.obs_buf.device if hasattr(self._env, "obs_buf") else None
        collected, batches = 0, []
        while collected < num_steps:
            need_envs = min(B, -(-(num_steps - collected) // max(1, int(max_path_length))))
            active = torch.arange(B, device=dev) < need_envs
            rb = batched_ast_sac_rollout(self._env, self._policy, max_path_length, self._replay_buffer, active,
                                         self._deterministic)
            n = rb.num_steps
            collected += n
            self._num_paths_total += int((rb.path_lengths() > 0).sum().item())
            batches.append(rb)
            if n == 0:
                break
        self._num_steps_total += collected
        self._epoch_batches.extend(batches)
        return batches

    # reference-compatible name: returns path dicts
    def collect_new_paths(self, max_path_length, num_steps, discard_incomplete_paths=False):
        from ..env import events_to_string
        paths = []
        for rb in self.collect_new_steps(max_path_length, num_steps, discard_incomplete_paths):
            paths.extend(rb.paths(events_to_string))
        return paths

    def get_epoch_paths(self):
        return self._epoch_batches

    def end_epoch(self, epoch):
        self._epoch_batches = deque(maxlen=self._epoch_batches.maxlen)

    def get_diagnostics(self):
        """path_collector.py:83-94: totals and the path-length statistics of this epoch's paths."""
        from .logging import create_stats_ordered_dict
        stats = OrderedDict([('num steps total', self._num_steps_total), ('num paths total', self._num_paths_total)])
        if self._epoch_batches:
            lens = torch.cat([rb.path_lengths()[rb.valid[0]] for rb in self._epoch_batches])
            stats.update(create_stats_ordered_dict('path length', lens))
        return stats

    def get_generic_path_information(self, stat_prefix=''):
        """eval_util.get_generic_path_information (eval_util.py:12-66) over this epoch's paths, computed from the
        batched rollouts on the device: Rewards / Returns / Actions statistics, 'Num Paths', 'Average Returns'.  (The
        env_infos of these envs hold a string and three booleans, which that function skips as non-numeric.)"""
        from .logging import create_stats_ordered_dict
        stats = OrderedDict()
        if not self._epoch_batches:
            return stats
        rew = torch.cat([rb.rewards.squeeze(-1)[rb.valid] for rb in self._epoch_batches])
        act = torch.cat([rb.actions[rb.valid].reshape(-1) for rb in self._epoch_batches])
        ret = torch.cat([rb.returns()[rb.valid[0]] for rb in self._epoch_batches])
        stats.update(create_stats_ordered_dict('Rewards', rew, stat_prefix=stat_prefix))
        stats.update(create_stats_ordered_dict('Returns', ret, stat_prefix=stat_prefix))
        stats.update(create_stats_ordered_dict('Actions', act, stat_prefix=stat_prefix))
        stats['Num Paths'] = int(ret.numel())
        stats[stat_prefix + 'Average Returns'] = float(ret.to(torch.float64).mean())
        return stats

    def get_snapshot(self):
        snap = dict(policy=self._policy)
        if self._save_env_in_snapshot:
            snap['env'] = self._env
        return snap

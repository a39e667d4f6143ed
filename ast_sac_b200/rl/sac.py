"""Soft Actor-Critic on stock PyTorch: the consumer of the batched env for BASELINE config 5.

Not a kernel target (north_star: "the SAC trainer [is] left on stock PyTorch/cuBLAS"): five small MLPs,
9 -> 256 -> 256 -> 1.  The update follows the reference trainer term by term
(ast_sac/torch/sac/sac.py:102-262): automatic entropy tuning, twin Q with soft target update, the
action-regularisation term and the Q-target clipping the reference added, reward_scale applied in the
target; networks and initialisation follow ast_sac/torch/networks/mlp.py:13-74 and
ast_sac/torch/sac/policies/gaussian_policy.py:68-126 (fan-in init, last layers U(-init_w, init_w)).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_SIG_MAX, LOG_SIG_MIN = 2.0, -20.0


def _fanin_init_(w: torch.Tensor):
    bound = 1.0 / math.sqrt(w.size(0))          # pytorch_util.fanin_init uses size[0] of the weight
    return w.data.uniform_(-bound, bound)


class Mlp(nn.Module):
    def __init__(self, hidden_sizes, output_size, input_size, init_w=3e-3, b_init_value=0.0):
        super().__init__()
        self.fcs = nn.ModuleList()
        in_size = input_size
        for h in hidden_sizes:
            fc = nn.Linear(in_size, h)
            _fanin_init_(fc.weight)
            fc.bias.data.fill_(b_init_value)
            self.fcs.append(fc)
            in_size = h
        self.last_fc = nn.Linear(in_size, output_size)
        self.last_fc.weight.data.uniform_(-init_w, init_w)
        self.last_fc.bias.data.fill_(0)

    def features(self, x):
        for fc in self.fcs:
            x = F.relu(fc(x))
        return x

    def forward(self, x):
        return self.last_fc(self.features(x))


class ConcatMlp(Mlp):
    """Q(s, a): inputs concatenated along dim 1 (mlp.py:121-136)."""

    def forward(self, *inputs):
        return super().forward(torch.cat(inputs, dim=1))


class TanhGaussianPolicy(Mlp):
    def __init__(self, hidden_sizes, obs_dim, action_dim, init_w=1e-3):
        super().__init__(hidden_sizes, output_size=action_dim, input_size=obs_dim, init_w=init_w)
        last = hidden_sizes[-1] if len(hidden_sizes) else obs_dim
        self.last_fc_log_std = nn.Linear(last, action_dim)
        self.last_fc_log_std.weight.data.uniform_(-init_w, init_w)
        self.last_fc_log_std.bias.data.uniform_(-init_w, init_w)

    def forward(self, obs):
        h = self.features(obs)
        mean = self.last_fc(h)
        log_std = torch.clamp(self.last_fc_log_std(h), LOG_SIG_MIN, LOG_SIG_MAX)
        return mean, torch.exp(log_std)

    def rsample_and_logprob(self, obs, return_dist=False):
        """TanhNormal.rsample_and_logprob (ast_sac/torch/core/distributions.py:318-447): a = tanh(z),
        log pi = log N(z) - log(1 - a^2) in the numerically stable form.  ``return_dist`` also returns the
        distribution's (normal_mean, normal_std) for the trainer's diagnostics."""
        mean, std = self(obs)
        z = mean + std * torch.randn_like(mean)
        a = torch.tanh(z)
        log_prob = -0.5 * ((z - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
        log_prob = log_prob - 2.0 * (math.log(2.0) - z - F.softplus(-2.0 * z))
        if return_dist:
            return a, log_prob.sum(dim=1), mean, std
        return a, log_prob.sum(dim=1)

    @torch.no_grad()
    def get_actions(self, obs, deterministic=False):
        mean, std = self(obs.to(self.last_fc.weight.dtype))
        if deterministic:
            return torch.tanh(mean)
        return torch.tanh(mean + std * torch.randn_like(mean))

    def get_action(self, obs_np, deterministic=False):
        """Single-observation interface of the reference policies (policies/base.py:24-37)."""
        dev = self.last_fc.weight.device
        a = self.get_actions(torch.as_tensor(np.asarray(obs_np)[None], dtype=torch.float32, device=dev), deterministic)
        return a[0].cpu().numpy(), {}

    def reset(self):
        pass


class MakeDeterministic:
    def __init__(self, policy):
        self._policy = policy

    def get_actions(self, obs, deterministic=True):
        return self._policy.get_actions(obs, deterministic=True)

    def get_action(self, obs_np):
        return self._policy.get_action(obs_np, deterministic=True)

    def reset(self):
        pass


class SACTrainer:
    def __init__(self, env, policy, qf1, qf2, target_qf1, target_qf2, discount=0.99, reward_scale=1.0,
                 policy_lr=1e-3, qf_lr=1e-3, soft_target_tau=1e-2, target_update_period=1,
                 use_automatic_entropy_tuning=True, target_entropy=None, action_reg_coeff=None, clip_val=np.inf,
                 device=None, capturable=False):
        self.policy, self.qf1, self.qf2 = policy, qf1, qf2
        self.target_qf1, self.target_qf2 = target_qf1, target_qf2
        # (like the reference, the target networks keep their own random initialisation: ast-sac_runner.py:134-145)
        self.soft_target_tau, self.target_update_period = soft_target_tau, target_update_period
        self.use_automatic_entropy_tuning = use_automatic_entropy_tuning
        dev = device or next(policy.parameters()).device
        if use_automatic_entropy_tuning:
            self.target_entropy = (-float(np.prod(env.action_space.shape)) if target_entropy is None else target_entropy)
            self.log_alpha = torch.zeros(1, requires_grad=True, device=dev)
            self.alpha_optimizer = torch.optim.Adam([self.log_alpha], lr=policy_lr, capturable=capturable)
        self.policy_optimizer = torch.optim.Adam(policy.parameters(), lr=policy_lr, capturable=capturable)
        self.qf1_optimizer = torch.optim.Adam(qf1.parameters(), lr=qf_lr, capturable=capturable)
        self.qf2_optimizer = torch.optim.Adam(qf2.parameters(), lr=qf_lr, capturable=capturable)
        self._graph = None
        self.discount, self.reward_scale = discount, reward_scale
        self.action_reg_coeff, self.clip_val = action_reg_coeff, clip_val
        self._n_train_steps_total = 0
        self.eval_statistics = OrderedDict()

    def compute_loss(self, batch):
        rewards, terminals = batch['rewards'], batch['terminals']
        obs, actions, next_obs = batch['observations'], batch['actions'], batch['next_observations']
        new_obs_actions, log_pi, pi_mean, pi_std = self.policy.rsample_and_logprob(obs, return_dist=True)
        log_pi = log_pi.unsqueeze(-1)
        if self.use_automatic_entropy_tuning:
            alpha_loss = -(self.log_alpha * (log_pi + self.target_entropy).detach()).mean()
            alpha = self.log_alpha.exp()
        else:
            alpha_loss, alpha = 0, 1
        q_new_actions = torch.min(self.qf1(obs, new_obs_actions), self.qf2(obs, new_obs_actions))
        policy_loss = (alpha * log_pi - q_new_actions).mean()
        if self.action_reg_coeff:
            policy_loss = policy_loss + self.action_reg_coeff * (new_obs_actions ** 2).mean()
        q1_pred, q2_pred = self.qf1(obs, actions), self.qf2(obs, actions)
        new_next_actions, new_log_pi = self.policy.rsample_and_logprob(next_obs)
        new_log_pi = new_log_pi.unsqueeze(-1)
        target_q_values = torch.min(self.target_qf1(next_obs, new_next_actions),
                                    self.target_qf2(next_obs, new_next_actions)) - alpha * new_log_pi
        q_target = self.reward_scale * rewards + (1.0 - terminals) * self.discount * target_q_values
        q_target = torch.clamp(q_target, min=-self.clip_val, max=self.clip_val)
        qf1_loss = F.mse_loss(q1_pred, q_target.detach())
        qf2_loss = F.mse_loss(q2_pred, q_target.detach())
        # what the reference's eval_statistics are made of (sac.py:262-292); kept as device tensors, read on demand
        self._stat_tensors = (q2_pred, q_target, pi_mean, pi_std)
        return policy_loss, qf1_loss, qf2_loss, alpha_loss, alpha, log_pi, q1_pred

    def train_from_torch(self, batch):
        policy_loss, qf1_loss, qf2_loss, alpha_loss, alpha, log_pi, q1_pred = self.compute_loss(batch)
        if self.use_automatic_entropy_tuning:
            self.alpha_optimizer.zero_grad()
            alpha_loss.backward()
            self.alpha_optimizer.step()
        self.policy_optimizer.zero_grad()
        policy_loss.backward()
        self.policy_optimizer.step()
        self.qf1_optimizer.zero_grad()
        qf1_loss.backward()
        self.qf1_optimizer.step()
        self.qf2_optimizer.zero_grad()
        qf2_loss.backward()
        self.qf2_optimizer.step()
        self._n_train_steps_total += 1
        if self._n_train_steps_total % self.target_update_period == 0:
            with torch.no_grad():
                for src, dst in ((self.qf1, self.target_qf1), (self.qf2, self.target_qf2)):
                    for ps, pd in zip(src.parameters(), dst.parameters()):
                        pd.mul_(1.0 - self.soft_target_tau).add_(ps, alpha=self.soft_target_tau)
        q2_pred, q_target, mean, std = self._stat_tensors
        self._last = tuple(x.detach() if torch.is_tensor(x) else x
                           for x in (policy_loss, qf1_loss, qf2_loss, alpha, log_pi, q1_pred, q2_pred, q_target,
                                     alpha_loss, mean, std))

    train = train_from_torch

    # -- one CUDA graph per update (stock torch.cuda.graphs): sample a batch from the GPU replay buffer + the whole
    # update above.  The five MLPs are tiny (9 -> 256 -> 256 -> 1, batch 256), so an eager update is ~100 kernel
    # launches of a few microseconds each; replaying them as one graph removes the launch overhead.
    def capture(self, replay_buffer, batch_size, warmup=3):
        """Capture one update (batch sampling included) as a CUDA graph.  The `warmup` eager updates torch needs
        before a capture run on live replay data, so the networks, the optimizer states and the update counter are
        snapshotted before them and restored IN PLACE afterwards (the graph holds the tensors' addresses): the graphed
        run performs exactly the updates the caller asks for with train_graphed() -- the same schedule as the eager
        trainer and the reference, none extra."""
        if self.target_update_period != 1:
            raise ValueError("graph capture assumes target_update_period == 1 (the reference default)")
        if not self.policy_optimizer.defaults.get("capturable", False):
            raise ValueError("construct the trainer with capturable=True to capture its optimizers")
        dev = next(self.policy.parameters()).device
        optimizers = [self.policy_optimizer, self.qf1_optimizer, self.qf2_optimizer]
        params = [p for net in self.networks for p in net.parameters()] + [b for net in self.networks for b in net.buffers()]
        if self.use_automatic_entropy_tuning:
            optimizers.append(self.alpha_optimizer)
            params.append(self.log_alpha)
        saved_params = [p.detach().clone() for p in params]
        saved_state = {id(t): t.detach().clone() for opt in optimizers for st in opt.state.values()
                       for t in st.values() if torch.is_tensor(t)}
        steps_before = self._n_train_steps_total
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.train_from_torch(replay_buffer.random_batch(batch_size))
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.train_from_torch(replay_buffer.random_batch(batch_size))
        with torch.no_grad():
            for p, q in zip(params, saved_params):
                p.copy_(q)
            for opt in optimizers:
                for st in opt.state.values():
                    for t in st.values():
                        if torch.is_tensor(t):
                            if id(t) in saved_state:
                                t.copy_(saved_state[id(t)])
                            else:
                                t.zero_()           # created by the warm-up: a fresh Adam state (step 0, zero moments)
        self._graph = graph
        self._n_train_steps_total = steps_before
        return graph

    def train_graphed(self):
        self._graph.replay()
        self._n_train_steps_total += 1

    def get_diagnostics(self):
        if not hasattr(self, "_last"):
            return OrderedDict()
        # the reference's key set and order (sac.py:262-292, torch_rl_algorithm.py:41-44), from the newest update
        from .logging import create_stats_ordered_dict as stats
        policy_loss, qf1_loss, qf2_loss, alpha, log_pi, q1_pred, q2_pred, q_target, alpha_loss, mean, std = self._last
        d = OrderedDict([('num train calls', self._n_train_steps_total),
                         ('QF1 Loss', float(qf1_loss)), ('QF2 Loss', float(qf2_loss)), ('Policy Loss', float(policy_loss))])
        d.update(stats('Q1 Predictions', q1_pred))
        d.update(stats('Q2 Predictions', q2_pred))
        d.update(stats('Q Targets', q_target))
        d.update(stats('Log Pis', log_pi))
        d.update(stats('policy/mean', torch.tanh(mean)))      # TanhNormal.mean (distributions.py:425-427)
        d.update(stats('policy/normal/std', std))
        d.update(stats('policy/normal/log_std', torch.log(std)))
        if self.use_automatic_entropy_tuning:
            d['Alpha'] = float(alpha)
            d['Alpha Loss'] = float(alpha_loss)
        return d

    def get_snapshot(self):
        """sac.py:318-325."""
        return dict(policy=self.policy, qf1=self.qf1, qf2=self.qf2, target_qf1=self.target_qf1, target_qf2=self.target_qf2)

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.target_qf1, self.target_qf2]

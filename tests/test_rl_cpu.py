"""Host-side logic of the batched RL consumers, on CPU tensors with a
duck-typed fake env: ring semantics of the replay buffer against a NumPy restatement of the reference's
SimpleReplayBuffer, the sampler's per-environment semantics against a sequential ast_sac_rollout loop, the
action normalisation arithmetic, and one SAC update."""
import math

import numpy as np
import pytest
import torch

from ast_sac_b200.rl import (BatchRLAlgorithm, ConcatMlp, GpuReplayBuffer, MakeDeterministic, NormalizedBoxEnv,
                             SACTrainer, TanhGaussianPolicy, VectorizedPathCollector, batched_ast_sac_rollout)
from ast_sac_b200.spaces import Box


class NumpyRing:
    """simple_replay_buffer.py:8-84 restated: add_sample / _advance / random_batch index range."""

    def __init__(self, n, od, ad):
        self.n, self.top, self.size = n, 0, 0
        self.o, self.no = np.zeros((n, od)), np.zeros((n, od))
        self.a, self.r, self.t = np.zeros((n, ad)), np.zeros((n, 1)), np.zeros((n, 1), dtype='uint8')

    def add_sample(self, o, a, r, no, t):
        self.o[self.top], self.a[self.top], self.r[self.top], self.t[self.top], self.no[self.top] = o, a, r, t, no
        self.top = (self.top + 1) % self.n
        self.size = min(self.size + 1, self.n)


def test_replay_ring_matches_reference_semantics_with_masks_and_wraparound():
    rng = np.random.default_rng(0)
    N, od, ad = 37, 8, 1
    buf = GpuReplayBuffer(N, observation_dim=od, action_dim=ad, device="cpu", dtype=torch.float64)
    ref = NumpyRing(N, od, ad)
    for it in range(40):
        B = int(rng.integers(1, 20))
        o, no = rng.normal(size=(B, od)), rng.normal(size=(B, od))
        a, r = rng.normal(size=(B, ad)), rng.normal(size=(B, 1))
        t = rng.integers(0, 2, size=(B, 1)).astype(np.uint8)
        mask = rng.random(B) < 0.7 if it % 3 else np.ones(B, bool)
        buf.add_batch(torch.tensor(o), torch.tensor(a), torch.tensor(r), torch.tensor(no), torch.tensor(t),
                      mask=None if it % 3 == 0 else torch.tensor(mask))
        for b in range(B):
            if mask[b]:
                ref.add_sample(o[b], a[b], r[b], no[b], t[b])
        assert buf._top == ref.top and buf._size == ref.size
    assert np.array_equal(buf._observations[:N].numpy(), ref.o)
    assert np.array_equal(buf._next_obs[:N].numpy(), ref.no)
    assert np.array_equal(buf._actions[:N].numpy(), ref.a)
    assert np.array_equal(buf._rewards[:N].numpy(), ref.r)
    assert np.array_equal(buf._terminals[:N].numpy(), ref.t)
    assert buf.num_steps_can_sample() == N and buf.get_diagnostics()['size'] == N


def test_replay_random_batch_and_path_api():
    buf = GpuReplayBuffer(100, observation_dim=8, action_dim=1, device="cpu", seed=3)
    path = dict(observations=np.arange(40, dtype=np.float32).reshape(5, 8), actions=np.ones((5, 1), np.float32) * 0.5,
                rewards=np.arange(5, dtype=np.float64).reshape(5, 1), next_observations=np.ones((5, 8), np.float32),
                terminals=np.array([[False]] * 4 + [[True]]), agent_infos=[{}] * 5, env_infos=[{}] * 5)
    buf.add_paths([path])
    buf.add_sample(np.zeros(8), np.array([0.25]), 7.0, np.ones(8), True, env_info={})
    assert buf.num_steps_can_sample() == 6
    batch = buf.random_batch(256)
    assert set(batch) == {"observations", "actions", "rewards", "terminals", "next_observations"}
    assert batch["observations"].shape == (256, 8) and batch["terminals"].dtype == torch.float32
    # only the 6 stored rows can be drawn, every one of them is (uniform with replacement)
    assert set(batch["rewards"].reshape(-1).tolist()) == {0.0, 1.0, 2.0, 3.0, 4.0, 7.0}
    assert set(batch["terminals"].reshape(-1).tolist()) == {0.0, 1.0}
    empty = GpuReplayBuffer(10, observation_dim=8, action_dim=1, device="cpu")
    assert empty.random_batch(4)["observations"].shape == (4, 8)       # no crash on an empty buffer
    with pytest.raises(ValueError):
        empty.add_batch(torch.zeros(11, 8), torch.zeros(11, 1), torch.zeros(11, 1), torch.zeros(11, 8), torch.zeros(11, 1))


class FakeBatchedEnv:
    """Duck-typed batched env on CPU tensors: returns VIEWS of its output buffers like the real one, leaves
    finished environments untouched, episode b ends after 2 + b % 5 steps (terminal if b is odd)."""

    def __init__(self, B):
        self.num_envs = B
        self.observation_space = Box(low=-np.ones(8, np.float32) * 10, high=np.ones(8, np.float32) * 10, dtype=np.float32)
        self.action_space = Box(low=np.array([-np.pi / 6], np.float32), high=np.array([np.pi / 6], np.float32), dtype=np.float32)
        self.obs_buf = torch.zeros(B, 8)
        self.reward_buf = torch.zeros(B, dtype=torch.float64)
        self.t = torch.zeros(B, dtype=torch.int64)
        self.done = torch.zeros(B, dtype=torch.bool)
        self.last_actions = None

    def reset(self, mask=None):
        self.obs_buf.copy_(torch.arange(self.num_envs, dtype=torch.float32)[:, None].expand(-1, 8) * 0.01)
        self.t.zero_()
        self.done.zero_()
        return self.obs_buf

    def step(self, a):
        a = a.reshape(-1).to(torch.float64)
        self.last_actions = a.clone()
        live = ~self.done
        self.obs_buf[live] += a[live].to(torch.float32)[:, None]
        self.reward_buf[live] = -a[live].abs() + 1.0
        self.t[live] += 1
        length = 2 + torch.arange(self.num_envs) % 5
        newly = live & (self.t >= length)
        self.done |= newly
        terminal = newly & (torch.arange(self.num_envs) % 2 == 1)
        info = dict(events=newly.to(torch.int32) * 32, terminal=terminal, test_ship_stop=newly, obs_ship_stop=newly,
                    substeps=live.to(torch.int32))
        return self.obs_buf, self.reward_buf, newly | (self.done & ~live & False), info


class LinearPolicy:
    def get_actions(self, obs, deterministic=False):
        return torch.tanh(obs[:, :1] * 3.0 - 0.2)

    def reset(self):
        pass


def test_normalized_env_matches_reference_arithmetic():
    env = NormalizedBoxEnv(FakeBatchedEnv(4), reward_scale=0.75)
    assert np.array_equal(env.action_space.low, [-1.0]) and np.array_equal(env.action_space.high, [1.0])
    a = torch.tensor([[-1.0], [0.3], [1.0], [1.7]])
    lb, ub = np.float32(-np.pi / 6), np.float32(np.pi / 6)
    want = np.clip(lb + (a.numpy() + np.float32(1.0)) * np.float32(0.5) * (ub - lb), lb, ub)   # normalized_box_env.py:48-51
    got = env.scale_action(a)
    assert got.dtype == torch.float32 and np.array_equal(got.numpy(), want)
    assert np.array_equal(env.scale_action(a.numpy()), want)
    env.reset()
    _, r, _, _ = env.step(a)
    assert torch.allclose(r, (1.0 - torch.tensor(want, dtype=torch.float64).abs().reshape(-1)) * 0.75)
    assert env.num_envs == 4                               # ProxyEnv attribute pass-through


def test_batched_rollout_equals_sequential_rollouts():
    B, T = 12, 9
    env = NormalizedBoxEnv(FakeBatchedEnv(B), reward_scale=0.75)
    buf = GpuReplayBuffer(1000, env=env, device="cpu")
    rb = batched_ast_sac_rollout(env, LinearPolicy(), T, replay_buffer=buf)
    lengths = rb.path_lengths().numpy()
    assert np.array_equal(lengths, 2 + np.arange(B) % 5)
    assert buf.num_steps_can_sample() == lengths.sum() == rb.num_steps
    paths = rb.paths()
    assert len(paths) == B
    for b, path in enumerate(paths):
        # the same episode, one environment at a time, with the reference's loop (rollout_functions.py:109-150)
        single = NormalizedBoxEnv(FakeBatchedEnv(B), reward_scale=0.75)
        o = single.reset().clone()
        n = 0
        while n < T:
            a = LinearPolicy().get_actions(o)
            no, r, done, info = single.step(a)
            assert np.allclose(path["observations"][n], o[b].numpy()) and np.allclose(path["actions"][n], a[b].numpy())
            assert np.allclose(path["next_observations"][n], no[b].numpy()) and np.isclose(path["rewards"][n, 0], float(r[b]))
            assert bool(path["terminals"][n, 0]) == bool(info["terminal"][b]) and bool(path["dones"][n, 0]) == bool(done[b])
            n += 1
            if bool(done[b]):
                break
            o = no.clone()
        assert n == len(path["actions"]) == lengths[b]
        assert path["rewards"].shape == (n, 1) and path["terminals"].shape == (n, 1)
    # what went into the replay buffer is exactly the valid transitions (step-major order)
    stored = buf._rewards[:rb.num_steps, 0].numpy()
    want = rb.rewards.squeeze(-1)[rb.valid].numpy()
    assert np.allclose(stored, want)


def test_collector_counts_and_limits_active_envs():
    env = NormalizedBoxEnv(FakeBatchedEnv(16))
    buf = GpuReplayBuffer(1000, env=env, device="cpu")
    col = VectorizedPathCollector(env, LinearPolicy(), replay_buffer=buf)
    batches = col.collect_new_steps(max_path_length=9, num_steps=20)
    got = sum(b.num_steps for b in batches)
    assert 20 <= got <= 20 + 16 * 9 and buf.num_steps_can_sample() == got
    assert int(batches[0].valid[0].sum()) == 3              # ceil(20 / 9) environments activated in the first wave
    d = col.get_diagnostics()
    assert d['num steps total'] == got and d['path length Max'] <= 6
    paths = col.collect_new_paths(9, 5)
    assert all(set(p) >= {"observations", "actions", "rewards", "next_observations", "terminals", "dones"} for p in paths)


def test_tanh_gaussian_logprob_and_one_sac_update():
    torch.manual_seed(0)
    pol = TanhGaussianPolicy([32, 32], obs_dim=8, action_dim=1)
    obs = torch.randn(64, 8)
    a, lp = pol.rsample_and_logprob(obs)
    mean, std = pol(obs)
    base = torch.distributions.Normal(mean, std)
    z = torch.atanh(a.clamp(-1 + 1e-6, 1 - 1e-6))
    want = (base.log_prob(z) - torch.log(1 - a ** 2 + 1e-12)).sum(1)
    assert torch.allclose(lp, want, atol=1e-3)
    env = FakeBatchedEnv(4)
    qf1, qf2, tq1, tq2 = (ConcatMlp([32, 32], 1, 9) for _ in range(4))
    tr = SACTrainer(env, pol, qf1, qf2, tq1, tq2, discount=0.965, reward_scale=0.75, policy_lr=8e-5, qf_lr=8e-5,
                    soft_target_tau=1e-3, action_reg_coeff=0.01, clip_val=100)
    assert tr.target_entropy == -1.0
    before = [p.clone() for p in tq1.parameters()]
    q_before = [p.clone() for p in qf1.parameters()]
    batch = dict(observations=obs, actions=torch.rand(64, 1) * 2 - 1, rewards=torch.randn(64, 1),
                 terminals=(torch.rand(64, 1) < 0.2).float(), next_observations=torch.randn(64, 8))
    tr.train_from_torch(batch)
    d = tr.get_diagnostics()
    assert all(math.isfinite(v) for v in d.values())
    # soft update: target <- (1 - tau) target + tau q  (pytorch_util.soft_update_from_to)
    for b, q, t in zip(before, qf1.parameters(), tq1.parameters()):
        assert torch.allclose(t, b * (1 - 1e-3) + q.detach() * 1e-3, atol=1e-7)
    assert any(not torch.equal(a_, b_) for a_, b_ in zip(q_before, qf1.parameters()))
    assert torch.equal(MakeDeterministic(pol).get_actions(obs), torch.tanh(pol(obs)[0]))


def test_batch_rl_algorithm_runs_on_fake_env():
    torch.manual_seed(1)
    env = NormalizedBoxEnv(FakeBatchedEnv(32), reward_scale=0.75)
    pol = TanhGaussianPolicy([16, 16], obs_dim=8, action_dim=1)
    qs = [ConcatMlp([16, 16], 1, 9) for _ in range(4)]
    buf = GpuReplayBuffer(5000, env=env, device="cpu", seed=0)
    tr = SACTrainer(env, pol, *qs, discount=0.965, reward_scale=0.75)
    logs = []
    alg = BatchRLAlgorithm(tr, VectorizedPathCollector(env, pol, replay_buffer=buf),
                           VectorizedPathCollector(env, MakeDeterministic(pol)), buf, batch_size=32, max_path_length=9,
                           num_epochs=2, num_eval_steps_per_epoch=40, num_expl_steps_per_train_loop=60,
                           num_trains_per_train_loop=3, min_num_steps_before_training=100, log=logs.append)
    hist = alg.train()
    assert len(hist) == 2 and logs[1]['epoch'] == 1
    assert hist[1]['replay_buffer/size'] >= 100 + 2 * 60
    assert tr._n_train_steps_total == 6 and 'trainer/QF1 Loss' in hist[0]


def test_run_directory_has_the_reference_loggers_format(tmp_path):
    """progress.csv / variant.json / debug.log / params.pkl of a training run.  The fixture
    tests/golden/progress_format.json was written by tests/golden/make_progress_golden.py from a run of the UNMODIFIED
    reference's TorchBatchRLAlgorithm + logger (two miniature epochs, snapshot mode 'last'): same file list, the same
    86 csv columns in the same (sorted) order, one row per epoch, the same snapshot keys."""
    import csv
    import json
    import os

    from ast_sac_b200.rl.logging import Logger, setup_logger
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "progress_format.json")))
    torch.manual_seed(3)
    env = NormalizedBoxEnv(FakeBatchedEnv(8), reward_scale=0.75)
    pol = TanhGaussianPolicy([16, 16], obs_dim=8, action_dim=1)
    qs = [ConcatMlp([16, 16], 1, 9) for _ in range(4)]
    buf = GpuReplayBuffer(1000, env=env, device="cpu", seed=0)
    tr = SACTrainer(env, pol, *qs, discount=0.965, reward_scale=0.75)
    lg = Logger()
    lg.set_print(False)
    variant = dict(ref["variant"])
    run_dir = setup_logger("ast-sac_maritime_logs", variant=variant, base_log_dir=str(tmp_path), target=lg, seed=5)
    # <base>/<prefix with '-'>/<prefix>_<timestamp>_<id>--s-<seed> (launcher_utils.py:181-225)
    assert os.path.basename(os.path.dirname(run_dir)) == "ast-sac-maritime-logs"
    assert os.path.basename(run_dir).startswith("ast-sac_maritime_logs_") and run_dir.endswith("_0000--s-5")
    # save_env_in_snapshot: the fake env is a plain picklable object like the real one
    alg = BatchRLAlgorithm(tr, VectorizedPathCollector(env, pol, replay_buffer=buf),
                           VectorizedPathCollector(env, MakeDeterministic(pol)), buf, batch_size=8, max_path_length=9,
                           num_epochs=2, num_eval_steps_per_epoch=9, num_expl_steps_per_train_loop=9,
                           num_trains_per_train_loop=2, min_num_steps_before_training=9, log=lambda s: None, logger=lg)
    alg.train()
    lg.close()
    assert sorted(os.listdir(run_dir)) == ref["files"]
    rows = list(csv.reader(open(os.path.join(run_dir, "progress.csv"))))
    assert rows[0] == ref["progress_columns"]
    assert len(rows) - 1 == ref["progress_rows"] == 2
    assert all(len(r) == len(rows[0]) and all(c != "" for c in r) for r in rows[1:])
    col = {k: i for i, k in enumerate(rows[0])}
    assert [r[col["Epoch"]] for r in rows[1:]] == ["0", "1"] and [r[col["epoch"]] for r in rows[1:]] == ["0", "1"]
    assert int(rows[2][col["trainer/num train calls"]]) == 4
    assert float(rows[2][col["expl/num steps total"]]) >= 27 and float(rows[1][col["eval/Num Paths"]]) >= 1
    snap = torch.load(os.path.join(run_dir, "params.pkl"), weights_only=False)
    assert sorted(snap.keys()) == ref["snapshot_keys"]
    assert {k: type(v).__name__ for k, v in snap.items()} == ref["snapshot_types"]
    assert json.load(open(os.path.join(run_dir, "variant.json"))) == ref["variant"]
    # later dumps with other keys keep the first dump's columns (logging.py:287-300)
    lg2 = Logger()
    lg2.set_print(False)
    lg2.add_tabular_output(str(tmp_path / "t.csv"))
    lg2.record_dict(dict(b=1, a=2)); lg2.dump_tabular()
    lg2.record_dict(dict(b=3, c=4)); lg2.dump_tabular()
    lg2.close()
    assert open(tmp_path / "t.csv").read().splitlines() == ["a,b", "2,1", ",3"]
    # snapshot modes (logging.py:314-336)
    lg3 = Logger()
    lg3.set_snapshot_dir(str(tmp_path)); lg3.set_snapshot_mode("gap_and_last"); lg3.set_snapshot_gap(2)
    for itr in range(3):
        lg3.save_itr_params(itr, dict(x=itr))
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".pkl")) == ["itr_0.pkl", "itr_2.pkl", "params.pkl"]
    assert torch.load(tmp_path / "params.pkl", weights_only=False) == dict(x=2)

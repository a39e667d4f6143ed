"""GPU tests of the batched RL consumers on the real environment."""
import json

import numpy as np
import pytest
import torch

from ast_sac_b200 import _lib as L
from ast_sac_b200 import scenarios as S
from ast_sac_b200.env import events_to_string
from ast_sac_b200.rl import (BatchRLAlgorithm, ConcatMlp, GpuReplayBuffer, MakeDeterministic, NormalizedBoxEnv,
                             SACTrainer, TanhGaussianPolicy, VectorizedPathCollector, batched_ast_sac_rollout)

from helpers import golden, golden_names
from product_helpers import env_from_meta

pytestmark = pytest.mark.gpu


class LinearTanhPolicy:
    """tests/golden/make_golden.py:LinearTanhPolicy on device tensors (float32)."""

    def __init__(self, w):
        self.w = [np.float32(x) for x in w]

    def reset(self):
        pass

    def get_actions(self, obs, deterministic=False):
        o = obs.to(torch.float32)
        z = float(self.w[0]) * o[:, 3] + float(self.w[1]) * o[:, 4] + float(self.w[2]) * o[:, 5] + float(self.w[3])
        return torch.tanh(z).reshape(-1, 1)


@pytest.mark.parametrize("name", golden_names("sampler_"))
def test_batched_rollout_matches_reference_sampler_golden(name):
    """The reference's ast_sac_rollout + NormalizedBoxEnv + MultiShipRLEnv (unmodified, one env) against the
    batched sampler on the CUDA env.  Tolerances: the reference evaluates tan() of the float32 scoping angle
    in float32, the product upcasts to FP64 first (quirk 12) -> waypoints differ by <= 1e-7 relative."""
    g = golden(name)
    meta = json.loads(str(g["meta"]))
    B = 4
    env, _ = env_from_meta(meta, num_envs=B)
    wrapped = NormalizedBoxEnv(env, reward_scale=meta["reward_scale"])
    buf = GpuReplayBuffer(1000, env=wrapped)
    rb = batched_ast_sac_rollout(wrapped, LinearTanhPolicy(meta["w"]), meta["max_path_length"], replay_buffer=buf)
    paths = rb.paths(events_to_string)
    assert len(paths) == B
    n = len(g["actions"])
    for p in paths:
        assert len(p["actions"]) == n
        np.testing.assert_allclose(p["observations"], g["observations"], rtol=1e-5, atol=0.05)
        np.testing.assert_allclose(p["next_observations"], g["next_observations"], rtol=1e-5, atol=0.05)
        np.testing.assert_allclose(p["actions"], g["actions"], atol=1e-5)
        np.testing.assert_allclose(p["rewards"], g["rewards"], rtol=1e-4, atol=1e-4)
        assert np.array_equal(p["terminals"].astype(np.uint8), g["terminals"])
        assert np.array_equal(p["dones"].astype(np.uint8), g["dones"])
        assert [i["events"] for i in p["env_infos"]] == [events_to_string(int(e)) for e in g["events"]]
    assert buf.num_steps_can_sample() == B * n
    env.close()


def test_collector_fills_gpu_replay_buffer_and_sac_trains(tmp_path):
    """config 5 in miniature: GPU-resident rollouts -> GPU replay buffer -> SAC updates, nothing on the host; the run
    directory (progress.csv, params.pkl with the pickled env, variant.json, debug.log) has the format of the
    reference's logger (tests/golden/progress_format.json, from a run of the unmodified reference)."""
    import csv
    import json
    import os
    from ast_sac_b200.rl.logging import Logger, setup_logger
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "progress_format.json")))
    lg = Logger()
    lg.set_print(False)
    run_dir = setup_logger("ast-sac_maritime_logs", variant=ref["variant"], base_log_dir=str(tmp_path), target=lg)
    torch.manual_seed(0)
    B = 2048
    args = S.get_env_args(time_step=4, collav_mode="none")
    env, _ = S.prepare_multiship_rl_env(args, num_envs=B)
    wrapped = NormalizedBoxEnv(env, reward_scale=0.75)
    dev = env.obs_buf.device
    pol = TanhGaussianPolicy([256, 256], obs_dim=8, action_dim=1).to(dev)
    qs = [ConcatMlp([256, 256], 1, 9).to(dev) for _ in range(4)]
    buf = GpuReplayBuffer(300000, env=wrapped, seed=0)
    tr = SACTrainer(wrapped, pol, *qs, discount=0.965, reward_scale=0.75, policy_lr=8e-5, qf_lr=8e-5,
                    soft_target_tau=1e-3, action_reg_coeff=0.01, clip_val=100)
    expl = VectorizedPathCollector(wrapped, pol, replay_buffer=buf)
    evalc = VectorizedPathCollector(wrapped, MakeDeterministic(pol))
    logs = []
    alg = BatchRLAlgorithm(tr, expl, evalc, buf, batch_size=256, max_path_length=9, num_epochs=2,
                           num_eval_steps_per_epoch=180, num_expl_steps_per_train_loop=4096,
                           num_trains_per_train_loop=20, min_num_steps_before_training=8192, log=logs.append, logger=lg)
    hist = alg.train()
    lg.close()
    rows = list(csv.reader(open(os.path.join(run_dir, "progress.csv"))))
    assert rows[0] == ref["progress_columns"] and len(rows) == 3 and sorted(os.listdir(run_dir)) == ref["files"]
    snap = torch.load(os.path.join(run_dir, "params.pkl"), weights_only=False)      # rebuilds the pickled env on the GPU
    assert sorted(snap.keys()) == ref["snapshot_keys"]
    assert {k: type(v).__name__ for k, v in snap.items()} == ref["snapshot_types"]
    assert snap["exploration/env"].wrapped_env.num_envs == B
    snap["exploration/env"].wrapped_env.close()
    if snap["evaluation/env"] is not snap["exploration/env"]:
        snap["evaluation/env"].wrapped_env.close()
    assert hist[-1]['replay_buffer/size'] >= 8192 + 2 * 4096
    assert buf._observations.device.type == "cuda" and buf.random_batch(256)["rewards"].device.type == "cuda"
    assert all(np.isfinite(v) for k, v in hist[-1].items() if k.startswith("trainer/"))
    # every stored transition is a real one: observations inside the map, |action| <= 1
    n = buf.num_steps_can_sample()
    assert bool((buf._actions[:n].abs() <= 1).all()) and bool(torch.isfinite(buf._rewards[:n]).all())
    assert 1.0 <= hist[-1]['exploration/path length Mean'] <= 9.0
    env.close()


def test_sac_update_as_cuda_graph_matches_eager_update():
    """One SAC update replayed as a CUDA graph (sample from the GPU replay buffer + the four optimiser steps + the
    soft target update) changes the networks exactly like the eager update on the same batch."""
    import copy
    dev = torch.device("cuda:0")
    torch.manual_seed(3)

    class E:
        action_space = type("B", (), {"shape": (1,)})()

    def make(capturable):
        torch.manual_seed(7)
        pol = TanhGaussianPolicy([64, 64], obs_dim=8, action_dim=1).to(dev)
        qs = [ConcatMlp([64, 64], 1, 9).to(dev) for _ in range(4)]
        return SACTrainer(E(), pol, *qs, discount=0.965, reward_scale=0.75, policy_lr=8e-5, qf_lr=8e-5,
                          soft_target_tau=1e-3, action_reg_coeff=0.01, clip_val=100, capturable=capturable)

    buf = GpuReplayBuffer(4096, observation_dim=8, action_dim=1, device=dev)
    n = 4096
    buf.add_batch(torch.randn(n, 8, device=dev), torch.rand(n, 1, device=dev) * 2 - 1, torch.randn(n, 1, device=dev),
                  torch.randn(n, 8, device=dev), (torch.rand(n, 1, device=dev) < 0.1))
    eager, graphed = make(False), make(True)
    before = [p.detach().clone() for p in graphed.qf1.parameters()]
    graphed.capture(buf, 256, warmup=3)
    # the capture's warm-up updates are undone: same parameters, fresh optimizer state, no update counted
    assert graphed._n_train_steps_total == 0
    assert all(torch.equal(a, b) for a, b in zip(before, graphed.qf1.parameters()))
    assert all(float(st["step"]) == 0 and float(st["exp_avg"].abs().sum()) == 0 for st in graphed.qf1_optimizer.state.values())
    # both see different random batches / noise, so compare statistics, not bits
    for _ in range(20):
        graphed.train_graphed()
        eager.train_from_torch(buf.random_batch(256))
    torch.cuda.synchronize()
    assert graphed._n_train_steps_total == eager._n_train_steps_total == 20
    dg, de = graphed.get_diagnostics(), eager.get_diagnostics()
    assert all(np.isfinite(v) for v in dg.values())
    assert abs(dg['Alpha'] - de['Alpha']) < 1e-3                      # 20 Adam steps of lr 8e-5 on log_alpha
    # the replayed graph really trains: parameters moved away from a fresh copy, targets follow slowly
    fresh = make(True)
    moved = sum(float((a - b).abs().sum()) for a, b in zip(graphed.qf1.parameters(), fresh.qf1.parameters()))
    assert moved > 0
    # soft update inside the graph: after 23 updates with tau = 1e-3 the target has covered ~2.3 % of its initial
    # distance to the (slowly moving) Q network
    d0 = sum(float((a - b).abs().sum()) for a, b in zip(fresh.target_qf1.parameters(), fresh.qf1.parameters()))
    d1 = sum(float((a - b).abs().sum()) for a, b in zip(graphed.target_qf1.parameters(), fresh.target_qf1.parameters()))
    assert 0.01 * d0 < d1 < 0.05 * d0, (d0, d1)

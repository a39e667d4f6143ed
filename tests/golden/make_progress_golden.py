"""On-disk formats of the reference's training run: column set of progress.csv, key set of
params.pkl / variant.json, made by running the UNMODIFIED reference's own TorchBatchRLAlgorithm + rllab logger for two
miniature epochs (ast_sac/core/logging.py:274-336, rl_algorithm.py:76-141, launchers/launcher_utils.py:227-300).

    python tests/golden/make_progress_golden.py        # needs /root/reference; writes tests/golden/progress_format.json

The stub gtimer below implements the four calls the reference makes (timed_for, stamp, get_times, reset_root) with
gtimer's semantics for them: stamps are per-iteration lists, a non-unique stamp accumulates within an iteration.
"""
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402


def install_gtimer():
    gt = types.ModuleType("gtimer")
    state = dict(itrs={}, last=time.perf_counter(), start=time.perf_counter(), itr=-1)

    def stamp(name, unique=True, **k):
        now = time.perf_counter()
        lst = state["itrs"].setdefault(name, [])
        while len(lst) <= state["itr"]:
            lst.append(0.0)
        lst[state["itr"]] += now - state["last"]
        state["last"] = now

    def timed_for(iterable, save_itrs=True, **k):
        for x in iterable:
            state["itr"] += 1
            yield x

    def get_times():
        return types.SimpleNamespace(stamps=types.SimpleNamespace(itrs=state["itrs"]),
                                     total=time.perf_counter() - state["start"])

    gt.stamp, gt.timed_for, gt.get_times = stamp, timed_for, get_times
    gt.reset_root = gt.blank_stamp = gt.reset = lambda *a, **k: None
    sys.modules["gtimer"] = gt


def main():
    assert H.reference_available(), "needs /root/reference"
    H.install_stubs()
    install_gtimer()
    spaces = sys.modules["gymnasium.spaces"]       # space classes env_replay_buffer.py / env_utils.py only test against
    for name in ("Discrete", "Tuple", "Dict", "MultiDiscrete", "MultiBinary"):
        if not hasattr(spaces, name):
            setattr(spaces, name, type(name, (), {}))
    import numpy as np
    import torch
    import ast_sac.torch.utils.pytorch_util as ptu
    from ast_sac.core.logging import logger
    from ast_sac.data_management.env_replay_buffer import EnvReplayBuffer
    from ast_sac.env_wrapper.normalized_box_env import NormalizedBoxEnv
    from ast_sac.samplers.data_collector.path_collector import MdpPathCollector
    from ast_sac.samplers.data_collector.rollout_functions import ast_sac_rollout
    from ast_sac.torch.core.torch_rl_algorithm import TorchBatchRLAlgorithm
    from ast_sac.torch.networks.mlp import ConcatMlp
    from ast_sac.torch.sac.policies.gaussian_policy import MakeDeterministic, TanhGaussianPolicy
    from ast_sac.torch.sac.sac import SACTrainer

    torch.manual_seed(0)
    np.random.seed(0)
    ptu.set_gpu_mode(False)
    variant = dict(   # run/ast-sac_runner.py:211-236 with miniature sizes
        algorithm="SAC", version="normal", layer_size=16, replay_buffer_size=1000,
        algorithm_kwargs=dict(num_epochs=2, num_eval_steps_per_epoch=9, num_trains_per_train_loop=2,
                              num_expl_steps_per_train_loop=9, min_num_steps_before_training=9, max_path_length=9,
                              batch_size=8),
        trainer_kwargs=dict(discount=0.965, soft_target_tau=1e-3, target_update_period=1, policy_lr=8e-5, qf_lr=8e-5,
                            reward_scale=0.75, use_automatic_entropy_tuning=True, action_reg_coeff=0.0, clip_val=np.inf))
    log_dir = tempfile.mkdtemp(prefix="ref_progress_")
    # what setup_logger does (launcher_utils.py:270-295), without its git / conf lookups
    logger.log_variant(os.path.join(log_dir, "variant.json"), variant)
    logger.add_text_output(os.path.join(log_dir, "debug.log"))
    logger.add_tabular_output(os.path.join(log_dir, "progress.csv"))
    logger.set_snapshot_dir(log_dir)
    logger.set_snapshot_mode("last")
    logger.set_snapshot_gap(1)
    logger.set_log_tabular_only(False)

    env, _ = H.make_rl_env(H.Args(time_step=4, collav_mode="none"))
    expl_env = NormalizedBoxEnv(env, reward_scale=0.75)
    eval_env = NormalizedBoxEnv(env, reward_scale=0.75)
    obs_dim, act_dim = expl_env.observation_space.low.size, expl_env.action_space.low.size
    M = variant["layer_size"]
    qf1, qf2, tq1, tq2 = (ConcatMlp(input_size=obs_dim + act_dim, output_size=1, hidden_sizes=[M, M]) for _ in range(4))
    policy = TanhGaussianPolicy(obs_dim=obs_dim, action_dim=act_dim, hidden_sizes=[M, M])
    trainer = SACTrainer(env=eval_env, policy=policy, qf1=qf1, qf2=qf2, target_qf1=tq1, target_qf2=tq2,
                         **variant["trainer_kwargs"])
    alg = TorchBatchRLAlgorithm(
        trainer=trainer, exploration_env=expl_env, evaluation_env=eval_env,
        exploration_data_collector=MdpPathCollector(expl_env, policy, rollout_fn=ast_sac_rollout),
        evaluation_data_collector=MdpPathCollector(eval_env, MakeDeterministic(policy), rollout_fn=ast_sac_rollout),
        replay_buffer=EnvReplayBuffer(variant["replay_buffer_size"], expl_env), **variant["algorithm_kwargs"])
    alg.to(ptu.device)
    alg.train()

    import csv
    with open(os.path.join(log_dir, "progress.csv")) as f:
        rows = list(csv.reader(f))
    snap = torch.load(os.path.join(log_dir, "params.pkl"), weights_only=False)
    with open(os.path.join(log_dir, "variant.json")) as f:
        var = json.load(f)
    out = dict(progress_columns=rows[0], progress_rows=len(rows) - 1,
               snapshot_keys=sorted(snap.keys()),
               snapshot_types={k: type(v).__name__ for k, v in snap.items()},
               variant=var, files=sorted(os.listdir(log_dir)))
    with open(os.path.join(HERE, "progress_format.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(rows[0]), "columns;", out["snapshot_keys"], out["files"])


if __name__ == "__main__":
    main()

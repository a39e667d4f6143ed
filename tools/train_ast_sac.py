#!/usr/bin/env python
"""BASELINE config 5: the AST-SAC training loop of run/ast-sac_runner.py with GPU-resident rollouts and a
GPU-resident replay buffer.  Flags and defaults follow run/ast-sac_runner.py:22-105; ``--num_envs`` is new
(the reference steps one environment).  The SAC trainer is stock PyTorch (ast_sac_b200/rl/sac.py).

    python tools/train_ast_sac.py --num_envs 4096 --num_epochs 5 --collav_mode sbmpc
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_sac_b200 import scenarios as S  # noqa: E402
from ast_sac_b200.rl import (BatchRLAlgorithm, ConcatMlp, GpuReplayBuffer, MakeDeterministic, NormalizedBoxEnv,  # noqa: E402
                             SACTrainer, TanhGaussianPolicy, VectorizedPathCollector)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--num_envs', type=int, default=4096)
    p.add_argument('--max_sampling_frequency', type=int, default=9)
    p.add_argument('--time_step', type=int, default=4)
    p.add_argument('--radius_of_acceptance', type=int, default=300)
    p.add_argument('--lookahead_distance', type=int, default=1000)
    p.add_argument('--collav_mode', type=str, default='sbmpc')
    p.add_argument('--layer_size', type=int, default=256)
    p.add_argument('--replay_buffer_size', type=int, default=300000)
    p.add_argument('--batch_size', type=int, default=256)
    p.add_argument('--num_epochs', type=int, default=500)
    p.add_argument('--num_eval_steps_per_epoch', type=int, default=180)
    p.add_argument('--num_trains_per_train_loop', type=int, default=240)
    p.add_argument('--num_expl_steps_per_train_loop', type=int, default=256)
    p.add_argument('--min_num_steps_before_training', type=int, default=8192)
    p.add_argument('--max_path_length', type=int, default=9)
    p.add_argument('--discount', type=float, default=0.965)
    p.add_argument('--soft_target_tau', type=float, default=1e-3)
    p.add_argument('--target_update_period', type=int, default=1)
    p.add_argument('--policy_lr', type=float, default=8e-5)
    p.add_argument('--qf_lr', type=float, default=8e-5)
    p.add_argument('--reward_scale', type=float, default=0.75)
    p.add_argument('--action_reg_coeff', type=float, default=0.01)
    p.add_argument('--clip_val', type=float, default=100)
    p.add_argument('--seed', type=int, default=0)
    p.add_argument('--cuda_graph', type=int, default=1, help="replay each SAC update as one CUDA graph (0: eager)")
    p.add_argument('--do_logging', type=int, default=0,
                   help="write the run directory of the reference's logger: progress.csv, variant.json, debug.log, "
                        "params.pkl (ast-sac_runner.py:240-241; ast_sac_b200/rl/logging.py)")
    p.add_argument('--base_log_dir', type=str, default=None, help="default: ./run/logs like the reference")
    p.add_argument('--snapshot_mode', type=str, default='last')
    return p.parse_args()


def main():
    a = parse()
    torch.manual_seed(a.seed)
    args = S.get_env_args(max_sampling_frequency=a.max_sampling_frequency, time_step=a.time_step,
                          radius_of_acceptance=a.radius_of_acceptance, lookahead_distance=a.lookahead_distance,
                          collav_mode=a.collav_mode)
    env, _ = S.prepare_multiship_rl_env(args, num_envs=a.num_envs)
    dev = env.obs_buf.device
    wrapped = NormalizedBoxEnv(env, reward_scale=a.reward_scale)
    M = a.layer_size
    qf1, qf2, tq1, tq2 = (ConcatMlp([M, M], 1, 9).to(dev) for _ in range(4))
    policy = TanhGaussianPolicy([M, M], obs_dim=8, action_dim=1).to(dev)
    buf = GpuReplayBuffer(a.replay_buffer_size, env=wrapped, seed=None if a.cuda_graph else a.seed)
    trainer = SACTrainer(wrapped, policy, qf1, qf2, tq1, tq2, discount=a.discount, reward_scale=a.reward_scale,
                         policy_lr=a.policy_lr, qf_lr=a.qf_lr, soft_target_tau=a.soft_target_tau,
                         target_update_period=a.target_update_period, action_reg_coeff=a.action_reg_coeff, clip_val=a.clip_val,
                         capturable=bool(a.cuda_graph))
    expl = VectorizedPathCollector(wrapped, policy, replay_buffer=buf)
    evalc = VectorizedPathCollector(wrapped, MakeDeterministic(policy))
    run_logger = None
    if a.do_logging:
        from ast_sac_b200.rl.logging import logger as run_logger, setup_logger
        variant = dict(   # run/ast-sac_runner.py:211-236
            algorithm="SAC", version="normal", layer_size=a.layer_size, replay_buffer_size=a.replay_buffer_size,
            algorithm_kwargs=dict(num_epochs=a.num_epochs, num_eval_steps_per_epoch=a.num_eval_steps_per_epoch,
                                  num_trains_per_train_loop=a.num_trains_per_train_loop,
                                  num_expl_steps_per_train_loop=a.num_expl_steps_per_train_loop,
                                  min_num_steps_before_training=a.min_num_steps_before_training,
                                  max_path_length=a.max_path_length, batch_size=a.batch_size),
            trainer_kwargs=dict(discount=a.discount, soft_target_tau=a.soft_target_tau,
                                target_update_period=a.target_update_period, policy_lr=a.policy_lr, qf_lr=a.qf_lr,
                                reward_scale=a.reward_scale, use_automatic_entropy_tuning=True,
                                action_reg_coeff=a.action_reg_coeff, clip_val=a.clip_val))
        run_logger.set_print(False)
        run_dir = setup_logger('ast-sac_maritime_logs', variant=variant, base_log_dir=a.base_log_dir,
                               snapshot_mode=a.snapshot_mode, seed=a.seed)
        print(json.dumps({"run_dir": run_dir}), flush=True)
    c0 = env.total_substeps()
    t0 = time.perf_counter()
    alg = BatchRLAlgorithm(trainer, expl, evalc, buf, batch_size=a.batch_size, max_path_length=a.max_path_length,
                           num_epochs=a.num_epochs, num_eval_steps_per_epoch=a.num_eval_steps_per_epoch,
                           num_expl_steps_per_train_loop=a.num_expl_steps_per_train_loop,
                           num_trains_per_train_loop=a.num_trains_per_train_loop,
                           min_num_steps_before_training=a.min_num_steps_before_training,
                           log=lambda s: print(json.dumps(s), flush=True), use_cuda_graph=bool(a.cuda_graph),
                           logger=run_logger)
    hist = alg.train()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sim_steps = env.total_substeps() - c0
    t_roll = sum(h['time/evaluation sampling (s)'] + h['time/exploration sampling (s)'] for h in hist) + \
        sum(h.get('initial exploration (s)', 0.0) for h in hist)
    t_train = sum(h['time/training (s)'] for h in hist)
    print(json.dumps({"summary": True, "wall_s": wall, "num_envs": a.num_envs, "collav_mode": a.collav_mode,
                      "rl_transitions": expl._num_steps_total + evalc._num_steps_total, "simulator_steps": sim_steps,
                      "rollout_s": t_roll, "training_s": t_train,
                      "env_steps_per_s_in_rollouts": sim_steps / max(t_roll, 1e-9),
                      "sac_updates_per_s": trainer._n_train_steps_total / max(t_train, 1e-9),
                      "cuda_graph": bool(a.cuda_graph)}), flush=True)
    env.close()


if __name__ == "__main__":
    main()

"""Batched consumers of the environment: SURVEY.md section 8(f) rows 2 and 3 and BASELINE config 5.

    normalized_env.NormalizedBoxEnv      ast_sac/env_wrapper/normalized_box_env.py:46-61, batched
    replay_buffer.GpuReplayBuffer        ast_sac/data_management/{simple,env}_replay_buffer.py on device tensors
    sampler.VectorizedPathCollector      ast_sac/samplers/data_collector/{rollout_functions,path_collector}.py, batched
    sac.SACTrainer + networks            ast_sac/torch/sac/sac.py on stock PyTorch (not a roofline target)
    algorithm.BatchRLAlgorithm           ast_sac/core/batch_rl_algorithm.py:47-106
"""
from .normalized_env import NormalizedBoxEnv  # noqa: F401
from .replay_buffer import GpuReplayBuffer  # noqa: F401
from .sampler import VectorizedPathCollector, batched_ast_sac_rollout  # noqa: F401
from .sac import ConcatMlp, MakeDeterministic, SACTrainer, TanhGaussianPolicy  # noqa: F401
from .algorithm import BatchRLAlgorithm  # noqa: F401

"""Multi-GPU plumbing: environments are independent, so they shard across ranks by contiguous env
index with NO per-step communication.  The only collective is an all-gather of a small fixed-size
episode-statistics record, NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

N_EVENTS = 11
STATS_LEN = N_EVENTS + 5   # per-event counts, episodes, env-steps, sum return, sum length(rl steps), max substeps


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of env indices owned by ``rank``: env b lives on rank floor(b * G / B)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = (total_envs * rank + world_size - 1) // world_size
    hi = (total_envs * (rank + 1) + world_size - 1) // world_size
    return lo, hi


def episode_stats(event_bits: torch.Tensor, episode_return: torch.Tensor, episode_env_steps: torch.Tensor,
                  episode_rl_steps: torch.Tensor) -> torch.Tensor:
    """Pack this rank's finished-episode statistics into one float64 vector of length STATS_LEN."""
    dev = event_bits.device
    out = torch.zeros(STATS_LEN, dtype=torch.float64, device=dev)
    bits = event_bits.to(torch.int64)
    for i in range(N_EVENTS):
        out[i] = ((bits >> i) & 1).sum()
    out[N_EVENTS + 0] = bits.numel()
    out[N_EVENTS + 1] = episode_env_steps.sum()
    out[N_EVENTS + 2] = episode_return.sum()
    out[N_EVENTS + 3] = episode_rl_steps.sum()
    out[N_EVENTS + 4] = episode_env_steps.max() if episode_env_steps.numel() else 0
    return out


def gather_stats(local: torch.Tensor) -> torch.Tensor:
    """all_gather of the per-rank statistics vectors -> [world_size, STATS_LEN] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.unsqueeze(0).clone()
    world = dist.get_world_size()
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous().view(-1))
    return out.view(world, local.numel())


def summarise(gathered: torch.Tensor) -> Dict[str, float]:
    tot = gathered[:, :N_EVENTS + 4].sum(dim=0).tolist()
    episodes = max(tot[N_EVENTS], 1.0)
    return {
        "episodes": tot[N_EVENTS], "env_steps": tot[N_EVENTS + 1], "mean_return": tot[N_EVENTS + 2] / episodes,
        "mean_rl_steps": tot[N_EVENTS + 3] / episodes, "event_counts": tot[:N_EVENTS],
        "max_env_steps_per_episode": float(gathered[:, N_EVENTS + 4].max().item()),
    }

"""Probe: bare ship loop (k_ship_rollout, no env logic) rate vs the env kernel's, same ship model."""
import sys, time
import torch
sys.path.insert(0, ".")
from ast_sac_b200 import scenarios as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
K = 200
for kind in ("colav", "rl"):
    args = S.get_env_args(time_step=4)
    if kind == "rl":
        env, _ = S.prepare_multiship_rl_env(args, num_envs=B)
    else:
        env, _ = S.prepare_colav_env(args, iw=True, num_envs=B)
    env.reset()
    env.ship_rollout(K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.ship_rollout(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{kind}: bare rollout {2 * B * K / ms / 1e6:.2f} G ship-steps/s = {B * K / ms / 1e6:.2f} G pair-steps/s ({ms:.3f} ms)")
    env.reset()
    env._step(K)          # first launch of this instantiation (lazy module load)
    env.reset()
    torch.cuda.synchronize()
    e0.record(); env._step(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n = int(env.nsub_buf.sum().item())
    print(f"{kind}: _step({K}) {n / ms / 1e6:.2f} G env-steps/s ({ms:.3f} ms, {n} steps)")
    env.close()

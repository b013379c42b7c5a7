"""Caller side of the hot path, batched: the reference's RandomController and the rollout loop of its runners.

``controller/random/RandomController.py:12`` returns the density map ``s0 + s1 - 10 s2 + s3`` of the observation it is
handed; ``runner/checkRL.py:33-36`` loops ``make_action -> step`` until the episode ends.  Here both run for B
environments at once and nothing leaves HBM: the map is formed by torch from the observation tensor, decoded by
``wrsn_decode_density_map`` and consumed by ``wrsn_rollout_step`` (which also resets finished episodes, as the trainers'
loops do: ``controller/ippo/IPPO.py:137-143``).  Trainers (PPO / IPPO) plug in at the same place: anything with a
``make_action(agent_id, state)`` that returns either [B, S, S] maps or [B, 3] actions.
"""
import torch

from .sharding import reduce_stats


class BatchedRandomController:
    """``RandomController.make_action`` for a batch: ``state`` is the [B, 4, S, S] observation tensor."""

    def make_action(self, agent_id, state, info=None, wrsn=None):
        return state[:, 0] + state[:, 1] - 10.0 * state[:, 2] + state[:, 3]


def rollout(env, controller, steps, obs=None, group=None):
    """``steps`` rollout steps of every environment of ``env`` (a ``BatchedWRSN``) under ``controller``.

    Returns the observation tensor and job-wide statistics (summed over ranks when torch.distributed is initialised:
    the only collective of the whole path): decisions handed out, simulated seconds, episodes finished, sum of rewards.
    """
    B, S, dev = env.B, env.S, env.device
    if obs is None:
        obs = torch.zeros((B, 4, S, S), dtype=torch.float32, device=dev)
        env.get_state(out=obs)
    action = torch.zeros((B, 3), dtype=torch.float64, device=dev)
    local = torch.zeros(2, dtype=torch.float64, device=dev)               # episodes finished, sum of rewards
    resets0 = env.req.stats[:, 2].sum()
    dec0, sim0 = reduce_stats(env, group)
    for _ in range(int(steps)):
        a = controller.make_action(env.req.agent_id, obs)
        if a.dim() == 3:                                                  # density maps: WRSN.step :293-297 on the device
            env.density_map_to_action(a.contiguous(), out=action)
        else:
            action.copy_(a)
        env.rollout_step(action, obs)
        local[1] += torch.nan_to_num(env.req.reward, nan=0.0).sum()
    local[0] = env.req.stats[:, 2].sum() - resets0                        # episodes that ended (and were reset) in these steps
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(local, group=group)
    dec1, sim1 = reduce_stats(env, group)
    return obs, dict(decisions=dec1 - dec0, simulated_seconds=sim1 - sim0, episodes=float(local[0].item()),
                     reward_sum=float(local[1].item()))

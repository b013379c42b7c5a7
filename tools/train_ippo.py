"""The reference's ``runner/IPPO.py`` / ``runner/PPO.py`` on the batched simulator: same scenario / charger / ``alg_args`` YAML
files, the reference's own ``UNet`` actor and ``CNNCritic`` (imported from a checkout of the reference, which is not part of
this repository), B environments per GPU instead of one.

    python tools/train_ippo.py --reference-root /path/to/multi_agent_rl_wrsn --envs 2048 --iterations 1000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_ippo.py --reference-root ... --envs 16384

Under torchrun the ``--envs`` environments are split into contiguous blocks, one per rank / GPU (``sharding.shard_range``); the
simulation has no collective, the update averages its gradients once per minibatch over NCCL.
"""
import argparse
import os
import random
import sys

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200 import BatchedIPPO, BatchedWRSN  # noqa: E402
from multi_agent_rl_wrsn_b200.sharding import shard_range  # noqa: E402


def main():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--reference-root", required=True, help="checkout of the reference (for controller/ppo/actor|critic and the YAML files)")
    p.add_argument("--scenario", default="physical_env/network/network_scenarios/hanoi1000n50.yaml")   # runner/IPPO.py:19
    p.add_argument("--agent-type", default="physical_env/mc/mc_types/default.yaml")
    p.add_argument("--alg-args", default="alg_args/ippo.yaml")
    p.add_argument("--num-agent", type=int, default=3)
    p.add_argument("--envs", type=int, default=2048, help="environments of the whole job")
    p.add_argument("--window", type=int, default=8, help="rollout steps per collection window")
    p.add_argument("--iterations", type=int, default=1000)
    p.add_argument("--save-folder", default="save_model/ippo")
    p.add_argument("--model-path", default=None)
    p.add_argument("--ppo", action="store_true", help="one shared network pair (controller/ppo/PPO.py) instead of one per agent")
    a = p.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    random.seed(0); np.random.seed(0); torch.manual_seed(0)                 # runner/IPPO.py:12-15 (identical replicas on every rank)
    torch.backends.cudnn.deterministic = True

    sys.path.insert(0, a.reference_root)
    from controller.ppo.actor.UnetActor import UNet                        # the reference's networks, unmodified
    from controller.ppo.critic.CNNCritic import CNNCritic
    ref = lambda rel: rel if os.path.isabs(rel) else os.path.join(a.reference_root, rel)
    with open(ref(a.alg_args)) as f:
        args = yaml.safe_load(f)["alg_args"]
    lo, hi = shard_range(a.envs, rank, world)
    env = BatchedWRSN([ref(a.scenario)], num_agent=a.num_agent, mc_type=ref(a.agent_type), num_envs=hi - lo, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)                                               # different action samples on every shard
    trainer = BatchedIPPO(args, env, device=dev, model_path=a.model_path, actor_factory=UNet, critic_factory=CNNCritic,
                          window=a.window, shared=a.ppo, generator=gen)
    trainer.train(a.iterations, a.save_folder)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

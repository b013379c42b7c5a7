"""BatchedIPPO (the reference's IPPO trainer on a BatchedWRSN) on CPU through the host emulation:
an iteration runs end to end, the checkpoint files have the reference's layout (controller/ippo/IPPO.py:296-309) and a
trainer restarted from them (IPPO.py:50-64) continues its counters.  Actors emit 3-vector actions here (the emulation has
no density-map decoder); the map path is the same code with action_shape=None and is exercised on the GPU."""
import csv
import os
import subprocess

import numpy as np
import pytest

from tests import helpers
import torch

from multi_agent_rl_wrsn_b200 import BatchedIPPO, BatchedWRSN, _lib, synthetic
from tests.helpers import REPO

EMU_DIR = os.path.join(REPO, "tests", "emu")
S = 16
ARGS = dict(seed=0, lr=3.0e-4, gamma=0.99, clip=0.2, batch_size=24, n_updates_per_iteration=2, save_freq=1, gae=True,
            norm_adv=True, minibatch_size=8, ent_coef=0.0, vf_coef=0.5, gae_lambda=0.95, max_grad_norm=0.5, clip_vloss=True)


@pytest.fixture()
def emu_library():
    subprocess.check_call(["make", "-C", EMU_DIR, "libwrsn_emu.so"], stdout=subprocess.DEVNULL)
    helpers.use_host_build()
    yield
    helpers.use_cuda_build_lazy()


class Actor(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = torch.nn.Linear(4 * S * S, 3)
        self.log_std = torch.nn.Parameter(torch.full((1, 3), -2.0))

    def forward(self, x):
        mean = torch.sigmoid(self.fc(x.flatten(1))) * torch.tensor([1.0, 1.0, 0.05])
        return mean, self.log_std.expand_as(mean)


class Critic(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = torch.nn.Linear(4 * S * S, 1)

    def forward(self, x):
        return self.fc(x.flatten(1))


def _params(nets):
    return torch.cat([p.detach().reshape(-1) for n in nets for p in n.parameters()])


def test_train_iteration_checkpoint_and_resume(emu_library, tmp_path):
    torch.manual_seed(0)
    scs = [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]
    env = BatchedWRSN(scs, num_agent=3, num_envs=6, device="cpu", map_size=S)
    gen = torch.Generator().manual_seed(1)
    trainer = BatchedIPPO(ARGS, env, model_path=None, actor_factory=Actor, critic_factory=Critic, window=6,
                          action_shape=(3,), generator=gen)
    before = _params(trainer.actors + trainer.critics)
    hist = trainer.train(0, str(tmp_path / "save"))                      # the reference's loop runs iterations + 1 times
    assert len(hist) == 3 and all(h["iteration"] == 1 for h in hist)
    assert not torch.equal(before, _params(trainer.actors + trainer.critics))
    assert min(trainer.last_rollout["transitions"]) >= ARGS["batch_size"]
    for i in range(3):
        folder = tmp_path / "save" / "1" / str(i)
        assert sorted(os.listdir(folder)) == ["actor.pth", "critic.pth", "log.csv"]
        rows = list(csv.reader(open(folder / "log.csv")))
        assert len(rows) == 1 and rows[0][0] == "1" and rows[0][1] == str(ARGS["batch_size"]) and len(rows[0]) == 7
        sd = torch.load(folder / "actor.pth")
        assert all(torch.equal(sd[k], v) for k, v in trainer.actors[i].state_dict().items())

    env2 = BatchedWRSN(scs, num_agent=3, num_envs=6, device="cpu", map_size=S)
    resumed = BatchedIPPO(ARGS, env2, model_path=str(tmp_path / "save" / "1"), actor_factory=Actor, critic_factory=Critic,
                          window=6, action_shape=(3,), generator=gen)
    assert torch.equal(_params(resumed.actors + resumed.critics), _params(trainer.actors + trainer.critics))
    assert all(lg["i_so_far"] == 1 and lg["t_so_far"] == ARGS["batch_size"] for lg in resumed.loggers)
    resumed.train(0, str(tmp_path / "save"))
    rows = list(csv.reader(open(tmp_path / "save" / "2" / "0" / "log.csv")))
    assert [r[0] for r in rows] == ["1", "2"]                           # the old log travels with the new checkpoint


def test_shared_network_is_ppo(emu_library, tmp_path):
    """shared=True = controller/ppo/PPO.py: one network pair, one pooled batch and one update per iteration, flat checkpoint folder."""
    scs = [synthetic(num_nodes=40, num_targets=120, seed=3, num_gateways=2)]
    env = BatchedWRSN(scs, num_agent=2, num_envs=4, device="cpu", map_size=S)
    t = BatchedIPPO(ARGS, env, actor_factory=Actor, critic_factory=Critic, window=4, action_shape=(3,), shared=True)
    assert t.actors[0] is t.actors[1] and t.critics[0] is t.critics[1] and t.optimizers[0] is t.optimizers[1]
    hist = t.train(0, str(tmp_path / "ppo"))
    assert len(hist) == 1                                                # one update per iteration, not one per agent
    assert sum(t.last_rollout["transitions"]) >= ARGS["batch_size"]
    assert sorted(os.listdir(tmp_path / "ppo" / "1")) == ["actor.pth", "critic.pth", "log.csv"]
    env2 = BatchedWRSN(scs, num_agent=2, num_envs=4, device="cpu", map_size=S)
    r = BatchedIPPO(ARGS, env2, model_path=str(tmp_path / "ppo" / "1"), actor_factory=Actor, critic_factory=Critic, window=4,
                    action_shape=(3,), shared=True)
    assert torch.equal(_params([r.actors[0], r.critics[0]]), _params([t.actors[0], t.critics[0]]))
    assert r.loggers[0]["i_so_far"] == 1


def _ddp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from multi_agent_rl_wrsn_b200.sharding import shard_range, shard_scenario_index
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    helpers.use_host_build()
    torch.manual_seed(0)                                                 # identical initial replicas
    scs = [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]
    lo, hi = shard_range(8, rank, world)
    env = BatchedWRSN(scs, num_agent=2, num_envs=hi - lo, device="cpu", map_size=S,
                      scenario_index=shard_scenario_index(8, len(scs), rank, world))
    a = dict(ARGS, batch_size=16, minibatch_size=8)
    t = BatchedIPPO(a, env, actor_factory=Actor, critic_factory=Critic, window=6, action_shape=(3,),
                    generator=torch.Generator().manual_seed(100 + rank))   # different samples, different shards
    t.train(0, os.path.join(out_dir, "save"))
    torch.save(dict(p=_params(t.actors + t.critics), dec=t.last_rollout["decisions"]), os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_keep_identical_replicas(tmp_path):
    """N > 1: every rank rolls out its own shard of environments; the gradient all-reduce keeps the replicas identical and
    only rank 0 writes checkpoints."""
    import socket
    import torch.multiprocessing as mp
    subprocess.check_call(["make", "-C", EMU_DIR, "libwrsn_emu.so"], stdout=subprocess.DEVNULL)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(2))
    assert torch.equal(r0["p"], r1["p"])
    torch.manual_seed(0)
    fresh = _params([Actor(), Actor(), Critic(), Critic()])
    assert r0["p"].shape == fresh.shape and not torch.equal(r0["p"], fresh)
    assert sorted(os.listdir(tmp_path / "save" / "1")) == ["0", "1"]


# ------------------------------------------------------------------ on the device, with the real networks
def test_in_tree_networks_have_the_reference_shapes():
    from multi_agent_rl_wrsn_b200.nets import CNNCritic, UNetActor, num_parameters
    a, c = UNetActor(), CNNCritic()
    assert num_parameters(a) == 936401 and num_parameters(c) == 1147513
    x = torch.rand(2, 4, 100, 100)
    mean, log_std = a(x)
    assert mean.shape == (2, 100, 100) and log_std.shape == (2, 100, 100) and c(x).shape == (2, 1)
    keys = set(a.state_dict())
    assert {"inc.conv.weight", "down1.conv_block.bn.running_mean", "up2.conv_block.conv.bias", "out_mean.conv.weight", "log_std"} <= keys
    assert set(c.state_dict()) == {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
                                   "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"}


@pytest.mark.gpu
def test_ippo_iteration_on_the_device_with_unet_and_cnn_critic(tmp_path):
    """BatchedIPPO.roll_out + ppo_update (IPPO.py:119-310) on the B200 with the U-Net actors / CNN critics and density-map
    actions decoded on the device, step budget on (rows whose step is in flight carry no request): every agent gets its
    batch, the update changes the weights, losses are finite, the checkpoint loads back."""
    helpers.use_cuda_build()
    torch.manual_seed(0)
    dev = "cuda:0"
    args = dict(ARGS, batch_size=64, minibatch_size=32, n_updates_per_iteration=2, save_freq=1)
    scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + k) for k in range(4)]
    env = BatchedWRSN(scs, num_agent=3, num_envs=96, device=dev, step_budget=60)
    gen = torch.Generator(device=dev).manual_seed(1)
    t = BatchedIPPO(args, env, window=4, generator=gen)
    before = _params(t.actors + t.critics).clone()
    hist = t.train(0, str(tmp_path / "save"))
    assert len(hist) == 3 and all(np.isfinite(h["loss"]) and np.isfinite(h["approx_kl"]) for h in hist)
    assert min(t.last_rollout["transitions"]) >= 64 and not torch.equal(before, _params(t.actors + t.critics))
    assert float(env.hdr("ERR").max()) == 0.0
    env2 = BatchedWRSN(scs, num_agent=3, num_envs=8, device=dev)
    r = BatchedIPPO(args, env2, model_path=str(tmp_path / "save" / "1"), window=2, generator=gen)
    assert torch.equal(_params(r.actors + r.critics), _params(t.actors + t.critics))

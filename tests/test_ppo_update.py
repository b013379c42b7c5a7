"""controllers.ppo_update (the update of IPPO.train / PPO.train, controller/ippo/IPPO.py:222-268) on CPU: one minibatch
step against the loss written out from the reference's formulas with torch.distributions, and the N > 1 path — two gloo
ranks with different batches keep identical replicas, equal to a single process that averages the two gradients itself."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_rl_wrsn_b200.controllers import ppo_update

S = 10
ARGS = dict(lr=3.0e-4, gamma=0.99, clip=0.2, batch_size=16, n_updates_per_iteration=1, gae=True, norm_adv=True,
            minibatch_size=16, ent_coef=0.01, vf_coef=0.5, gae_lambda=0.95, max_grad_norm=0.5, clip_vloss=True)


class Actor(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(4, 1, 3, padding=1)
        self.log_std = torch.nn.Parameter(torch.full((1, 1, S, S), -0.5))

    def forward(self, x):
        mean = self.conv(x)
        return mean.squeeze(), self.log_std.expand_as(mean).squeeze()


class Critic(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = torch.nn.Linear(4 * S * S, 1)

    def forward(self, x):
        return self.fc(x.flatten(1))


def _models():
    torch.manual_seed(0)
    return Actor(), Critic()


def _batch(seed, n=16):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(states=torch.rand(n, 4, S, S, generator=g), actions=r(n, S, S), log_probs=-90.0 + r(n), advantages=r(n),
                returns=r(n), values=r(n))


def _reference_grads(actor, critic, b, a):
    """Loss of IPPO.py:233-262 written out literally for one minibatch = the whole batch, in order."""
    from torch.distributions.normal import Normal
    mean, log_std = actor(b["states"])
    d = Normal(mean, torch.exp(log_std))
    newlogprob, entropy = d.log_prob(b["actions"]).sum((1, 2)), d.entropy().sum((1, 2))
    newvalue = torch.squeeze(critic(b["states"]).sum(1)).view(-1)
    ratio = (newlogprob - b["log_probs"]).exp()
    adv = (b["advantages"] - b["advantages"].mean()) / (b["advantages"].std() + 1e-8)
    pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - a["clip"], 1 + a["clip"])).mean()
    v_clipped = b["values"] + torch.clamp(newvalue - b["values"], -a["clip"], a["clip"])
    v_loss = 0.5 * torch.max((newvalue - b["returns"]) ** 2, (v_clipped - b["returns"]) ** 2).mean()
    loss = pg_loss - a["ent_coef"] * entropy.mean() + v_loss * a["vf_coef"]
    for p in list(actor.parameters()) + list(critic.parameters()):
        p.grad = None
    loss.backward()
    return loss.detach()


def _apply(actor, critic, opt, a):
    torch.nn.utils.clip_grad_norm_(actor.parameters(), a["max_grad_norm"])
    torch.nn.utils.clip_grad_norm_(critic.parameters(), a["max_grad_norm"])
    opt.step()


def _params(*mods):
    return torch.cat([p.detach().reshape(-1) for m in mods for p in m.parameters()])


def test_one_minibatch_equals_reference_formulas():
    b = _batch(1)
    actor, critic = _models()
    opt = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=ARGS["lr"])
    actor2, critic2 = copy.deepcopy(actor), copy.deepcopy(critic)
    opt2 = torch.optim.Adam(list(actor2.parameters()) + list(critic2.parameters()), lr=ARGS["lr"])
    st = ppo_update(actor, critic, opt, b, ARGS)
    loss = _reference_grads(actor2, critic2, b, ARGS)
    _apply(actor2, critic2, opt2, ARGS)
    assert torch.allclose(st["loss"], loss, rtol=1e-6, atol=1e-7)
    assert torch.allclose(_params(actor, critic), _params(actor2, critic2), rtol=1e-6, atol=1e-8)
    assert not torch.equal(_params(actor, critic), _params(*_models()))


def test_minibatches_and_passes_cover_the_batch():
    b = _batch(2, n=24)
    actor, critic = _models()
    opt = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=ARGS["lr"])
    a = dict(ARGS, minibatch_size=16, n_updates_per_iteration=3, clip_vloss=False, norm_adv=False)
    st = ppo_update(actor, critic, opt, b, a, generator=torch.Generator().manual_seed(0))
    assert st["minibatch"] == 8 and opt.state[next(actor.parameters())]["step"] == 6     # ceil(24 / 16) * 3 steps
    assert 0.0 <= float(st["clipfrac"]) <= 1.0


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    actor, critic = _models()
    opt = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=ARGS["lr"])
    ppo_update(actor, critic, opt, _batch(10 + rank), ARGS)
    torch.save(_params(actor, critic), os.path.join(out_dir, "p%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_average_gradients(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = (torch.load(os.path.join(str(tmp_path), "p%d.pt" % r)) for r in range(2))
    assert torch.equal(p0, p1)                                                           # replicas stay identical
    actor, critic = _models()
    opt = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=ARGS["lr"])
    params = list(actor.parameters()) + list(critic.parameters())
    grads = []
    for r in range(2):
        _reference_grads(actor, critic, _batch(10 + r), ARGS)
        grads.append([p.grad.clone() for p in params])
    for p, g0, g1 in zip(params, *grads):
        p.grad = (g0 + g1) / 2
    _apply(actor, critic, opt, ARGS)
    assert torch.allclose(_params(actor, critic), p0, rtol=1e-6, atol=1e-8)

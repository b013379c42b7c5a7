"""``BatchedIPPO`` — the reference's IPPO trainer (``controller/ippo/IPPO.py``) on a ``BatchedWRSN``.

Same constructor arguments (``args`` = the ``alg_args`` mapping of ``alg_args/ippo.yaml``, ``env``, ``device``,
``model_path``), same ``train(trained_iterations, save_folder)`` and the same files on disk
(``<save_folder>/<iteration>/<agent>/actor.pth | critic.pth | log.csv``, ``IPPO.py:296-309``; ``model_path`` is such an
``<iteration>`` folder, ``:50-64``), so checkpoints move freely between the two.  What changes is where the work happens:
``roll_out`` runs every environment of the batch at once (``controllers.IPPORollout``: observations, maps and the
transition record never leave HBM), the advantage recursion covers all (environment, episode) sequences in one pass,
and with several ranks (one process per GPU, each with its own shard of environments) the gradients are averaged by a
single all-reduce per minibatch (``controllers.ppo_update``).

``actor_factory`` / ``critic_factory`` build one actor / critic per agent: by default ``nets.UNetActor`` / ``nets.CNNCritic``,
written from the shapes of the reference's ``UNet`` and ``CNNCritic`` (``controller/ppo/actor/UnetActor.py``,
``controller/ppo/critic/CNNCritic.py``; same parameter names, so checkpoints are interchangeable) — or the reference's own
classes, or anything with the same interface (actor: obs -> (mean, log_std) maps; critic: obs -> [n, k] summed over k).

Several ranks (one process per GPU, ``torch.distributed`` initialised by the caller): every rank holds its own shard of
environments and an identical replica of the networks — the constructor broadcasts rank 0's parameters, so replicas are
identical whatever the caller seeded; every rank feeds ITS ``batch_size`` transitions per agent into an update, gradients are
averaged (one all-reduce per minibatch), i.e. the effective batch is ``world_size * batch_size`` while ``t_so_far`` counts
``batch_size`` per iteration as the reference's log does; the episode statistics of ``log.csv`` are summed over ranks.
``shared=True`` is the reference's ``PPO`` (``controller/ppo/PPO.py``): ONE actor / critic pair for all chargers, the
transitions of all agents pooled into one batch of ``batch_size`` (``PPO.py:125-197``), one update per iteration and the flat
checkpoint folder ``<save_folder>/<iteration>/actor.pth | critic.pth | log.csv`` (``PPO.py:281-294``, ``:48-55``).
"""
import csv
import os
import shutil
import time

import torch

from .controllers import IPPORollout, PerAgentPolicy, ppo_update, select_batch


class BatchedIPPO:
    def __init__(self, args, env, device=None, model_path=None, actor_factory=None, critic_factory=None, window=8,
                 shared=False, group=None, generator=None, action_shape=None):
        if actor_factory is None or critic_factory is None:
            from .nets import CNNCritic, UNetActor
            actor_factory = actor_factory or (lambda: UNetActor(map_size=env.S))
            critic_factory = critic_factory or (lambda: CNNCritic(map_size=env.S))
        # window: rollout steps per collection window; the record holds (window + 1) x B observations (160 KB each at map size
        # 100) plus, for the decisions still open at a window end, B x num_agent more (keep_open: nothing is lost at the boundary).
        self.env, self.args, self.group, self.generator = env, dict(args), group, generator
        self.num_agent = env.num_agent
        self.device = torch.device(device) if device is not None else env.device
        self.batch_size, self.save_freq = int(args["batch_size"]), int(args["save_freq"])
        n_nets = 1 if shared else self.num_agent
        nets_a = [actor_factory().to(self.device) for _ in range(n_nets)]
        nets_c = [critic_factory().to(self.device) for _ in range(n_nets)]
        self.actors = [nets_a[0 if shared else i] for i in range(self.num_agent)]
        self.critics = [nets_c[0 if shared else i] for i in range(self.num_agent)]
        self.shared = bool(shared)
        self.log_file = [None] * self.num_agent
        self.loggers = [dict(i_so_far=0, t_so_far=0, losses=[], delta_t=time.time_ns()) for _ in range(self.num_agent)]
        if model_path is not None:                                        # IPPO.py:50-64 / PPO.py:48-55
            for agent_folder in ([None] if shared else sorted(os.listdir(model_path))):
                i = 0 if shared else int(agent_folder)
                path = model_path if shared else os.path.join(model_path, agent_folder)
                self.critics[i].load_state_dict(torch.load(os.path.join(path, "critic.pth"), map_location=self.device))
                self.actors[i].load_state_dict(torch.load(os.path.join(path, "actor.pth"), map_location=self.device))
                self.log_file[i] = os.path.join(path, "log.csv")
                with open(self.log_file[i], "r") as f:
                    rows = list(csv.reader(f))
                if rows:
                    self.loggers[i]["i_so_far"], self.loggers[i]["t_so_far"] = int(rows[-1][0]), int(rows[-1][1])
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for net in nets_a + nets_c:                                   # identical replicas: rank 0's parameters and buffers
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.optimizers = []
        for i in range(self.num_agent):                                   # one Adam over actor + critic, IPPO.py:65-69
            if shared and i > 0:
                self.optimizers.append(self.optimizers[0])
                continue
            params = list(self.actors[i].parameters()) + list(self.critics[i].parameters())
            self.optimizers.append(torch.optim.Adam(params, lr=float(args["lr"])))
        # action_shape: None = S x S density maps decoded on the device (the runners' density_map=True); (3,) = direct actions
        self.rollout = IPPORollout(env, int(window), action_shape=action_shape, keep_open=True)
        self.policy = PerAgentPolicy(self.actors, generator=generator, action_shape=action_shape)
        self.last_rollout = {}

    def get_value(self, agent_id, state):                                 # IPPO.py:117-119
        return self.critics[agent_id](state.to(torch.float32)).sum(1)

    def roll_out(self):
        """``IPPO.roll_out`` (``:119-210``): windows of rollout steps until every agent has ``batch_size`` transitions, returns
        and advantages per (environment, episode) sequence, then the reward-outlier selection of ``batch_size`` rows."""
        a, ro = self.args, self.rollout
        acc = [dict(states=[], actions=[], log_probs=[], rewards=[], advantages=[], returns=[], values=[]) for _ in range(self.num_agent)]
        counts = [0] * self.num_agent
        st0 = self.env.req.stats.sum(0).clone()
        while (sum(counts) if self.shared else min(counts)) < self.batch_size:        # PPO.py:126 / IPPO.py:183-189
            ro.carry_over()
            ro.collect(self.policy)
            for i in range(self.num_agent):
                returns, advantages, values, bt = ro.cal_rt_adv(i, lambda s, i=i: self.get_value(i, s), float(a["gamma"]),
                                                                float(a["gae_lambda"]), gae=bool(a["gae"]))
                n = int(bt["rewards"].shape[0])
                if n == 0:
                    continue
                counts[i] += n
                for k, v in (("states", bt["states"]), ("actions", bt["actions"]), ("log_probs", bt["log_probs"]),
                             ("rewards", bt["rewards"]), ("advantages", advantages), ("returns", returns), ("values", values)):
                    acc[i][k].append(v)
        delta = (self.env.req.stats.sum(0) - st0).clone()
        all_dead = self.env.raise_on_error()                               # engine errors of any step of these windows, reset or not
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(delta, group=self.group)                      # job-wide statistics for log.csv
        self.last_rollout = dict(decisions=float(delta[0]), simulated_seconds=float(delta[1]), episodes=float(delta[2]),
                                 transitions=list(counts), environments_that_saw_every_charger_dead=all_dead)
        if self.shared:                                                  # one pooled batch, agents in id order (PPO.py:164-175)
            acc = [{k: [x for i in range(self.num_agent) for x in acc[i][k]] for k in acc[0]}]
        out = []
        for i in range(len(acc)):
            full = {k: torch.cat(v) for k, v in acc[i].items()}
            idx = select_batch(full["rewards"], self.batch_size, generator=self.generator)     # :193-200
            sel = {k: v[idx.to(v.device)] for k, v in full.items()}
            sel["mean_reward_all"] = float(full["rewards"].mean())
            out.append(sel)
        return out

    def train(self, trained_iterations, save_folder):
        """``IPPO.train`` (``:212-310``): roll out, update every agent, log, save every ``save_freq`` iterations.  With several
        ranks only rank 0 writes files."""
        dist = torch.distributed
        rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank(self.group) == 0
        logs = [[] for _ in range(self.num_agent)]
        i_so_far = 0
        history = []
        while i_so_far <= trained_iterations:                                                  # sic: one more than asked (:220)
            batches = self.roll_out()
            i_so_far += 1
            for i in range(1 if self.shared else self.num_agent):
                lg = self.loggers[i]
                lg["t_so_far"] += self.batch_size
                lg["i_so_far"] += 1
                stats = ppo_update(self.actors[i], self.critics[i], self.optimizers[i], batches[i], self.args, group=self.group,
                                   generator=self.generator)
                row = self._log_summary(i, stats, batches[i]["mean_reward_all"])
                logs[i].append(row)
                history.append(dict(agent=i, iteration=lg["i_so_far"], **{k: (float(v) if v is not None else None)
                                                                          for k, v in stats.items()}))
                if lg["i_so_far"] % self.save_freq == 0 and rank0:                               # :296-309
                    folder = os.path.join(save_folder, str(lg["i_so_far"])) if self.shared else \
                        os.path.join(save_folder, str(lg["i_so_far"]), str(i))
                    os.makedirs(folder, exist_ok=True)
                    torch.save(self.actors[i].state_dict(), os.path.join(folder, "actor.pth"))
                    torch.save(self.critics[i].state_dict(), os.path.join(folder, "critic.pth"))
                    if self.log_file[i] is not None and os.path.abspath(os.path.dirname(self.log_file[i])) != os.path.abspath(folder):
                        shutil.copy(self.log_file[i], folder)
                    with open(os.path.join(folder, "log.csv"), "a", newline="") as f:
                        w = csv.writer(f)
                        for r in logs[i]:
                            w.writerow(r)
                    logs[i] = []
                    self.log_file[i] = os.path.join(folder, "log.csv")
        return history

    def _log_summary(self, i, stats, mean_reward):
        """The row of ``IPPO._log_summary`` (``:312-349``): iteration, timesteps, mean episode length (decisions per finished
        episode), mean episode lifetime (simulated seconds per finished episode + warm-up), loss, mean reward, seconds."""
        lg, lr = self.loggers[i], self.last_rollout
        now = time.time_ns()
        delta_t = (now - lg["delta_t"]) / 1e9
        lg["delta_t"] = now
        ep = max(lr.get("episodes", 0.0), 1.0)
        avg_len = lr.get("decisions", 0.0) / ep
        avg_life = self.env.warm_up_time + lr.get("simulated_seconds", 0.0) / ep
        return [lg["i_so_far"], lg["t_so_far"], str(round(avg_len, 2)), str(round(avg_life, 2)),
                str(round(float(stats["loss"]), 5)), str(round(mean_reward, 5)), str(round(delta_t, 2))]

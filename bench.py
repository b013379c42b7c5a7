#!/usr/bin/env python
"""bench.py — agent-decisions/s of the batched WRSN hot path (reset / step / observation / reward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--nodes 100] [--chargers 3]
    python bench.py --impl reference ...        # the CPU restatement of the reference on the host cores

One *step* = one pass of the hot path over the batch: ``step`` for every environment (advance to the
next charger decision), the observation raster of every deciding charger, and the reset of every
environment that terminated.  Workload at N = 1: BASELINE.json configs[1] — 100 nodes / 3 chargers / 4096
environments; controller = counter-based uniform 3-vector actions (a ~ U[0,1]^3, a[2] *= 0.05).
Environments shard across ranks by index with no data-path collective (weak scaling: B per GPU is fixed).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "agent-decisions/sec"
UNIT = "decisions/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=60)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    p.add_argument("--nodes", type=int, default=100)
    p.add_argument("--targets", type=int, default=None)
    p.add_argument("--chargers", type=int, default=3)
    p.add_argument("--topologies", type=int, default=64, help="distinct synthetic scenarios per GPU")
    p.add_argument("--scenario", default=None,
                   help="a shipped scenario of the reference, replicated in every environment: name of a committed fixture "
                        "that carries it (e.g. net_hanoi1000n100; the reference's YAML files do not travel to the GPU box)")
    p.add_argument("--threads", type=int, default=0)
    p.add_argument("--preroll", type=int, default=160, help="untimed steps before warm-up that desynchronise the episodes")
    p.add_argument("--groups", type=int, default=8, help="asynchronous environment groups (CUDA streams) per GPU")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--engine-switch", type=int, default=0, help="diagnostic: hdr[OPT_NOBATCH] of every environment (see include/wrsn_b200.h)")
    return p.parse_args()


def workload_name(a):
    if a.scenario:
        return "%s of the reference replicated, %d chargers, %d envs per GPU, uniform 3-vector actions" % (
            a.scenario.replace("net_", ""), a.chargers, a.envs)
    return "%d-node/%d-charger synthetic WRSN, %d envs per GPU, uniform 3-vector actions" % (a.nodes, a.chargers, a.envs)


def scenarios_for(a, rank):
    from multi_agent_rl_wrsn_b200 import Scenario, synthetic
    if a.scenario:                                   # the arrays of the reference's YAML as stored in the golden fixture
        g = np.load(os.path.join(REPO, "tests", "golden", a.scenario + ".npz"), allow_pickle=False)
        q = g["sc_par"]
        spe = dict(capacity=q[0], threshold=q[1], com_range=q[2], sen_range=q[3], prob_gp=q[4], package_size=q[5],
                   er=q[6], et=q[7], efs=q[8], emp=q[9])
        return [Scenario(nodes=g["sc_nodes"].reshape(-1, 2), targets=g["sc_targets"].reshape(-1, 2),
                         base_station=g["sc_bs"], node_phy_spe=spe, max_time=float(q[10]), name=a.scenario)]
    T = a.nodes if a.targets is None else a.targets
    return [synthetic(num_nodes=a.nodes, num_targets=T, seed=1000 + rank * a.topologies + k) for k in range(a.topologies)]


# ----------------------------------------------------------------------------------------------- CPU side
def _cpu_worker(args):
    """One host core: oracle environments (C restatement of the reference) stepped with the same action law."""
    sc_dict, M, seed, budget_s, want_state = args
    from oracle.wrsn_oracle import OracleWRSN, scenario_from_dict
    rng = np.random.default_rng(seed)
    o = OracleWRSN(scenario_from_dict(sc_dict), num_agent=M)
    t0 = time.perf_counter()
    r = o.reset(want_state=want_state)
    n = 1
    ticks = 0.0
    while time.perf_counter() - t0 < budget_s:
        if r["raw_agent_id"] < 0:
            ticks += o.now
            r = o.reset(want_state=want_state)
            n += 1
            continue
        act = rng.uniform(0, 1, 3)
        act[2] *= 0.05
        r = o.step(r["raw_agent_id"], act, want_state=want_state)
        if r["raw_agent_id"] >= 0:
            n += 1
    ticks += o.now
    return n, time.perf_counter() - t0, ticks


def cpu_baseline(a, cores, budget_s):
    import multiprocessing as mp
    scs = scenarios_for(a, 0)
    jobs = [(scs[k % len(scs)].to_dict(), a.chargers, k, budget_s, True) for k in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    per = sum(r[0] / r[1] for r in res)
    return dict(value=per, unit=UNIT, cores=cores, kind="port",
                sample="%d oracle envs (C restatement of rl_env/WRSN.py, one per core), %.0f s each, reset+step+get_state, "
                       "%d decisions, %.0f simulated s" % (cores, budget_s, n, sum(r[2] for r in res)),
                wall_s=wall)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
    vals = []
    for k in range(a.warmup + a.steps):
        cb = cpu_baseline(a, cores, per_step)
        if k >= a.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    ms = 1e3 * float(np.mean([c["wall_s"] for c in vals]))
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=workload_name(a), note="each step = a %.0f s sample on every host core" % per_step),
                impl="reference",
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=vals[-1]["sample"]),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU side
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_b200(a):
    import torch
    import torch.distributed as dist
    from multi_agent_rl_wrsn_b200 import BatchedWRSN
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, M, S, G = a.envs, a.chargers, 100, max(1, a.groups)
    if B % G:
        raise SystemExit("--envs must be a multiple of --groups")
    Bg = B // G
    scs = scenarios_for(a, rank)
    # G asynchronous groups of environments, one CUDA stream each: a group's launch waits only for ITS slowest
    # environment, the other groups keep the SMs busy meanwhile (environments are independent; no collective).
    groups = [BatchedWRSN(scs, num_agent=M, num_envs=Bg, device=dev, threads=a.threads, map_size=S,
                          scenario_index=(np.arange(Bg) + g * Bg) % len(scs)) for g in range(G)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    if a.engine_switch:
        for env in groups:
            env.hdr("OPT_NOBATCH")[:] = float(a.engine_switch)
            env.view("hdr", env._snap)[:, env.E["WRSN_H_OPT_NOBATCH"]] = float(a.engine_switch)
    N, T = groups[0].N, groups[0].T
    obs = [torch.zeros((Bg, 4, S, S), dtype=torch.float32, device=dev) for _ in range(G)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    scale = torch.tensor([1.0, 1.0, 0.05], dtype=torch.float64, device=dev)
    total = a.warmup + a.steps

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident run (value): the actions of every step are already in HBM
    actions = [torch.rand((total, Bg, 3), generator=gen, dtype=torch.float64, device=dev) * scale for _ in range(G)]
    for g in range(G):
        groups[g].reset()
        groups[g].get_state(out=obs[g])
    sync_all()

    def group_step(g, k):
        with torch.cuda.stream(streams[g]):
            groups[g].rollout_step(actions[g][k % total], obs[g])   # 3 launches: step | reset finished episodes | observe

    def totals():
        """(decisions, simulated seconds) so far, from the kernels' own running totals"""
        st = torch.stack([env.req.stats.sum(0) for env in groups]).sum(0)
        return float(st[0].item()), float(st[1].item())

    def resets():
        return float(sum(env.req.stats[:, 2].sum().item() for env in groups))

    # untimed pre-roll: all episodes start at the same instant; run long enough for their phases to spread out, so the
    # timed window sees the steady state of a rollout (resets, death ticks and charging phases mixed) whatever K is
    for k in range(a.preroll):
        for g in range(G):
            group_step(g, k)
    for k in range(a.warmup):
        for g in range(G):
            group_step(g, k)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    dec0, sim0 = totals()
    ep0 = resets()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    for k in range(a.steps):
        for g in range(G):
            group_step(g, a.warmup + k)
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    e1.record(torch.cuda.current_stream(dev))
    sync_all()
    t_wall1 = time.perf_counter()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    dec1, sim1 = totals()
    decisions, ticks = dec1 - dec0, sim1 - sim0
    episodes = resets() - ep0
    n_launch = 3 * G * a.steps

    # ---- end-to-end run (e2e): per step and group, actions come from pinned host memory and the request record is
    # read back on the host before that group's next step is issued
    host_actions = [(torch.rand((a.steps, Bg, 3), dtype=torch.float64) * scale.cpu()).pin_memory() for _ in range(G)]
    host_req = [dict(agent_id=torch.zeros(Bg, dtype=torch.int32).pin_memory(), reward=torch.zeros(Bg, dtype=torch.float64).pin_memory(),
                     terminal=torch.zeros(Bg, dtype=torch.uint8).pin_memory(), now=torch.zeros(Bg, dtype=torch.float64).pin_memory())
                for _ in range(G)]
    dev_action = [torch.zeros((Bg, 3), dtype=torch.float64, device=dev) for _ in range(G)]
    done_ev = [torch.cuda.Event() for _ in range(G)]
    h2d = G * host_actions[0][0].numel() * 8
    d2h = G * sum(v.numel() * v.element_size() for v in host_req[0].values())
    sync_all()
    e2e_dec = 0
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    for k in range(a.steps + 1):
        for g in range(G):
            if k > 0:
                done_ev[g].synchronize()                   # the caller reads step k-1's request of this group
                e2e_dec += int((host_req[g]["agent_id"] >= 0).sum())
            if k == a.steps:
                continue
            with torch.cuda.stream(streams[g]):
                dev_action[g].copy_(host_actions[g][k], non_blocking=True)
                groups[g].rollout_step(dev_action[g], obs[g])
                for name, v in host_req[g].items():
                    v.copy_(getattr(groups[g].req, name), non_blocking=True)
                done_ev[g].record(streams[g])
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    f1.record(torch.cuda.current_stream(dev))
    sync_all()
    e2e_ms = f0.elapsed_time(f1)

    # ---- roofline pass: the same hot path issued serially on ONE stream through the separate entry points, CUDA events
    # around every launch (concurrent groups would time-share the SMs and blur per-launch durations)
    R = min(a.steps, 12)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(R * G)]
    sync_all()
    rd0, rs0 = totals()
    for k in range(R):
        for g in range(G):
            env = groups[g]
            act = actions[g][a.warmup + (k % a.steps)]
            aid = env.req.agent_id
            m = aid >= 0
            e = ev[k * G + g]
            e[0].record()
            env.step(aid, act, mask=m)
            e[1].record()
            env.get_state(out=obs[g])
            e[2].record()
            done = env.req.agent_id < 0
            env.reset(mask=done)
            env.get_state(out=obs[g], agent_id=torch.where(done, env.req.agent_id, torch.full_like(aid, -1)))
    sync_all()
    step_ms = sum(x[0].elapsed_time(x[1]) for x in ev)
    obs_ms = sum(x[1].elapsed_time(x[2]) for x in ev)
    rd1, rs1 = totals()
    r_decisions, r_ticks = rd1 - rd0, rs1 - rs0

    # ---- the density-map decoder, the one streaming kernel of the path: every launch reads another group's
    # maps (G x Bg x S x S float32 = 164 MB at the default sizes, more than the 126 MB L2)
    maps = [torch.rand((Bg, S, S), generator=gen, dtype=torch.float32, device=dev) for _ in range(G)]
    dec_out = torch.zeros((Bg, 3), dtype=torch.float64, device=dev)
    all_agents = torch.zeros(Bg, dtype=torch.int32, device=dev)
    for g in range(G):
        groups[g].density_map_to_action(maps[g], agent_id=all_agents, out=dec_out)
    # (a launch lasts ~10 us, less than the host needs to issue it: the launches are queued behind a spin kernel so that the
    # events bracket back-to-back device execution, not host latency)
    n_dec = 6 * G
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    torch.cuda._sleep(int(2e7))                        # ~10 ms at 1.9 GHz
    d0.record()
    for k in range(n_dec):
        groups[k % G].density_map_to_action(maps[k % G], agent_id=all_agents, out=dec_out)
    d1.record()
    sync_all()
    dec_ms = d0.elapsed_time(d1) / n_dec
    dec_bytes = Bg * (S * S * 4 + 24)
    del maps

    # ---- the dense charging kernel (wrsn_k_charge): not launched by the rollout (the step kernel applies the same
    # model incrementally), timed alone like the decoder; a failure here only drops the figure
    chg = None
    try:
        import ctypes as C
        from multi_agent_rl_wrsn_b200 import _lib
        node_rate = torch.zeros((Bg, N), dtype=torch.float64, device=dev)
        mc_rate = torch.zeros((Bg, M), dtype=torch.float64, device=dev)

        def charge_launch(env):
            _lib.check(env.L.wrsn_k_charge(C.byref(env.dims), env.scen.data_ptr(), env.scen_id.data_ptr(), env.state.data_ptr(),
                                           None, node_rate.data_ptr(), mc_rate.data_ptr(), env._stream()), env.L)

        for g in range(G):
            charge_launch(groups[g])
        n_chg = 6 * G
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        torch.cuda._sleep(int(2e7))
        c0.record()
        for k in range(n_chg):
            charge_launch(groups[k % G])
        c1.record()
        sync_all()
        chg_ms = c0.elapsed_time(c1) / n_chg
        chg_bytes = Bg * (17 * N + 8 * N + 8 * M)
        chg = dict(ms_per_launch=chg_ms, algorithmic_bytes=chg_bytes, gbs=chg_bytes / (chg_ms * 1e-3) / 1e9, share=0.0,
                   note="dense node x charger charging model, standalone (unit parity / profiling); not launched by the "
                        "rollout, which applies the model incrementally inside the step kernel")
        del node_rate, mc_rate
    except Exception as ex:                              # noqa: BLE001 - diagnostic figure only
        chg = dict(error=str(ex)[:200])

    # ---- the reference's RandomController (controller/random/RandomController.py:12: map = s0 + s1 - 10 s2 + s3) as the
    # action source: torch forms the map from the observation in HBM, the decoder turns it into the action, rollout_step
    # consumes it.  A short device-timed run, reported beside the headline (which feeds 3-vector actions).
    K2 = min(a.steps, 30)
    act_buf = [torch.zeros((Bg, 3), dtype=torch.float64, device=dev) for _ in range(G)]

    def map_step(g):
        with torch.cuda.stream(streams[g]):
            o = obs[g]
            dm = o[:, 0] + o[:, 1] - 10.0 * o[:, 2] + o[:, 3]
            groups[g].density_map_to_action(dm, out=act_buf[g])
            groups[g].rollout_step(act_buf[g], obs[g])

    for k in range(3):
        for g in range(G):
            map_step(g)
    sync_all()
    md0, ms0 = totals()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    for k in range(K2):
        for g in range(G):
            map_step(g)
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    m1.record(torch.cuda.current_stream(dev))
    sync_all()
    md1, ms1 = totals()
    map_ms, map_dec, map_sim = m0.elapsed_time(m1), md1 - md0, ms1 - ms0

    # ---- the rollout loop of the reference's IPPO trainer (controller/ippo/IPPO.py:128-155) with its transition record
    # kept in HBM (controllers.IPPORollout): a Gaussian policy around the RandomController map stands in
    # for the actors (map, sample, log-probability by torch), the decoder and rollout_step do the rest.  Windows of T3 steps.
    from multi_agent_rl_wrsn_b200.controllers import IPPORollout
    T3, W3 = 6, 3
    sigma = 1e-3

    def gauss_policy(agent_id, o):
        mean = o[:, 0] + o[:, 1] - 10.0 * o[:, 2] + o[:, 3]
        x = torch.randn_like(mean).mul_(sigma).add_(mean)
        lp = (-0.5 * ((x - mean) / sigma) ** 2).sum((1, 2)) - S * S * (math.log(sigma) + 0.5 * math.log(2.0 * math.pi))
        return x, lp

    ros = [IPPORollout(groups[g], T3) for g in range(G)]
    sync_all()                                   # the records were initialised on the current stream; the groups' streams follow

    def ippo_window():
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                ros[g].carry_over()
                ros[g].collect(gauss_policy)

    ippo_window()
    sync_all()
    id0, is0 = totals()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    for k in range(W3):
        ippo_window()
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    i1.record(torch.cuda.current_stream(dev))
    sync_all()
    id1, is1 = totals()
    ippo_ms, ippo_dec, ippo_sim = i0.elapsed_time(i1), id1 - id0, is1 - is0
    ippo_tr = float(sum(sum(int(r.transitions(i)[0].numel()) for i in range(M)) for r in ros))   # of the last window
    del ros

    # ---- reduce over ranks: max time, summed work
    t = torch.tensor([elapsed_ms, e2e_ms, map_ms, ippo_ms], dtype=torch.float64, device=dev)
    w = torch.tensor([decisions, ticks, float(e2e_dec), map_dec, episodes, ippo_dec, ippo_tr, map_sim, ippo_sim], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, map_ms, ippo_ms = [float(x) for x in t.tolist()]
    decisions_all, ticks_all, e2e_dec_all, map_dec_all, episodes_all, ippo_dec_all, ippo_tr_all, map_sim_all, ippo_sim_all = [float(x) for x in w.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = decisions_all / (elapsed_ms * 1e-3)
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # algorithmic bytes per SURVEY §8(d): 82 N + T per environment and simulated second, 16 T + 88 per decision in the
    # step kernel, 4 S^2 float32 per decision in the observation kernel
    n_l = R * G
    step_bytes = (r_ticks * (82 * N + T) + r_decisions * (16 * T + 88)) / n_l
    obs_bytes = r_decisions * (4 * S * S * 4) / n_l
    step_gbs = step_bytes / (step_ms / n_l * 1e-3) / 1e9
    obs_gbs = obs_bytes / (obs_ms / n_l * 1e-3) / 1e9
    traffic = None          # DRAM bytes per launch of the dominant kernel from the committed ncu capture, same launch shape only
    try:
        with open(os.path.join(REPO, "profiles", "traffic_r01.json")) as f:
            tj = json.load(f)
        if tj.get("nodes") == N and tj.get("chargers") == M and tj.get("envs_per_launch") == Bg:
            traffic = tj
    except Exception:
        pass
    dominant = "k_env<MODE_STEP>" if step_ms >= obs_ms else "k_observe<float>"
    ach = step_gbs if step_ms >= obs_ms else obs_gbs
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
        ms_per_step=elapsed_ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
        data="synthetic",
        config=dict(workload=workload_name(a), nodes=N, targets=T, chargers=M, envs_per_gpu=B, map_size=S,
                    groups="%d asynchronous groups of %d environments, one CUDA stream each" % (G, Bg),
                    topologies_per_gpu=a.topologies, threads_per_env=int(groups[0].dims.threads),
                    observation="float32 [B,4,100,100] kept in HBM",
                    l2="working set (state %.0f MB + observations %.0f MB per GPU) exceeds the 126 MB L2; no explicit flush"
                       % (B * groups[0].dims.state_bytes / 1e6, B * 4 * S * S * 4 / 1e6),
                    sim_seconds_per_decision=ticks_all / max(decisions_all, 1.0),
                    env_ticks_per_s=ticks_all / (elapsed_ms * 1e-3),
                    episodes_per_s=episodes_all / (elapsed_ms * 1e-3),
                    resets="finished episodes are reset inside the timed step (wrsn_rollout_step: step | reset | observe)",
                    random_controller_map=dict(value=map_dec_all / (map_ms * 1e-3), unit=UNIT, steps=K2,
                                               sim_seconds_per_decision=map_sim_all / max(map_dec_all, 1.0),
                                               note="same environments driven by the reference's RandomController density map "
                                                    "(s0 + s1 - 10 s2 + s3), decoded on the device by wrsn_decode_density_map; "
                                                    "device-resident, not the headline workload"),
                    ippo_rollout_record=dict(value=ippo_dec_all / (ippo_ms * 1e-3), unit=UNIT, steps=T3 * W3,
                                             sim_seconds_per_decision=ippo_sim_all / max(ippo_dec_all, 1.0),
                                             transitions_last_window=ippo_tr_all,
                                             note="IPPO.roll_out's loop (IPPO.py:128-155) with the per-agent transition record "
                                                  "(state, map action, log-probability, reward, next state) kept in HBM by "
                                                  "controllers.IPPORollout; Gaussian policy around the RandomController map as "
                                                  "the stand-in actor, maps decoded on the device; windows of %d steps" % T3)),
        clocks=clocks,
        e2e=dict(value=e2e_dec_all / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 note="per step: actions from pinned host memory, rollout_step (step | reset | observe), request record "
                      "(agent_id, reward, terminal, now) read back on the host; observations stay in HBM for the policy networks"),
        gpu_launches=n_launch,
        roofline=dict(bound="hbm", kernel=dominant, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                      traffic=(traffic or {}).get(dominant), traffic_source=(traffic or {}).get("source"),
                      peak_source=peak_src,
                      how="serialized pass of %d launches per kernel on one stream right after the timed region" % n_l,
                      kernels={"k_env<MODE_STEP>": dict(ms_per_launch=step_ms / n_l, algorithmic_bytes=step_bytes, gbs=step_gbs,
                                                        share=step_ms / (step_ms + obs_ms)),
                               "k_observe<float>": dict(ms_per_launch=obs_ms / n_l, algorithmic_bytes=obs_bytes, gbs=obs_gbs,
                                                        share=obs_ms / (step_ms + obs_ms)),
                               "k_decode_map<float>": dict(ms_per_launch=dec_ms, algorithmic_bytes=dec_bytes,
                                                           gbs=dec_bytes / (dec_ms * 1e-3) / 1e9,
                                                           frac=dec_bytes / (dec_ms * 1e-3) / 1e9 / peak,
                                                           share=0.0, note="not launched by the headline workload (3-vector actions); "
                                                                           "timed alone on maps larger than L2"),
                               "k_charge": chg}),
    )
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(a, 1, a.cpu_seconds).items() if k != "wall_s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

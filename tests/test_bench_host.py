"""Host-side pieces of bench.py that run without a GPU: the workloads it builds and the reference arm's CPU sample."""
import argparse
import importlib.util
import os

from tests.helpers import REPO


def _bench():
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    base = dict(nodes=100, targets=None, chargers=3, envs=4096, topologies=3, scenario=None, actions="controller")
    base.update(kw)
    return argparse.Namespace(**base)


def test_workloads():
    b = _bench()
    scs = b.scenarios_for(_args(), rank=1)
    assert len(scs) == 3 and all(s.N == 100 and s.T == 100 for s in scs)
    other = b.scenarios_for(_args(), rank=0)
    assert any((x.nodes != y.nodes).any() for x, y in zip(scs, other))          # every rank has its own topologies
    shipped = b.scenarios_for(_args(scenario="net_hanoi1000n100"), rank=0)      # the reference's hanoi1000n100.yaml
    assert len(shipped) == 1 and (shipped[0].N, shipped[0].T) == (98, 100)
    assert "hanoi1000n100" in b.workload_name(_args(scenario="net_hanoi1000n100"))
    assert "100-node/3-charger" in b.workload_name(_args())


def test_cpu_sample_of_the_reference_arm():
    """One short single-core sample of the C restatement, as `cpu_baseline` / `--impl reference` take it."""
    b = _bench()
    out = b.cpu_baseline(_args(topologies=1, actions="uniform"), cores=1, budget_s=1.0)
    assert out["kind"] == "port" and out["cores"] == 1 and out["unit"] == b.UNIT and out["value"] > 10.0
    # the headline law: the RandomController map decoded by the reference's own numpy / scipy statements (slow: L-BFGS-B)
    out = b.cpu_baseline(_args(topologies=1), cores=1, budget_s=2.0)
    assert out["value"] > 0.5 and "RandomController" in out["sample"]
    assert "RandomController" in b.workload_name(_args()) and "uniform" in b.workload_name(_args(actions="uniform"))

"""B200-native batched WRSN simulator — drop-in for the hot path of ``rl_env/WRSN.py``.

``BatchedWRSN`` advances thousands of independent network + mobile-charger instances per launch;
``WRSN`` is the single-environment façade with the reference's constructor and request dict.
"""
from .scenario import Scenario, build_static, load_mc_type, synthetic  # noqa: F401
from .batched import BatchedWRSN, Requests  # noqa: F401
from .wrsn import WRSN  # noqa: F401
from .ippo import BatchedIPPO  # noqa: F401
from .controllers import (BatchedRandomController, IPPORollout, PerAgentPolicy, allreduce_gradients, ppo_update,  # noqa: F401
                          rollout, select_batch)  # noqa: F401

__all__ = ["Scenario", "build_static", "load_mc_type", "synthetic", "BatchedWRSN", "Requests", "WRSN", "BatchedIPPO",
           "BatchedRandomController", "IPPORollout", "PerAgentPolicy", "allreduce_gradients", "ppo_update", "rollout", "select_batch"]

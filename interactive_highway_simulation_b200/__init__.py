"""B200-native rollout path of einsteinguang/interactive_highway_simulation.

    backend          ctypes layer over libmcsim_b200.so (C ABI: include/mcsim.h) — DeviceScene, rollouts, decisions
    mc_sim_highway   drop-in for the reference's precompiled boost-python module of the same name
    synthetic        synthetic scenes in the shape mcs_wrapper.py hands to the inner simulator
    sharding         rollout-range partition over ranks and the fixed-order feature exchange (torch.distributed)

There is no CPU implementation in this package: without the CUDA library and a device every compute call raises.
"""
import sys as _sys

__version__ = "0.1.0"


def install_as_mc_sim_highway():
    """Register the drop-in under the reference's module name, so that the reference's own Python
    (`from mc_sim_highway import ...` at behaviors.py:5-7, mcs_wrapper.py:3-5, type.py:4-6) imports it unchanged."""
    from . import mc_sim_highway as m
    _sys.modules["mc_sim_highway"] = m
    return m

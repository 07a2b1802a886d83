#!/usr/bin/env python
"""Profile driver for the outer-environment kernels: bench.py's outer_env workload (512 scenes x 200 vehicles), a few
mce_env_match calls.   ncu --set full -k regex:env_ -c 4 python scripts/env_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from interactive_highway_simulation_b200 import outer_env  # noqa: E402

corridors, x, y = bench.outer_env_workload()
dm = outer_env.DeviceMap(corridors, 0)
for _ in range(int(os.environ.get("REPS", "3"))):
    t = dm.match(x, y, x.shape[0])
print("kernel ms", dm.last_kernel_ms(), "matched", float((t.lane >= 0).mean()))

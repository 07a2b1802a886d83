#!/usr/bin/env python
"""CPU baseline of BASELINE.json configs[0] (build container only: needs /root/reference and oracle/_ref).

Times THE REFERENCE'S OWN PYTHON outer loop (simulation_multilane / environment / agent / behaviors / mcs_wrapper, imported
unchanged from /root/reference) headless on its shipped test_scene — dt 0.2, 100 steps, 21 vehicles — with
  * tests/refshim/shapely  standing in for shapely/GEOS (a pure-Python restatement: slower than GEOS), and
  * tests/refshim/ref_mc_sim_highway.py, i.e. the reference's machine code behind a pipe, as `mc_sim_highway`.
The reference cannot travel to the GPU box, so the figure is recorded here and bench.py reports it next to the product's
outer loop with exactly this label:  python scripts/time_reference_outer_loop.py  ->  profiles/r02_reference_outer_loop_cpu.json
"""
import json
import os
import platform
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TESTS = os.path.join(ROOT, "tests")
REF = "/root/reference"
sys.path[:0] = [os.path.join(TESTS, "refshim"), ROOT, TESTS, REF]

import numpy as np  # noqa: E402
import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402


def once(steps, dt=0.2):
    random.seed(0); np.random.seed(0); backend.seed(1000)
    backend.RECORD.clear()
    sim = sm.Simulation(None, sm.SimulationParameter(dt, steps, render=False, write_gif=False))
    sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
    n0 = len(sim.environment.agents)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.environment.step(dt)
    return time.perf_counter() - t0, n0, len(backend.RECORD)


def main():
    steps = 100
    once(5)
    best, n0, calls = min(once(steps) for _ in range(3))
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"workload": "test_scene epoch0, %d vehicles, dt 0.2, %d steps, render=False" % (n0, steps),
           "wall_s": best, "ms_per_step": best / steps * 1e3, "inner_calls": calls, "cores": 1, "kind": "reference",
           "host": "%s (%d vCPU), %s" % (cpu, os.cpu_count(), platform.python_version()),
           "what": "the reference's own Python outer loop, unchanged, over tests/refshim (pure-Python shapely stand-in) and the "
                   "reference's machine code behind a pipe; recorded in the build container, NOT on the GPU box"}
    path = os.path.join(ROOT, "profiles", "r02_reference_outer_loop_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

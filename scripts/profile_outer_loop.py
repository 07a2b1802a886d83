"""Host profile (cProfile) of configs[0] free-running on the device: where a step of the sequential outer loop spends its time.
usage: python scripts/profile_outer_loop.py [steps]   (diagnostic; needs a GPU)"""
import cProfile
import os
import pstats
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interactive_highway_simulation_b200 import mc_sim_highway as ms, simulation_multilane as sm  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
scene = os.path.join(ROOT, "tests", "golden", "test_scene_epoch0")
for rep in range(2):
    random.seed(0); np.random.seed(0); ms.seed(1000)
    sim = sm.Simulation(None, sm.SimulationParameter(0.2, steps, render=False, write_gif=False))
    sim.load_initial_scene(scene)
    if rep == 1:
        pr = cProfile.Profile()
        pr.enable()
    sim.run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

"""Per-phase cycle counts of the device walk (library built with -DMCW_TIMING; diagnostic).
usage: walk_timing.py path/to/lib_with_MCW_TIMING.so [steps]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interactive_highway_simulation_b200 import backend
backend.LIB_PATH = os.path.abspath(sys.argv[1])
from interactive_highway_simulation_b200 import mc_sim_highway as ms, simulation_multilane as sm  # noqa: E402
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
random.seed(0); np.random.seed(0); ms.seed(1000)
sim = sm.Simulation(None, sm.SimulationParameter(0.2, steps, render=False, write_gif=False))
sim.load_initial_scene(os.path.join(ROOT, "tests", "golden", "test_scene_epoch0"))
sim.run()

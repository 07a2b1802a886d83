#!/usr/bin/env python
"""Writes tests/golden/test_scene_epoch0/{agents,map}.yml: the reference's shipped scene (configs[0]) loaded by the
reference's own simulation_multilane.Simulation.load_initial_scene and written back by its own dump format
(initial_agents_dict = Agent.to_dict, Map.to_dict; simulation_multilane.py:117-135) — the scene the outer-loop tests load
on the GPU box, where /root/reference does not exist.

    python tests/golden/make_test_scene_fixture.py          (build container only: needs /root/reference)
"""
import os
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(TESTS, "refshim"), os.path.dirname(TESTS), TESTS, REF]

import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402

sim = sm.Simulation(None, sm.SimulationParameter(0.2, 1, render=False, write_gif=False))
sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
out = os.path.join(HERE, "test_scene_epoch0")
os.makedirs(out, exist_ok=True)
for name, data in (("agents", sim.initial_agents_dict), ("map", sim.environment.map.to_dict())):
    with open(os.path.join(out, name + ".yml"), "w") as f:
        yaml.dump(data, f, default_flow_style=False)
print("->", out, {i: a["behavior_model"] for i, a in sim.initial_agents_dict.items()})

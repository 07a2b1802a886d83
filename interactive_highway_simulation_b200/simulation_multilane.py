"""Headless mirror of the reference's outer simulation driver (simulation_multilane.py:9-244): SimulationParameter,
Simulation (reset, load_initial_scene, dump_initial_scene, if_terminate, run) and the yaml scene format
(`agents.yml` / `map.yml`, simulation_multilane.py:117-154).  Rendering, gif writing and matplotlib state are out of
scope: `render` / `write_gif` are accepted and ignored.  Every step is `outer_env.Environment.step`: device
lane match / neighbour tables, behaviours answered by the rollout and decision kernels, device step_along_path.
"""
import datetime
import os

import yaml

from .agent import create_agent_from_dict
from .outer_env import Environment, Map, create_corridor_from_dict


class SimulationParameter:                                   # simulation_multilane.py:9-16
    def __init__(self, dt, total_steps, render=True, write_gif=True, show_debug_info=False, scenario="lane_change"):
        self.dt, self.total_steps, self.render, self.write_gif = dt, total_steps, render, write_gif
        self.show_debug_info, self.scenario = show_debug_info, scenario


# yaml.safe_load's scanner and parser in C where PyYAML was built with libyaml (the constructors — what turns the scalars
# into Python values — are the same Python code either way): loading an epoch's two files is a third of a learned-policy
# evaluation with the pure-Python scanner (0.05 of 0.19 s)
_SAFE_LOADER = getattr(yaml, "CSafeLoader", yaml.SafeLoader)


def load_yaml(path):                                         # simulation_multilane.py:19-22
    with open(path) as infile:
        return yaml.load(infile, Loader=_SAFE_LOADER)


class Simulation:
    def __init__(self, env, param, device=0):                # simulation_multilane.py:26-42
        self.device = device
        self.write_dir = os.path.join(os.getcwd(), datetime.datetime.now().strftime('%Y-%m-%d_%H-%M-%S'))
        self.epoch = 0
        self.ego_id = -1
        self.reset(env, param)

    def reset(self, env, param):                             # simulation_multilane.py:44-53
        self.environment = env
        self.initial_agents_dict = {i: a.to_dict() for i, a in env.agents.items()} if env is not None else {}
        self.param = param
        self.step = 0

    def if_terminate(self):                                  # simulation_multilane.py:105-113
        if self.param.scenario == "exit" and self.environment.all_finish_exit():
            return True
        if self.param.scenario == "merging" and self.environment.all_finish_merging():
            return True
        return False

    def load_initial_scene(self, path):                      # simulation_multilane.py:115-131
        self.step = 0
        agents = [create_agent_from_dict(a) for a in load_yaml(os.path.join(path, "agents.yml")).values()]
        corridors = [create_corridor_from_dict(c) for c in load_yaml(os.path.join(path, "map.yml")).values()]
        self.environment = Environment(Map(corridors, self.device), agents)
        self.initial_agents_dict = {i: a.to_dict() for i, a in self.environment.agents.items()}

    def dump_initial_scene(self):                            # simulation_multilane.py:133-135
        self.to_yaml(self.initial_agents_dict, "agents")
        self.to_yaml(self.environment.map.to_dict(), "map")

    def dump_print_outputs(self, output):                    # simulation_multilane.py:137-141
        os.makedirs(self.write_dir, exist_ok=True)
        with open(os.path.join(self.write_dir, 'console_output.txt'), 'w') as f:
            f.write(output)

    def to_yaml(self, data, data_name):                      # simulation_multilane.py:143-152
        data_dir = os.path.join(self.write_dir, "epoch" + str(self.epoch))
        os.makedirs(data_dir, exist_ok=True)
        path = os.path.join(data_dir, data_name + '.yml')
        with open(path, 'w') as outfile:
            yaml.dump(data, outfile, default_flow_style=False)
        return path

    def run(self, write_fig_dir=None):                       # simulation_multilane.py:172-222 without the drawing
        terminate_counter = 0
        for step in range(self.param.total_steps):
            self.environment.step(self.param.dt)
            self.step = step
            if self.if_terminate():                          # one more second of simulation after termination
                if terminate_counter * self.param.dt > 0.4:
                    break
                terminate_counter += 1

"""The reference's three evaluation scripts (evaluation_merging.py, evaluation_lane_change.py, evaluation_exit.py) on the
package's headless outer loop: scene generation into the reference's on-disk format (`<dir>/epoch<N>/{agents,map}.yml`),
the choice of ego vehicles, and `evaluate_recorded_*_scene` with the statistics they accumulate.  Same function names,
arguments, return strings and statistics side effects as the reference; plotting is out of scope.
"""
import os

import numpy as np

from . import traffic_generator as tg
from .agent import LaneChangeStatistics, MergingAndExitStatistics, create_agents_from_dicts  # noqa: F401  (re-exported)
from .outer_env import Environment, Map, create_corridor_from_dict
from .simulation_multilane import Simulation, SimulationParameter  # noqa: F401

SCENARIOS = {   # kind: (v_limit candidates, SimulationParameter.scenario)   evaluation_{merging,lane_change,exit}.py
    "merging": ([80 / 3.6, 60 / 3.6, 100 / 3.6], "merging"),
    "lane_change": ([100 / 3.6], "lane_change"),
    "exit": ([100 / 3.6], "exit"),
}


def random_environment(kind, device=0, generator=None):
    """One epoch of generate_random_{merging,lc,exit}_traffic (evaluation_merging.py:17-46, evaluation_lane_change.py:18-51,
    evaluation_exit.py:18-49): draws v_limit, builds the map, generates the traffic, passes it through the yaml dict format.
    Consumes numpy's and `random`'s global streams in the reference's order."""
    v_limit = float(np.random.choice(SCENARIOS[kind][0], 1)[0])
    spec, params, unsafe, truck = tg.evaluation_map(kind, v_limit=v_limit)
    gen = generator or tg.RandomTrafficGenerator()
    merging, main = gen.random_agents_on_map(tg.lanes_from_spec(spec), {i: tg.RandomDensityAndVelocityParameter(*p) for i, p in params.items()},
                                             unsafe, truck)
    if kind != "merging":      # evaluation_lane_change.py:56, evaluation_exit.py:52 restart the ids every epoch; evaluation_merging.py
        gen.vehicle_id = 0     # never does: its epochs after the first carry cumulative vehicle ids
    corridors = [create_corridor_from_dict({**c, "left_id": "none" if c["left_id"] is None else c["left_id"],
                                            "right_id": "none" if c["right_id"] is None else c["right_id"]}) for c in spec]
    return Environment(Map(corridors, device), create_agents_from_dicts(merging) + create_agents_from_dicts(main))


def generate_random_traffic(kind, n_epochs, sim):
    """The generation loops of the three scripts: n_epochs scenes dumped under sim.write_dir/epoch<k>/ (agents.yml, map.yml)."""
    gen = tg.RandomTrafficGenerator()
    for _ in range(n_epochs):
        sim.reset(random_environment(kind, sim.device, gen), sim.param)
        sim.dump_initial_scene()
        sim.epoch += 1


def ego_candidates(env, kind):
    """Which vehicles the scripts evaluate as ego: every merging vehicle (evaluation_merging.py:56-62); the 3rd..5th vehicle
    of every main lane (evaluation_lane_change.py:118-123); the 3rd..4th of lanes 2 and 3 (evaluation_exit.py:123-128);
    never a truck."""
    if kind == "merging":
        return [a.id for a in env.agents.values() if "merging" in a.behavior_model]
    out = []
    for c_id, sorted_agents in env.matched_agents_sorted_by_arc_length.items():
        if kind == "lane_change" and env.map.corridors[c_id].type == "main":
            out += [a.id for _, a in sorted_agents[2:5] if a.l <= 7]
        if kind == "exit" and c_id in (2, 3):
            out += [a.id for _, a in sorted_agents[2:4] if a.l <= 7]
    return out


def _invalid(statistics, with_t):
    for lst in (statistics.utility_ego, statistics.comfort_ego, statistics.utility_other, statistics.comfort_other):
        lst.append(-10)
    if with_t:
        statistics.total_t.append(-10)
    statistics.epoch_and_ego_id_history.append([-1, -1])
    return statistics.__repr__() + "\n"


def _ego_and_others(sim, ego_id, statistics):
    env = sim.environment
    agent = env.agents[ego_id] if ego_id in env.agents else env.removed_agents[ego_id]
    statistics.utility_ego.append(agent.average_utility())
    statistics.comfort_ego.append(agent.minimum_comfort())
    utility_others, comfort_others = 1e10, 1e10
    for a in list(env.agents.values()) + list(env.removed_agents.values()):
        if a.id != ego_id:
            comfort_others = min(a.minimum_comfort(), comfort_others)
            utility_others = min(a.average_utility(), utility_others)
    statistics.utility_other.append(utility_others)
    statistics.comfort_other.append(comfort_others)
    return agent


def _known(sim, ego_id):
    return ego_id in sim.environment.agents or ego_id in sim.environment.removed_agents


def evaluate_recorded_merging_scene(path, sim, ego_policy, epoch_id, statistics, max_fall_back_ratio=1.0):   # evaluation_merging.py:53-106
    sim.reset(None, sim.param)
    sim.load_initial_scene(path)
    ego_ids = []
    for a in sim.environment.agents.values():
        if "merging" in a.behavior_model:
            ego_ids.append(a.id)
            a.reset_behavior_model(ego_policy)
            if "learn" in ego_policy:
                a.learned_merging_model.learned_model_parameter.max_fall_back_rate = max_fall_back_ratio
    sim.run(os.path.join(path, "fig"))
    print_output = ego_policy + " : " + "\n"
    for ego_id in ego_ids:
        if not _known(sim, ego_id):
            return print_output + _invalid(statistics, True)
        env = sim.environment
        agent = env.agents[ego_id] if ego_id in env.agents else env.removed_agents[ego_id]
        statistics.num_total_agents += 1
        t = agent.merging_flag.t_finish_merging
        statistics.total_t.append(t if t is not None else sim.param.dt * sim.param.total_steps)
        statistics.num_finish += agent.merging_flag.finish_merging
        statistics.num_fallback += agent.merging_flag.once_fallback
        statistics.epoch_and_ego_id_history.append([epoch_id, ego_id])
        _ego_and_others(sim, ego_id, statistics)
        print_output += statistics.__repr__() + "\n"
    return print_output


def evaluate_recorded_lc_scene(path, sim, ego_policy, epoch_id, ego_id, statistics):                       # evaluation_lane_change.py:58-101
    sim.reset(None, sim.param)
    sim.load_initial_scene(path)
    for a in sim.environment.agents.values():
        if a.id == ego_id:
            a.reset_behavior_model(ego_policy)
    sim.run(os.path.join(path, "fig"))
    print_output = ego_policy + " with id {}: ".format(ego_id)
    if not _known(sim, ego_id):
        return print_output + _invalid(statistics, False)
    env = sim.environment
    agent = env.agents[ego_id] if ego_id in env.agents else env.removed_agents[ego_id]
    if agent.had_harsh_brake():
        statistics.num_fallback += 1
    statistics.num_lane_change += agent.num_lane_change()
    statistics.lane_id_history.append(agent.corridor_history[-1])
    statistics.epoch_and_ego_id_history.append([epoch_id, ego_id])
    _ego_and_others(sim, ego_id, statistics)
    return print_output + statistics.__repr__() + "\n"


def evaluate_recorded_exit_scene(path, sim, ego_policy, epoch_id, ego_id, statistics, max_fall_back_ratio=1.0):   # evaluation_exit.py:55-105
    sim.reset(None, sim.param)
    sim.load_initial_scene(path)
    for a in sim.environment.agents.values():
        if a.id == ego_id:
            a.reset_behavior_model(ego_policy)
            if "learn" in ego_policy:
                a.learned_merging_model.learned_model_parameter.max_fall_back_rate = max_fall_back_ratio
    sim.run(os.path.join(path, "fig"))
    print_output = ego_policy + " with id {}: ".format(ego_id)
    if not _known(sim, ego_id):
        return print_output + _invalid(statistics, True)
    env = sim.environment
    agent = env.agents[ego_id] if ego_id in env.agents else env.removed_agents[ego_id]
    statistics.num_total_agents += 1
    t = agent.exit_flag.t_finish_exit
    statistics.total_t.append(t if t is not None else sim.param.dt * sim.param.total_steps)
    statistics.num_finish += agent.exit_flag.finish_exit
    statistics.num_fallback += agent.exit_flag.once_fallback
    statistics.epoch_and_ego_id_history.append([epoch_id, ego_id])
    _ego_and_others(sim, ego_id, statistics)
    return print_output + statistics.__repr__() + "\n"

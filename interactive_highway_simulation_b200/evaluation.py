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


# ---- many evaluations side by side ------------------------------------------------------------------------------------
# The scripts' main loops (evaluation_merging.py:108-140, evaluation_lane_change.py:103-140, evaluation_exit.py:107-150) call
# evaluate_recorded_*_scene once per (epoch, ego, policy), 500 epochs one after the other.  An evaluation is a sequential outer
# loop of ~100 steps — a chain of small, latency-bound launches and the host bookkeeping of ~25 vehicles; the evaluations
# themselves are independent, so they shard: evaluate_scenes_parallel runs them in worker processes, task i on device
# devices[i % len(devices)], and merges the statistics in task order.  Each task seeds the three random streams itself
# (numpy's, Python's, the drop-in's), so a task's result does not depend on which worker or GPU runs it or on what ran
# before it — unlike the scripts' single run, where evaluation k starts from the stream positions evaluation k-1 left
# behind.  Measured (scripts/eval_parallel.py): the shard unit is the GPU, not the host core — worker processes that share
# ONE GPU do not overlap (their contexts time-slice the device, and an evaluation keeps it occupied ~2/3 of the time:
# 5.5 evaluations/s with 1 worker, 5.8 with 15), workers on different GPUs scale with the number of GPUs.

def _new_statistics(kind):
    return LaneChangeStatistics() if kind == "lane_change" else MergingAndExitStatistics()


def _evaluate_task(task):
    """One (epoch, ego, policy) evaluation in a worker: (kind, path, policy, epoch_id, ego_id, max_fall_back_ratio, seed, dt, steps,
    device) -> (console text, the statistics object's fields)."""
    import random
    from . import mc_sim_highway as ms
    kind, path, policy, epoch_id, ego_id, ratio, seed, dt, steps, device = task
    if _WORKER_DEVICE is not None:
        device = _WORKER_DEVICE                       # a pool worker stays on the GPU it was given
    random.seed(seed); np.random.seed(seed); ms.seed(seed)
    sim = Simulation(None, SimulationParameter(dt, steps, render=False, write_gif=False, scenario=kind))
    sim.device = device
    st = _new_statistics(kind)
    if kind == "merging":
        text = evaluate_recorded_merging_scene(path, sim, policy, epoch_id, st, ratio)
    elif kind == "lane_change":
        text = evaluate_recorded_lc_scene(path, sim, policy, epoch_id, ego_id, st)
    else:
        text = evaluate_recorded_exit_scene(path, sim, policy, epoch_id, ego_id, st, ratio)
    return text, dict(st.__dict__)


def merge_statistics(kind, parts):
    """The statistics object a single run over the same tasks in the same order would have built: lists concatenated,
    counters summed."""
    total = _new_statistics(kind)
    for part in parts:
        for name, value in part.items():
            if isinstance(value, list):
                getattr(total, name).extend(value)
            elif isinstance(value, (int, float)) and not isinstance(value, bool):
                setattr(total, name, getattr(total, name) + value)
    return total


_WORKER_DEVICE = None


def _bind_worker(counter, devices):
    global _WORKER_DEVICE
    with counter.get_lock():
        k = counter.value
        counter.value += 1
    _WORKER_DEVICE = devices[k % len(devices)]


class EvaluationPool:
    """Worker processes for evaluate_scenes_parallel, spawned once: a worker keeps its CUDA context and the loaded library,
    and is bound to one GPU (worker k to devices[k % len(devices)])."""

    def __init__(self, workers, devices=(0,)):
        import multiprocessing
        ctx = multiprocessing.get_context("spawn")
        self.workers = workers
        self._pool = ctx.Pool(workers, initializer=_bind_worker, initargs=(ctx.Value("i", 0), list(devices)))

    def run(self, tasks):
        return self._pool.map(_evaluate_task, tasks, chunksize=1)

    def close(self):
        self._pool.close()
        self._pool.join()


def evaluate_scenes_parallel(kind, tasks, workers=None, pool=None, dt=0.2, steps=100, devices=(0,)):
    """tasks: [(path, policy, epoch_id, ego_id, max_fall_back_ratio, seed)] -> ([console texts], merged statistics).
    devices: the GPUs of the pool's workers (default worker count: one per device); a pool passed in keeps its own."""
    devices = list(devices)
    full = [(kind, path, policy, epoch_id, ego_id, ratio, seed, dt, steps, devices[0])
            for path, policy, epoch_id, ego_id, ratio, seed in tasks]
    own = pool is None
    if own:
        pool = EvaluationPool(workers or min(len(full), len(devices)), devices)
    try:
        results = pool.run(full)
    finally:
        if own:
            pool.close()
    return [r[0] for r in results], merge_statistics(kind, [r[1] for r in results])

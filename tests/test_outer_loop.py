"""configs[0] free-running: the product's outer loop (simulation_multilane.Simulation -> outer_env.Environment.step ->
agent.plan / step -> behaviors) loads the shipped scene from yaml and steps it; after every step every vehicle must be
where the reference's own Python had it (tests/golden/scenario_*.json.gz `trace`, written by
tests/golden/make_scenario_golden.py), and the vehicles leave the map at the same steps.

CPU: host logic only (geometry by the oracle, inner-simulator answers from the records — see outer_loop_cases.py).
GPU: the real thing — device geometry kernels, MOBIL / gap decisions and rollouts from the CUDA kernels."""
import pytest

from outer_loop_cases import (SCENE_DIR, OracleDeviceMap, RecordedInnerSimulator, compare_trace, load_scenario, seed_streams)


def run_test_scene(sim_mod, rec, overrides=None, seed=0, check=None):
    seed_streams(seed)
    sim = sim_mod.Simulation(None, sim_mod.SimulationParameter(rec["dt"], rec["steps"], render=False, write_gif=False))
    sim.load_initial_scene(SCENE_DIR)
    for a in sim.environment.agents.values():
        if overrides and a.id in overrides:
            a.reset_behavior_model(overrides[a.id])
    worst = 0.0
    for step in range(rec["steps"]):
        sim.environment.step(rec["dt"])
        worst = max(worst, compare_trace(sim.environment, rec["trace"][step], step))
    return sim, worst


def learned_overrides(sim_mod, tag):
    """the vehicles make_scenario_golden.py switched to learned behaviours (picked by lane type on the initial match)"""
    if tag == "test_scene":
        return None, 0
    seed_streams(0)
    sim = sim_mod.Simulation(None, sim_mod.SimulationParameter(0.2, 1, render=False, write_gif=False))
    sim.load_initial_scene(SCENE_DIR)
    env = sim.environment
    main_ids = [a.id for a in env.agents.values() if env.corridor_for_agent(a.id).type == "main"]
    merging_ids = [a.id for a in env.agents.values() if env.corridor_for_agent(a.id).type == "merging"]
    if tag == "lane_change_learned":
        return {main_ids[len(main_ids) // 2]: "lane_change_learned", main_ids[1]: "lane_change_learned"}, 1
    return {merging_ids[0]: "merging_learned"}, 2


def evaluation_env(kind, seed):
    """one epoch of generate_random_{merging,lc,exit}_traffic (evaluation_merging.py:4-50, evaluation_lane_change.py:4-55,
    evaluation_exit.py:4-52) with the package's generator, agents through the yaml dict format as the scripts do"""
    import numpy as np
    from interactive_highway_simulation_b200 import agent as ag, outer_env, traffic_generator as tg
    seed_streams(seed)
    choices = [80 / 3.6, 60 / 3.6, 100 / 3.6] if kind == "merging" else [100 / 3.6]
    v_limit = float(np.random.choice(choices, 1)[0])
    spec, params, unsafe, truck = tg.evaluation_map(kind, v_limit=v_limit)
    gen = tg.RandomTrafficGenerator()
    merging, main = gen.random_agents_on_map(tg.lanes_from_spec(spec), {i: tg.RandomDensityAndVelocityParameter(*p) for i, p in params.items()},
                                             unsafe, truck)
    agents = ag.create_agents_from_dicts(merging) + ag.create_agents_from_dicts(main)
    return outer_env.Environment(outer_env.Map([outer_env.create_corridor_from_dict({**c, "left_id": "none" if c["left_id"] is None else c["left_id"],
                                                                                     "right_id": "none" if c["right_id"] is None else c["right_id"]})
                                                for c in spec]), agents)


EVALUATIONS = {   # tag: (map, seed, steps come from the record, who becomes the ego, fall-back rate)  make_scenario_golden.py main()
    "eval_merging_learned": ("merging", 3, lambda env: {a.id: "merging_learned" for a in env.agents.values() if "merging" in a.behavior_model}, None),
    "eval_lane_change_learned": ("lane_change", 4, lambda env: {[a for _, a in env.matched_agents_sorted_by_arc_length[2][2:4] if a.l <= 7][0].id:
                                                                "lane_change_learned"}, None),
    "eval_exit_learned": ("exit", 5, lambda env: {[a for _, a in env.matched_agents_sorted_by_arc_length[2][2:4] if a.l <= 7][0].id: "exit_learned"}, 0.2),
    "eval_exit_closest_gap": ("exit", 6, lambda env: {[a for _, a in env.matched_agents_sorted_by_arc_length[1][2:4] if a.l <= 7][0].id: "exit"}, None),
}


def run_evaluation(tag, rec):
    """evaluate_recorded_*_scene (evaluation_exit.py:55-64 etc.): reset the ego's behaviour model, step the scene"""
    kind, seed, egos, fall_back = EVALUATIONS[tag]
    env = evaluation_env(kind, seed)
    for i, model in egos(env).items():
        a = env.agents[i]
        a.reset_behavior_model(model)
        if fall_back is not None and "learn" in model and hasattr(a, "learned_merging_model"):
            a.learned_merging_model.learned_model_parameter.max_fall_back_rate = fall_back
    worst = 0.0
    for step in range(rec["steps"]):
        env.step(rec["dt"])
        worst = max(worst, compare_trace(env, rec["trace"][step], step))
    return env, worst


@pytest.mark.parametrize("tag", sorted(EVALUATIONS))
def test_evaluation_epochs_host_logic_reproduces_the_reference_trace(monkeypatch, tag):
    """configs[1..3]: generator -> environment -> 12-20 steps with the learned / closest-gap ego"""
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, outer_env
    monkeypatch.setattr(outer_env, "DeviceMap", OracleDeviceMap)
    rec = load_scenario(tag)
    inner = RecordedInnerSimulator(ms, rec["calls"])
    inner.install(monkeypatch)
    _, worst = run_evaluation(tag, rec)
    assert inner.k == len(rec["calls"]) and worst < 1e-9


def stream_states():
    """where the three host random streams stand: numpy's global stream, Python's `random`, the drop-in's seed sequence"""
    import random
    import numpy as np
    from interactive_highway_simulation_b200 import mc_sim_highway as ms
    st = np.random.get_state()
    return (st[1].tobytes(), st[2], st[3], st[4]), random.getstate(), ms._peek_seed()


@pytest.mark.gpu
@pytest.mark.parametrize("walk", ["device_walk", "host_walk"])
@pytest.mark.parametrize("tag", sorted(EVALUATIONS))
def test_evaluation_epochs_on_the_device_reproduce_the_reference_trace(backend_lib, monkeypatch, tag, walk):
    """Both forms of the sequential step: the walk of the vehicles on the device (mce_walk_step: one launch per run of
    device-plannable vehicles) and on the host (one launch per query)."""
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, outer_env
    monkeypatch.setattr(outer_env.Environment, "walk_on_device", walk == "device_walk")
    ms.seed(2000 + EVALUATIONS[tag][1])
    _, worst = run_evaluation(tag, load_scenario(tag))
    assert worst < 1e-9        # kernels == oracle bit for bit, inner answers == reference bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(EVALUATIONS))
def test_device_walk_leaves_the_random_streams_where_the_host_walk_does(backend_lib, monkeypatch, tag):
    """The walk kernel consumes numpy's, Python's and the drop-in's streams in the reference's order: after the same steps
    every vehicle AND all three stream states equal the host path's."""
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, outer_env
    ends = []
    for on_device in (True, False):
        monkeypatch.setattr(outer_env.Environment, "walk_on_device", on_device)
        ms.seed(2000 + EVALUATIONS[tag][1])
        env, _ = run_evaluation(tag, load_scenario(tag))
        vehicles = sorted((a.id, a.x, a.y, a.v, a.a, a.vy, a.commit_lane_change, a.commit_keep_lane_step, a.target_lane_id,
                           tuple(sorted((k, it.yielding, it.initial_probability) for k, it in a.yielding_intention_to_others.items())))
                          for a in list(env.agents.values()) + list(env.removed_agents.values()))
        ends.append((vehicles, stream_states()))
    assert ends[0][0] == ends[1][0]
    assert ends[0][1] == ends[1][1]


@pytest.mark.parametrize("tag", ["test_scene", "lane_change_learned", "merging_learned"])
def test_outer_loop_host_logic_reproduces_the_reference_trace(monkeypatch, tag):
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, outer_env, simulation_multilane as sim_mod
    monkeypatch.setattr(outer_env, "DeviceMap", OracleDeviceMap)
    rec = load_scenario(tag)
    overrides, seed = learned_overrides(sim_mod, tag)
    inner = RecordedInnerSimulator(ms, rec["calls"])
    inner.install(monkeypatch)
    sim, worst = run_test_scene(sim_mod, rec, overrides, seed)
    assert inner.k == len(rec["calls"]), "fewer boundary calls than the reference made"
    assert worst < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("walk", ["device_walk", "host_walk"])
@pytest.mark.parametrize("tag", ["test_scene", "lane_change_learned", "merging_learned"])
def test_outer_loop_on_the_device_reproduces_the_reference_trace(backend_lib, monkeypatch, tag, walk):
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, outer_env, simulation_multilane as sim_mod
    monkeypatch.setattr(outer_env.Environment, "walk_on_device", walk == "device_walk")
    rec = load_scenario(tag)
    overrides, seed = learned_overrides(sim_mod, tag)
    ms.seed(1000 + seed)
    _, worst = run_test_scene(sim_mod, rec, overrides, seed)
    assert worst < 1e-9

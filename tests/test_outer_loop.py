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
@pytest.mark.parametrize("tag", ["test_scene", "lane_change_learned", "merging_learned"])
def test_outer_loop_on_the_device_reproduces_the_reference_trace(backend_lib, tag):
    from interactive_highway_simulation_b200 import mc_sim_highway as ms, simulation_multilane as sim_mod
    rec = load_scenario(tag)
    overrides, seed = learned_overrides(sim_mod, tag)
    ms.seed(1000 + seed)
    run_test_scene(sim_mod, rec, overrides, seed)

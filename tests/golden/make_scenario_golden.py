#!/usr/bin/env python
"""Runs THE REFERENCE'S OWN PYTHON (simulation_multilane / environment / agent / behaviors / mcs_wrapper, imported
unchanged from /root/reference) headless on its shipped test_scene and on learned-behaviour variants of it, with the
reference's own machine code as the inner simulator (tests/refshim/ref_mc_sim_highway.py), and records every call that
crosses the mc_sim_highway boundary with its inputs and the reference's answer:

    python tests/golden/make_scenario_golden.py          (build container only: needs /root/reference and oracle/_ref)

-> tests/golden/scenario_*.json.gz.  Replayed by tests/test_scenario_replay.py through the oracle (CPU) and through the
product's drop-in module on the GPU: identical answers to every recorded call == identical outer simulation.
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
REF = "/root/reference"
sys.path[:0] = [os.path.join(TESTS, "refshim"), ROOT, TESTS, REF]

import numpy as np  # noqa: E402
import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402  (the reference's module, unchanged)


def run_scene(tag, steps, behavior_overrides=None, seed=0, dt=0.2):
    random.seed(seed)
    np.random.seed(seed)
    backend.seed(1000 + seed)
    backend.RECORD.clear()
    param = sm.SimulationParameter(dt, steps, render=False, write_gif=False)
    sim = sm.Simulation(None, param)
    sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
    for a in sim.environment.agents.values():
        if behavior_overrides and a.id in behavior_overrides:
            a.reset_behavior_model(behavior_overrides[a.id])
    trace = []
    for step in range(steps):
        sim.environment.step(dt)
        trace.append({str(a.id): [a.x, a.y, a.v, a.a] for a in sim.environment.agents.values()})
    rec = {"tag": tag, "kind": "scenario", "steps": steps, "dt": dt, "calls": list(backend.RECORD), "trace": trace}
    path = os.path.join(HERE, "scenario_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(rec, f, separators=(",", ":"))
    fns = {}
    for c in rec["calls"]:
        fns[c["fn"]] = fns.get(c["fn"], 0) + 1
    print(tag, "steps", steps, "calls", fns, "->", os.path.getsize(path), "bytes")


def record_env(tag, env, steps, ego_policies, seed, dt=0.2, fall_back=None):
    """evaluate_recorded_*_scene (evaluation_exit.py:55-64 etc.): reset the ego's behaviour model, step the scene."""
    backend.seed(2000 + seed)
    backend.RECORD.clear()
    for a in env.agents.values():
        if a.id in ego_policies:
            a.reset_behavior_model(ego_policies[a.id])
            if fall_back is not None and "learn" in ego_policies[a.id] and hasattr(a, "learned_merging_model"):
                a.learned_merging_model.learned_model_parameter.max_fall_back_rate = fall_back
    trace = []
    for step in range(steps):
        env.step(dt)
        trace.append({str(a.id): [a.x, a.y, a.v, a.a] for a in env.agents.values()})
    rec = {"tag": tag, "kind": "scenario", "steps": steps, "dt": dt, "calls": list(backend.RECORD), "trace": trace}
    path = os.path.join(HERE, "scenario_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(rec, f, separators=(",", ":"))
    fns = {}
    for c in rec["calls"]:
        fns[c["fn"]] = fns.get(c["fn"], 0) + 1
    print(tag, "steps", steps, "agents", len(env.agents), "calls", fns, "->", os.path.getsize(path), "bytes")


def evaluation_env(kind, seed):
    """One epoch of generate_random_{merging,lc,exit}_traffic (evaluation_merging.py:4-50, evaluation_lane_change.py:4-55,
    evaluation_exit.py:4-52), built with the reference's own Corridor / Map / RandomTrafficGenerator / Environment."""
    random.seed(seed); np.random.seed(seed)
    R = sm.RandomDensityAndVelocityParameter
    gen = sm.RandomTrafficGenerator()
    if kind == "merging":
        params = {0: R(1.0, 0.1, -2, 2, None, 2), 1: R(1.6, 0.4, 0, 2, [0, 400], None), 2: R(2.0, 1.0, 0, 2, [0, 400], None)}
        v_limit = float(np.random.choice([80 / 3.6, 60 / 3.6, 100 / 3.6], 1)[0])
        length = v_limit * 3.6 * 2 + 20
        cs = [sm.Corridor(0, "merging", [[0, 3.5], [length + 10, 3.5]], [[0, 0], [length, 0]], [[0, 1.75], [length + 5, 1.75]], v_limit),
              sm.Corridor(1, "main", [[-100, 7], [500, 7]], [[-100, 3.5], [500, 3.5]], [[-100, 5.25], [500, 5.25]], v_limit),
              sm.Corridor(2, "main", [[-100, 10.5], [500, 10.5]], [[-100, 7], [500, 7]], [[-100, 8.75], [500, 8.75]], v_limit + 20 / 3.6)]
        cs[0].set_left_and_right_corridor_id(1, None); cs[1].set_left_and_right_corridor_id(2, 0); cs[2].set_left_and_right_corridor_id(None, 1)
    else:
        first = R(1.0, 0.1, -2, 2, None, 2 if kind == "lane_change" else 0)
        second = R(1.8 if kind == "lane_change" else 1.6, 0.4, 0, 2, [0, 500], None)
        params = {0: first, 1: second, 2: R(2.0, 1.0, 0, 2, [0, 500], None), 3: R(2.0, 1.0, 0, 2, [0, 500], None)}
        v_limit = float(np.random.choice([100 / 3.6], 1)[0])
        if kind == "lane_change":
            length = v_limit ** 2 / 4 + 20
            c0 = sm.Corridor(0, "merging", [[-10, 3.5], [length + 10, 3.5]], [[-10, 0], [length, 0]], [[-10, 1.75], [length + 5, 1.75]], v_limit)
        else:
            c0 = sm.Corridor(0, "exit", [[245, 3.5], [500, 3.5]], [[255, 0], [500, 0]], [[250, 1.75], [500 + 5, 1.75]], v_limit)
        cs = [c0,
              sm.Corridor(1, "main", [[-200, 7], [500, 7]], [[-200, 3.5], [500, 3.5]], [[-200, 5.25], [500, 5.25]], v_limit),
              sm.Corridor(2, "main", [[-200, 10.5], [500, 10.5]], [[-200, 7], [500, 7]], [[-200, 8.75], [500, 8.75]], v_limit + 20 / 3.6),
              sm.Corridor(3, "main", [[-200, 14.0], [500, 14.0]], [[-200, 10.5], [500, 10.5]], [[-200, 12.25], [500, 12.25]], v_limit + 30 / 3.6)]
        cs[0].set_left_and_right_corridor_id(1, None); cs[1].set_left_and_right_corridor_id(2, 0)
        cs[2].set_left_and_right_corridor_id(3, 1); cs[3].set_left_and_right_corridor_id(None, 2)
    m = sm.Map(cs)
    agents_merging, agents_main = gen.random_agents_on_map(m, params, 0., 0.4)
    agents = sm.create_agents_from_dicts(sm.create_dicts_from_agents(agents_merging or [])) + \
        sm.create_agents_from_dicts(sm.create_dicts_from_agents(agents_main))
    return sm.Environment(m, agents)


def main():
    steps = int(os.environ.get("SCENARIO_STEPS", "100"))      # configs[0]: dt 0.2, 100 steps
    # configs[0]: the shipped scene as is (idm / idm_mobil_lane_change_safe / merging_closest_gap_policy agents)
    run_scene("test_scene", steps)
    # learned behaviours on the same scene: find suitable vehicles by their lane type
    random.seed(0); np.random.seed(0)
    sim = sm.Simulation(None, sm.SimulationParameter(0.2, 1, render=False, write_gif=False))
    sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
    env = sim.environment
    main_ids = [a.id for a in env.agents.values() if env.corridor_for_agent(a.id).type == "main"]
    merging_ids = [a.id for a in env.agents.values() if env.corridor_for_agent(a.id).type == "merging"]
    run_scene("lane_change_learned", 10, {main_ids[len(main_ids) // 2]: "lane_change_learned", main_ids[1]: "lane_change_learned"}, seed=1)
    run_scene("merging_learned", 10, {merging_ids[0]: "merging_learned"}, seed=2)
    # configs[1..3]: one epoch of each evaluation script's random scene, egos chosen as the scripts choose them
    env = evaluation_env("merging", 3)                                      # evaluation_merging.py:56-62: every merging agent
    egos = {a.id: "merging_learned" for a in env.agents.values() if "merging" in a.behavior_model}
    record_env("eval_merging_learned", env, 12, egos, 3)
    env = evaluation_env("lane_change", 4)                                  # evaluation_lane_change.py: 2nd..4th vehicle of a lane
    lane2 = [a for _, a in env.matched_agents_sorted_by_arc_length[2][2:4] if a.l <= 7]
    record_env("eval_lane_change_learned", env, 12, {lane2[0].id: "lane_change_learned"}, 4)
    env = evaluation_env("exit", 5)                                         # evaluation_exit.py:123-128
    lane2 = [a for _, a in env.matched_agents_sorted_by_arc_length[2][2:4] if a.l <= 7]
    record_env("eval_exit_learned", env, 12, {lane2[0].id: "exit_learned"}, 5, fall_back=0.2)
    env = evaluation_env("exit", 6)
    lane1 = [a for _, a in env.matched_agents_sorted_by_arc_length[1][2:4] if a.l <= 7]
    record_env("eval_exit_closest_gap", env, 20, {lane1[0].id: "exit"}, 6)


if __name__ == "__main__":
    main()

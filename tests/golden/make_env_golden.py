#!/usr/bin/env python
"""Writes tests/golden/env_states.json.gz with THE REFERENCE'S OWN PYTHON (environment.py / type.py / utility.py /
mcs_wrapper.py imported unchanged from /root/reference; shapely replaced by tests/refshim): for a set of environment
states it records the inputs (map, every vehicle) and the reference's answers for

  * Environment.match_agents_on_corridor: lane of every vehicle, the per-lane vectors sorted by arc length
  * front_agent / following_agent / neighbor_around_agent of every vehicle (ids and IdmMatch.dis), also for vehicles
    that have moved since the match (the reference's sequential plan-and-move loop)
  * sorted_agents_on_corridor_within_range_of_vehicle
  * LineSimplified(extend_line_both_sides(center, 1000)).to_frenet_frame of every vehicle on every corridor
  * step_along_path
  * merging_mcs_sim_from_env / exit_mcs_sim_from_env / lane_change_mcs_sim_from_env (mcs_wrapper.py) with numpy's
    random stream seeded per call

    python tests/golden/make_env_golden.py          (build container only: needs /root/reference)
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
sys.path[:0] = [os.path.join(TESTS, "refshim"), ROOT, TESTS, REF, HERE]

import numpy as np  # noqa: E402
import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402
import mcs_wrapper as mw  # noqa: E402
from utility import step_along_path, extend_line_both_sides  # noqa: E402
from type import LineSimplified  # noqa: E402
from make_scenario_golden import evaluation_env  # noqa: E402

NB = ("left_ff", "left_f", "left_b", "left_bb", "f", "b", "right_ff", "right_f", "right_b", "right_bb")


def fl(v):
    return float(v)


def corridor_rec(c):
    return {"id": c.id, "type": c.type, "left": [[fl(p[0]), fl(p[1])] for p in c.left], "right": [[fl(p[0]), fl(p[1])] for p in c.right],
            "center": [[fl(p[0]), fl(p[1])] for p in c.center], "v_limit": fl(c.v_limit), "left_id": c.left_id, "right_id": c.right_id}


def agent_rec(a):
    ip, rp, mp = a.idm_param, a.rss_param, a.mobil_param
    return {"id": a.id, "x": fl(a.x), "y": fl(a.y), "v": fl(a.v), "vy": fl(a.vy), "a": fl(a.a), "l": fl(a.l), "w": fl(a.w),
            "behavior_model": a.behavior_model,
            "idm": [fl(ip.vd), fl(ip.acc_max), fl(ip.acc_min), fl(ip.min_dis), fl(ip.acc_com), fl(ip.thw)],
            "rss": [fl(rp.r_t_other), fl(rp.r_t_ego), fl(rp.a_min), fl(rp.a_max), fl(rp.soft_acc), fl(rp.soft_dcc)],
            "mobil": [fl(mp.politeness_factor), fl(mp.delta_a), fl(mp.bias_a)],
            "yielding": [fl(v) for v in a.yielding_param]}


def neighbor_rec(env, idx):
    nb = env.neighbor_around_agent(idx)
    return [[getattr(nb, k).id, fl(getattr(nb, k).dis)] if getattr(nb, k) else None for k in NB]


def ms_corridor_rec(c):
    return [c.id, c.type, c.left.p1.x, c.left.p1.y, c.left.p2.x, c.left.p2.y, c.right.p1.x, c.right.p1.y, c.right.p2.x, c.right.p2.y,
            c.center.p1.x, c.center.p1.y, c.center.p2.x, c.center.p2.y, c.vLimit]


def ms_agent_rec(a):
    return [a.id, a.x, a.y, a.vx, a.vy, a.vd, a.a, a.l, a.w] + list(a.idm._memory_order()) + list(a.rss._memory_order()) + \
        list(a.yp._memory_order()) + [a.behavior]


def builder_rec(env, name, agent, seed):
    fn = getattr(mw, name)
    np.random.seed(seed)
    got = fn(env, agent, env.corridor_for_agent(agent.id))
    corridors, agents = got[0], got[1]
    actions = got[2] if len(got) > 2 else None            # lane_change_mcs_sim_from_env returns two values
    return {"builder": name, "ego": agent.id, "np_seed": seed, "corridors": [ms_corridor_rec(c) for c in corridors],
            "agents": [ms_agent_rec(a) for a in agents], "possible_actions": [int(v) for v in actions] if actions is not None else None}


def decision_recs(env, rng):
    """behaviors.py:89-176 on the reference's environment with the reference's machine code as inner simulator: the
    chosen command and the resulting action of a few egos per behaviour."""
    import behaviors as bh
    out = []
    has_exit = any(k.type == "exit" for k in env.map.corridors.values())
    picks = []
    for a in env.agents.values():
        c = env.corridor_for_agent(a.id)
        if c.type == "merging" and c.left_id is not None:
            picks.append((a, "merging_learned", "cloned_merging_behavior"))
            picks.append((a, "merging_closest_gap_policy", "merging_behavior_closest_gap"))
        elif c.type == "main" and has_exit and a.l <= 7 and rng.random() < 0.2:
            picks.append((a, "exit_learned", "cloned_exit_behavior"))
            picks.append((a, "exit", "exit_behavior_closest_gap"))
        elif c.type == "main" and a.l <= 7 and rng.random() < 0.12:
            picks.append((a, "lane_change_learned", "lane_change_behavior_learned"))
    seed = 500
    for a, model, fn in picks[:8]:
        before = a.behavior_model
        a.reset_behavior_model(model)
        np.random.seed(seed); backend.seed(7000 + seed); backend.RECORD.clear()
        act = getattr(bh, fn)(env, a)
        out.append({"fn": fn, "ego": a.id, "np_seed": seed, "backend_seed": 7000 + seed, "agent": agent_rec(a),
                    "ax": fl(act.ax), "vy": fl(act.vy), "decision": getattr(a, "current_decision", None),
                    "ref_lane": env.corridor_for_agent(a.id).id, "calls": len(backend.RECORD)})
        a.reset_behavior_model(before)
        seed += 1
    backend.RECORD.clear()
    return out


def state_rec(tag, env, all_agents, rng, builders=True):
    """env was just matched on the positions of all_agents (dict order)."""
    rec = {"tag": tag, "corridors": [corridor_rec(c) for c in env.map.corridors.values()], "agents": all_agents}
    rec["lane"] = {str(a["id"]): (env.matched_agents[a["id"]].corridor_id if a["id"] in env.matched_agents else None) for a in all_agents}
    rec["lane_vectors"] = {str(c): [[fl(k), a.id] for k, a in lst] for c, lst in env.matched_agents_sorted_by_arc_length.items()}
    rec["neighbors"] = {str(i): neighbor_rec(env, i) for i in env.agents}
    ids = list(env.agents)
    rec["within_range"] = []
    for _ in range(min(12, len(ids))):
        ego = rng.choice(ids)
        c = rng.choice(list(env.map.corridors))
        d = rng.choice([80, 30, 200])
        got = env.sorted_agents_on_corridor_within_range_of_vehicle(c, ego, d)
        rec["within_range"].append({"corridor": c, "ego": ego, "max_dis": d, "ids": [a.id for a in got]})
    rec["frenet"] = {}
    for c in env.map.corridors.values():
        ref = LineSimplified(extend_line_both_sides(c.center, 1000))
        rec["frenet"][str(c.id)] = [[fl(v) for v in ref.to_frenet_frame([a.x, a.y])] for a in env.agents.values()]
    rec["frenet_ids"] = ids
    rec["steps"] = []
    for a in env.agents.values():
        ax, vy, dt = rng.uniform(-6, 2), rng.choice([0.0, 0.0, 1.0, -1.0, 0.37]), rng.choice([0.2, 0.3])
        c = env.corridor_for_agent(a.id)
        x, y, v, yaw = step_along_path(c.centerSimplified, a, ax, vy, dt)
        rec["steps"].append({"id": a.id, "lane": c.id, "ax": ax, "vy": vy, "dt": dt, "out": [fl(x), fl(y), fl(v), fl(yaw)]})
    rec["builders"] = []
    if builders:
        seed = 100
        for a in env.agents.values():
            c = env.corridor_for_agent(a.id)
            left = env.map.corridors[c.left_id] if c.left_id is not None else None
            if c.type == "merging" and left is not None:
                rec["builders"].append(builder_rec(env, "merging_mcs_sim_from_env", a, seed)); seed += 1
            has_exit = any(k.type == "exit" for k in env.map.corridors.values())
            if c.type == "main" and has_exit and rng.random() < 0.25:
                rec["builders"].append(builder_rec(env, "exit_mcs_sim_from_env", a, seed)); seed += 1
            if c.type == "main" and rng.random() < 0.25:
                rec["builders"].append(builder_rec(env, "lane_change_mcs_sim_from_env", a, seed)); seed += 1
    rec["decisions"] = decision_recs(env, rng) if builders else []
    # the sequential loop: the first vehicles move, nobody is re-matched, and they and their neighbours are asked again
    moved = []
    for a in list(env.agents.values())[: max(3, len(env.agents) // 3)]:
        c = env.corridor_for_agent(a.id)
        a.x, a.y, a.v, _ = step_along_path(c.centerSimplified, a, rng.uniform(-3, 1.5), rng.choice([0.0, 1.0, -1.0]), rng.choice([0.2, 2.0]))
        moved.append({"id": a.id, "x": fl(a.x), "y": fl(a.y)})
    rec["moved"] = moved
    rec["neighbors_after_move"] = {str(i): neighbor_rec(env, i) for i in env.agents}
    return rec


def perturb(env, rng, mode):
    agents = list(env.agents.values())
    for a in agents:
        if mode == "jitter":
            a.y += rng.gauss(0, 1.2)
            a.x += rng.gauss(0, 3.0)
        elif mode == "ties":
            a.x = float(round(a.x / 25.0) * 25.0)             # equal arc lengths inside a lane and across lanes
        elif mode == "edges":
            r = rng.random()
            if r < 0.15:
                a.y = rng.choice([0.0, 3.5, 7.0, 10.5])        # exactly on a border line: within() is False or next lane
            elif r < 0.25:
                a.x = rng.choice([-500.0, 900.0])              # beyond the map: removed
            elif r < 0.4:
                a.x = rng.choice([499.0, 499.99, -99.5, 480.0 + rng.random() * 30])


def main():
    rng = random.Random(20261020)
    out = []
    # the shipped scene
    random.seed(0); np.random.seed(0)
    sim = sm.Simulation(None, sm.SimulationParameter(0.2, 1, render=False, write_gif=False))
    sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
    envs = [("test_scene", sim.environment)]
    for kind, seeds in (("merging", (11, 12)), ("lane_change", (13, 14)), ("exit", (15, 16))):
        for s in seeds:
            envs.append(("%s_seed%d" % (kind, s), evaluation_env(kind, s)))
    for tag, env in envs:
        all_agents = [agent_rec(a) for a in env.agents.values()]
        out.append(state_rec(tag, env, all_agents, rng))
        print(tag, len(all_agents), "agents", len(out[-1]["builders"]), "builder calls")
    for mode in ("jitter", "ties", "edges"):
        for kind, s in (("merging", 21), ("lane_change", 22), ("exit", 23)):
            env = evaluation_env(kind, s)
            perturb(env, rng, mode)
            all_agents = [agent_rec(a) for a in env.agents.values()]
            env.match_agents_on_corridor()
            tag = "%s_%s" % (kind, mode)
            out.append(state_rec(tag, env, all_agents, rng, builders=(mode == "jitter")))
            print(tag, len(all_agents), "agents,", len(env.agents), "matched,", len(out[-1]["builders"]), "builder calls")
    path = os.path.join(HERE, "env_states.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

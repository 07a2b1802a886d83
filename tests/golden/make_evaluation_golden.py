#!/usr/bin/env python
"""Writes tests/golden/evaluation_stats.json.gz with THE REFERENCE'S OWN evaluation scripts (evaluation_merging.py,
evaluation_lane_change.py, evaluation_exit.py, imported unchanged; shapely / matplotlib replaced by tests/refshim, the
reference's machine code as mc_sim_highway): per scenario one generated epoch is dumped with the reference's
Simulation.dump_initial_scene, the egos are chosen as the scripts choose them, and evaluate_recorded_*_scene is run for
the baseline and the learned policy.  Recorded: the dumped scene (agents / map dicts), every boundary call with the
reference's answer, the returned console text and the statistics object after each evaluation.

    python tests/golden/make_evaluation_golden.py          (build container only: needs /root/reference and oracle/_ref)
"""
import gzip
import json
import os
import random
import sys
import tempfile

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
import evaluation_exit as ee  # noqa: E402
import evaluation_lane_change as el  # noqa: E402
import evaluation_merging as em  # noqa: E402
from make_scenario_golden import evaluation_env  # noqa: E402

STEPS = int(os.environ.get("EVALUATION_STEPS", "12"))


def plain(v):
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int, bool, np.bool_)):
        return int(v)
    return v


def stats_rec(st):
    return {k: plain(v) for k, v in st.__dict__.items()}


def main():
    out = []
    tmp = tempfile.mkdtemp(prefix="evalgolden")
    for kind, seed in (("merging", 41), ("lane_change", 42), ("exit", 43)):
        env = evaluation_env(kind, seed)
        param = sm.SimulationParameter(0.2, STEPS, render=False, write_gif=False, show_debug_info=False, scenario=kind)
        sim = sm.Simulation(None, param)
        sim.write_dir = os.path.join(tmp, kind)
        sim.reset(env, param)
        sim.dump_initial_scene()
        path = os.path.join(sim.write_dir, "epoch0")
        rec = {"kind": kind, "steps": STEPS, "agents": sm.load_yaml(os.path.join(path, "agents.yml")),
               "map": sm.load_yaml(os.path.join(path, "map.yml")), "evaluations": []}
        sim.reset(None, sim.param)
        sim.load_initial_scene(path)
        if kind == "merging":
            egos = [None]
            runs = [("merging_closest_gap_policy", 1.0), ("merging_learned", 1.0), ("merging_learned", 0.2)]
        else:
            egos = []
            for c_id, sorted_agents in sim.environment.matched_agents_sorted_by_arc_length.items():
                if kind == "lane_change" and sim.environment.map.corridors[c_id].type == "main":
                    egos += [a.id for _, a in sorted_agents[2:5] if a.l <= 7]
                if kind == "exit" and c_id in (2, 3):
                    egos += [a.id for _, a in sorted_agents[2:4] if a.l <= 7]
            rec["ego_candidates"] = list(egos)
            egos = egos[:1]
            runs = [("idm", None), ("idm_mobil_lane_change_safe", None), ("lane_change_learned", None)] if kind == "lane_change" else \
                [("exit", 1.0), ("exit_learned", 1.0), ("exit_learned", 0.2)]
        stats = {}
        n = 0
        for ego in egos:
            for policy, fb in runs:
                key = "%s/%s" % (policy, fb)
                if key not in stats:
                    stats[key] = sm.LaneChangeStatistics() if kind == "lane_change" else sm.MergingAndExitStatistics()
                s = 9000 + 10 * seed + n
                n += 1
                random.seed(s); np.random.seed(s); backend.seed(s); backend.RECORD.clear()
                if kind == "merging":
                    text = em.evaluate_recorded_merging_scene(path, sim, policy, 0, stats[key], fb)
                elif kind == "lane_change":
                    text = el.evaluate_recorded_lc_scene(path, sim, policy, 0, ego, stats[key])
                else:
                    text = ee.evaluate_recorded_exit_scene(path, sim, policy, 0, ego, stats[key], fb)
                rec["evaluations"].append({"policy": policy, "ego": ego, "fall_back": fb, "seed": s, "stats_key": key, "text": text,
                                           "calls": list(backend.RECORD), "statistics": stats_rec(stats[key]),
                                           "executed_steps": sim.step + 1})
                print(kind, policy, ego, fb, "calls", len(backend.RECORD), "|", text.strip().splitlines()[-1][:150])
        out.append(rec)
    path = os.path.join(HERE, "evaluation_stats.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

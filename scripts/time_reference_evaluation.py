#!/usr/bin/env python
"""CPU baseline of the evaluation scripts (build container only: needs /root/reference and oracle/_ref).

Times THE REFERENCE'S OWN evaluate_recorded_{merging,lc,exit}_scene (evaluation_merging.py:53-106, evaluation_lane_change.py:
58-101, evaluation_exit.py:55-105, imported unchanged) with the LEARNED ego policy on the epochs of
tests/golden/evaluation_stats.json.gz — the scenes bench.py's `evaluation` block runs through the product — for the
scripts' own horizon (dt 0.2, 100 steps), over tests/refshim (pure-Python shapely stand-in, matplotlib stub) and the
reference's machine code behind a pipe as `mc_sim_highway` (its 8-thread runMultiThreads per candidate).  The reference
cannot travel to the GPU box: the figures are recorded here and bench.py reports them with this label.
    python scripts/time_reference_evaluation.py  ->  profiles/r02_reference_evaluation_cpu.json
"""
import gzip
import json
import os
import platform
import random
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TESTS = os.path.join(ROOT, "tests")
REF = "/root/reference"
sys.path[:0] = [os.path.join(TESTS, "refshim"), ROOT, TESTS, REF]

import numpy as np  # noqa: E402
import yaml  # noqa: E402
import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402
import evaluation_exit as ee  # noqa: E402
import evaluation_lane_change as el  # noqa: E402
import evaluation_merging as em  # noqa: E402

STEPS = int(os.environ.get("EVALUATION_STEPS", "100"))
LEARNED = {"merging": "merging_learned", "lane_change": "lane_change_learned", "exit": "exit_learned"}


def main():
    with gzip.open(os.path.join(TESTS, "golden", "evaluation_stats.json.gz"), "rt") as f:
        records = json.load(f)
    tmp = tempfile.mkdtemp(prefix="evaltime")
    out = {}
    for rec in records:
        kind = rec["kind"]
        path = os.path.join(tmp, kind, "epoch0")
        os.makedirs(path)
        with open(os.path.join(path, "agents.yml"), "w") as f:
            yaml.dump({int(k): v for k, v in rec["agents"].items()}, f)
        with open(os.path.join(path, "map.yml"), "w") as f:
            yaml.dump({int(k): v for k, v in rec["map"].items()}, f)
        sim = sm.Simulation(None, sm.SimulationParameter(0.2, STEPS, render=False, write_gif=False, show_debug_info=False, scenario=kind))
        ego = None if kind == "merging" else rec["ego_candidates"][0]
        stats = sm.LaneChangeStatistics() if kind == "lane_change" else sm.MergingAndExitStatistics()
        random.seed(7); np.random.seed(7); backend.seed(7); backend.RECORD.clear()
        t0 = time.perf_counter()
        if kind == "merging":
            em.evaluate_recorded_merging_scene(path, sim, LEARNED[kind], 0, stats, 1.0)
        elif kind == "lane_change":
            el.evaluate_recorded_lc_scene(path, sim, LEARNED[kind], 0, ego, stats)
        else:
            ee.evaluate_recorded_exit_scene(path, sim, LEARNED[kind], 0, ego, stats, 1.0)
        wall = time.perf_counter() - t0
        decisions = sum(1 for c in backend.RECORD if c["fn"] in ("fixedGapMergingBehavior", "fixedDirectionLaneChangeBehavior"))
        out[kind] = {"policy": LEARNED[kind], "vehicles": len(rec["agents"]), "executed_steps": sim.step + 1, "wall_s": wall,
                     "evaluations_per_s": 1.0 / wall, "learned_decisions": decisions, "boundary_calls": len(backend.RECORD)}
        print(kind, out[kind])
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    res = {"horizon": "dt 0.2, %d steps" % STEPS, "kinds": out, "kind": "reference",
           "host": "%s (%d vCPU), %s" % (cpu, os.cpu_count(), platform.python_version()),
           "what": "the reference's own evaluate_recorded_*_scene with the learned ego policy, unchanged, over tests/refshim and the "
                   "reference's machine code (runMultiThreads, 8 threads) behind a pipe; recorded in the build container, NOT on the GPU box"}
    with open(os.path.join(ROOT, "profiles", "r02_reference_evaluation_cpu.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

"""Host profile (cProfile) of one learned lane-change evaluation epoch (diagnostic; needs a GPU)."""
import cProfile, gzip, json, os, pstats, random, sys, tempfile
import numpy as np
import yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interactive_highway_simulation_b200 import evaluation as ev, mc_sim_highway as ms
kind = sys.argv[1] if len(sys.argv) > 1 else "lane_change"
with gzip.open(os.path.join(ROOT, "tests", "golden", "evaluation_stats.json.gz"), "rt") as f:
    rec = [r for r in json.load(f) if r["kind"] == kind][0]
path = os.path.join(tempfile.mkdtemp(prefix="prof_eval"), "epoch0"); os.makedirs(path)
for name in ("agents", "map"):
    with open(os.path.join(path, name + ".yml"), "w") as f:
        yaml.dump({int(k): v for k, v in rec[name].items()}, f, default_flow_style=False)
policy = {"merging": "merging_learned", "lane_change": "lane_change_learned", "exit": "exit_learned"}[kind]
for rep in range(2):
    sim = ev.Simulation(None, ev.SimulationParameter(0.2, 100, render=False, write_gif=False, scenario=kind))
    st = ev.LaneChangeStatistics() if kind == "lane_change" else ev.MergingAndExitStatistics()
    random.seed(7); np.random.seed(7); ms.seed(7)
    if rep == 1:
        pr = cProfile.Profile(); pr.enable()
    if kind == "merging":
        ev.evaluate_recorded_merging_scene(path, sim, policy, 0, st, 1.0)
    elif kind == "lane_change":
        ev.evaluate_recorded_lc_scene(path, sim, policy, 0, rec["ego_candidates"][0], st)
    else:
        ev.evaluate_recorded_exit_scene(path, sim, policy, 0, rec["ego_candidates"][0], st, 1.0)
pr.disable()
pstats.Stats(pr).sort_stats(sys.argv[2] if len(sys.argv) > 2 else "cumulative").print_stats(38)

"""BASELINE-size golden vectors straight from the reference binary (build container only: needs oracle/_ref).

Same record format as make_golden.py's `rollouts` fixtures, gzip-compressed, no per-step histories (the outcome record —
utility / comfort over the whole trajectory of every vehicle, finish step, draw count — is a checksum of the rollout):
  large_lc100_*    configs[2] shape: 4 main lanes, 100 vehicles, lane-change commands
  large_exit200    configs[3] shape: exit lane + 3 main lanes, 200 vehicles, 50-step horizon, ego `exit`
  large_lc256      configs[4] shape: 256 vehicles x 100 steps, ego keep_lane, minSteps = steps
  large_merge64_*  merging lane + 3 main lanes, 64 vehicles (block-size boundary), two gaps
  *_decisions      the Python-facing single-decision wrappers on the 100- and 64-vehicle scenes
usage: python tests/golden/make_large_golden.py
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg   # noqa: E402  (RefHarness plumbing, rollout_case)
import scenes               # noqa: E402


def main():
    cases = []
    rng = random.Random(20261020)
    s, ego = scenes.random_scene(rng, 4, 100, spacing=(3, 45))
    for act in ("keep_lane", "left", "acc"):
        cases.append(mg.rollout_case("large_lc100_%s" % act, s, ego, act, -2, 0.3, 40, 30, 11, 0, 4, n_hist=0))
    # the single-decision wrappers (MOBIL with / without RSS, the five fixed directions) for every fourth vehicle
    cases.append(mg.decision_cases("large_lc100_decisions", s, [a["id"] for a in s.agents[::4]]))
    s, ego = scenes.random_scene(rng, n_agents=200, lane_types=["exit", "main", "main", "main"], ego_behavior="exit", ego_lane=1,
                                 spacing=(3, 45))
    cases.append(mg.rollout_case("large_exit200", s, ego, "right", -1, 0.3, 50, 10, 12, 0, 3, n_hist=0))
    s, ego = scenes.random_scene(rng, 4, 256, spacing=(3, 45))
    cases.append(mg.rollout_case("large_lc256", s, ego, "keep_lane", -2, 0.3, 100, 100, 13, 0, 2, n_hist=0))
    s, ego = scenes.random_scene(rng, n_agents=64, lane_types=["merging", "main", "main", "main"], ego_behavior="merging", ego_lane=0,
                                 spacing=(3, 45))
    on_left = sorted((a for a in s.agents if 3.5 <= a["y"] < 7.0), key=lambda a: a["x"])
    gaps = [-1, on_left[len(on_left) // 2]["id"]]
    for gi, gap in enumerate(gaps):
        cases.append(mg.rollout_case("large_merge64_gap%d" % gi, s, ego, "left", gap, 0.3, 40, 10, 14, 0, 3, n_hist=0))
    cases.append(mg.merging_decision_cases("large_merge64_decisions", s, "left", gaps))
    for c in cases:
        with gzip.open(os.path.join(HERE, c["tag"] + ".json.gz"), "wt") as f:
            json.dump(c, f, separators=(",", ":"))
        print(c["tag"], [(r.get("ok"), r.get("finishStep"), r.get("draws")) for r in c["rollouts"]] if "rollouts" in c else "%d calls" % len(c["calls"]))
    print("wrote %d fixtures" % len(cases))


if __name__ == "__main__":
    main()

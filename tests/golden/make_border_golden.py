#!/usr/bin/env python
"""Writes tests/golden/env_border.json.gz with THE REFERENCE'S OWN type.py (Corridor, LineSimplified; shapely replaced by
tests/refshim): for straight evaluation maps, the shipped test_scene map and a bent three-lane map it records, for
random vehicle poses around every corridor,

  * Corridor.distance_to_border(hull, "left" / "right")            type.py:183-206 (hull: utility.py:12-20)
  * corridor.centerSimplified.signed_distance([x, y])             type.py:307-315
  * Corridor.distance_to_corridor_end([x, y])                      type.py:180-181

    python tests/golden/make_border_golden.py          (build container only: needs /root/reference)
"""
import gzip
import json
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
REF = "/root/reference"
sys.path[:0] = [os.path.join(TESTS, "refshim"), ROOT, TESTS, REF, HERE]

import ref_mc_sim_highway as backend  # noqa: E402

sys.modules["mc_sim_highway"] = backend
os.chdir(REF)
import simulation_multilane as sm  # noqa: E402
from utility import vehicle_pose_to_rect  # noqa: E402
from make_env_golden import corridor_rec  # noqa: E402
from make_scenario_golden import evaluation_env  # noqa: E402


def bent_corridors():
    def line(off):
        pts = [(-50.0, 0.0), (60.0, 0.0), (140.0, 12.0), (260.0, 12.0), (330.0, -6.0)]
        out = []
        for i, (x, y) in enumerate(pts):
            a, b = pts[max(i - 1, 0)], pts[min(i + 1, len(pts) - 1)]
            yaw = math.atan2(b[1] - a[1], b[0] - a[0])
            out.append([x - off * math.sin(yaw), y + off * math.cos(yaw)])
        return out
    cs = []
    for k in range(3):
        cs.append(sm.Corridor(k, "main", line(3.5 * (k + 1)), line(3.5 * k), line(3.5 * k + 1.75), 25.0))
    cs[0].set_left_and_right_corridor_id(1, None); cs[1].set_left_and_right_corridor_id(2, 0); cs[2].set_left_and_right_corridor_id(None, 1)
    return cs


def main():
    rng = random.Random(77)
    maps = []
    sim = sm.Simulation(None, sm.SimulationParameter(0.2, 1, render=False, write_gif=False))
    sim.load_initial_scene(os.path.join(REF, "test_scene", "epoch0"))
    maps.append(("test_scene", list(sim.environment.map.corridors.values())))
    for kind in ("merging", "lane_change", "exit"):
        maps.append((kind, list(evaluation_env(kind, 31).map.corridors.values())))
    maps.append(("bent", bent_corridors()))
    out = []
    for tag, corridors in maps:
        rec = {"tag": tag, "corridors": [corridor_rec(c) for c in corridors], "samples": []}
        for c in corridors:
            for k in range(40):
                x_min, x_max = c.x_min, c.x_max
                s = rng.uniform(0, c.centerSimplified.path_length) if k % 8 else rng.choice([0.0, c.centerSimplified.path_length + 3.0])
                lat = rng.choice([0.0, rng.gauss(0, 0.8), rng.uniform(-2.5, 2.5), 1.75, -1.75, 0.85])
                x, y, yaw = c.centerSimplified.to_global_coordinate(s, lat)
                x, y, yaw = float(x), float(y), float(yaw)
                yaw += rng.choice([0.0, 0.0, rng.gauss(0, 0.08), rng.uniform(-0.6, 0.6)])
                if k % 10 == 9:
                    x = rng.uniform(x_min - 20, x_max + 20)
                l, w = rng.choice([(4.5, 1.8), (5.0, 2.0), (12.0, 2.5)])
                hull = vehicle_pose_to_rect(x, y, yaw, l, w)
                rec["samples"].append({
                    "corridor": c.id, "x": x, "y": y, "hull": [[float(p[0]), float(p[1])] for p in hull],
                    "left": float(c.distance_to_border(hull, "left")), "right": float(c.distance_to_border(hull, "right")),
                    "signed": float(c.centerSimplified.signed_distance([x, y])),
                    "to_end": float(c.distance_to_corridor_end([x, y]))})
        out.append(rec)
        print(tag, len(rec["samples"]), "samples",
              sum(1 for s in rec["samples"] if s["left"] < 0), "past left,", sum(1 for s in rec["samples"] if s["right"] < 0), "past right,",
              sum(1 for s in rec["samples"] if s["left"] == 0 or s["right"] == 0), "touching")
    path = os.path.join(HERE, "env_border.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Runs THE REFERENCE'S OWN random_traffic_generator.py (imported unchanged from /root/reference; shapely replaced by
tests/refshim) on the three maps its evaluation scripts build, for a few seeds, and records every generated agent:

    python tests/golden/make_generator_golden.py          (build container only: needs /root/reference)

-> tests/golden/generator_scenes.json.gz.  tests/test_traffic_generator.py replays the same seeds through the package's
port (interactive_highway_simulation_b200/traffic_generator.py) and expects identical agents, field for field, bit for bit.
Maps and parameters: evaluation_lane_change.py:9-44, evaluation_merging.py:8-37, evaluation_exit.py:9-43.
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
import simulation_multilane as sm  # noqa: E402  (pulls in random_traffic_generator, environment, type unchanged)

sys.path.insert(0, ROOT)
from interactive_highway_simulation_b200 import traffic_generator as tg  # noqa: E402  (only for the shared map specs)


def ref_map(spec):
    corridors = []
    for c in spec:
        o = sm.Corridor(c["id"], c["type"], c["left"], c["right"], c["center"], c["v_limit"])
        o.set_left_and_right_corridor_id(c["left_id"], c["right_id"])
        corridors.append(o)
    return sm.Map(corridors)


def main():
    out = []
    for name in ("lane_change", "merging", "exit"):
        for seed in (0, 1, 2):
            spec, params, unsafe, truck = tg.evaluation_map(name, v_limit=100 / 3.6)
            np.random.seed(seed); random.seed(seed)
            gen = sm.RandomTrafficGenerator()
            ref_params = {i: sm.RandomDensityAndVelocityParameter(*p) for i, p in params.items()}
            merging, main_ = gen.random_agents_on_map(ref_map(spec), ref_params, unsafe, truck)
            rec = {"map": name, "seed": seed, "merging": [a.to_dict() for a in merging], "main": [a.to_dict() for a in main_]}
            out.append(rec)
            print(name, seed, len(merging), len(main_))
    with gzip.open(os.path.join(HERE, "generator_scenes.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()

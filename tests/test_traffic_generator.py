"""The package's scene generator against records written by the reference's own random_traffic_generator.py
(tests/golden/make_generator_golden.py): same seeds -> the same agents, field for field, bit for bit."""
import gzip
import json
import os
import random

import numpy as np
import pytest

from interactive_highway_simulation_b200 import traffic_generator as tg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_scenes.json.gz")
with gzip.open(GOLDEN, "rt") as f:
    RECORDS = json.load(f)


@pytest.mark.parametrize("rec", RECORDS, ids=["%s_seed%d" % (r["map"], r["seed"]) for r in RECORDS])
def test_generated_agents_match_the_reference(rec):
    spec, params, unsafe, truck = tg.evaluation_map(rec["map"], v_limit=100 / 3.6)
    np.random.seed(rec["seed"]); random.seed(rec["seed"])
    gen = tg.RandomTrafficGenerator()
    merging, main = gen.random_agents_on_map(tg.lanes_from_spec(spec), {i: tg.RandomDensityAndVelocityParameter(*p) for i, p in params.items()},
                                             unsafe, truck)
    for got, want, what in ((merging, rec["merging"], "merging"), (main, rec["main"], "main")):
        assert len(got) == len(want), what
        for g, w in zip(got, want):
            assert g == w, (what, g["id"], {k: (g[k], w[k]) for k in w if g.get(k) != w[k]})


def test_truck_lane_and_lane_geometry():
    spec, _, _, _ = tg.evaluation_map("lane_change")
    lanes = tg.lanes_from_spec(spec)
    assert tg.truck_lane_id(lanes) == 1                       # the lane left of the merging lane
    spec, _, _, _ = tg.evaluation_map("exit")
    assert tg.truck_lane_id(tg.lanes_from_spec(spec)) == 1    # ... of the exit lane
    plain = {k: v for k, v in lanes.items() if v.type == "main"}
    for v in plain.values():
        v.right_id = None if v.id == 1 else v.right_id
    assert tg.truck_lane_id(plain) == 1                       # no ramp: the right-most lane
    c = lanes[1]
    assert c.path_length == 700.0
    x, y, yaw = c.to_global_coordinate(250.0, 0.5)
    assert (x, y, yaw) == (pytest.approx(50.0, abs=1e-9), 5.75, 0.0)      # the reference interpolates from the far end
    x, y, _ = c.to_global_coordinate(800.0, 0.0)              # past the end: extrapolated along the last segment
    assert (x, y) == (pytest.approx(600.0, abs=1e-9), 5.25)

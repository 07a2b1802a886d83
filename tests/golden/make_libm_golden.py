#!/usr/bin/env python
"""Golden vectors from the STOCK reference — its machine code with the platform libm's exp / pow, exactly as it ships
(no pinned math interposed) — at the shape of every BASELINE.json configuration.  Build container only (oracle/_ref):

    python tests/golden/make_libm_golden.py

The product computes exp / pow with include/mcsim_math.h, the reference as shipped with glibc's; the two differ by at
most 1 ulp per call.  These records are what "trajectories within a stated fp64 tolerance per step, discrete outcomes
bit-exact" (BASELINE.json north_star) is checked against: tests/test_libm_parity.py (oracle, CPU) and
tests/test_gpu_parity.py::test_against_the_stock_reference (CUDA path).  Random draws are the injected counter stream
of include/mcsim_rng.h, as everywhere.  Same record format as make_golden.py's `rollouts` fixtures, kind
`rollouts_libm`, file names libm_*.json.gz.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import make_golden as mg                                          # noqa: E402
import scenes                                                     # noqa: E402
from bench import scene_spec                                      # noqa: E402
from interactive_highway_simulation_b200 import synthetic         # noqa: E402


def main():
    cases = []

    def add(tag, spec, ego, action, gap, dt, steps, min_steps, seed, count, n_hist=1):
        c = mg.rollout_case(tag, spec, ego, action, gap, dt, steps, min_steps, seed, 0, count, n_hist=n_hist, math="libm")
        c["kind"] = "rollouts_libm"
        cases.append(c)

    # configs[0]: the inner scenes the outer test_scene loop builds are 8-25 vehicles on <= 3 lanes (mcs_wrapper's 80 m
    # cut): the probe scene of SURVEY Appendix A.7 and a 21-vehicle merging scene, reference horizon SimParameter(0.3, 40, 30)
    for act in scenes.ACTIONS:
        add("libm_cfg0_probe_%s" % act, scenes.probe_scene(), 0, act, -2, 0.3, 40, 30, 7, 8)
    c, a, ego, gaps = synthetic.merging_scene(21, 2, n_main=2)
    for gi, g in enumerate(gaps[:2]):
        add("libm_cfg0_merge21_gap%d" % gi, scene_spec(c, a), ego, "left", g, 0.3, 40, 10, 8, 8)
    # configs[1]: merging lane + 3 main lanes, 40 vehicles, every candidate gap of the decision bench.py times
    c, a, ego, gaps = synthetic.merging_scene(40, 0)
    for gi, g in enumerate(gaps):
        add("libm_cfg1_merge40_gap%d" % gi, scene_spec(c, a), ego, "left", g, 0.3, 40, 10, 1000, 16)
    # configs[2]: 4 lanes, 100 vehicles, the five lane-change commands
    c, a, ego = synthetic.lane_change_scene(100, 4, seed=1)
    for act in ("keep_lane", "dcc", "acc", "left", "right"):
        add("libm_cfg2_lc100_%s" % act, scene_spec(c, a), ego, act, -2, 0.3, 40, 30, 2, 8)
    # configs[3]: exit lane + 3 main lanes, 200 vehicles, 50-step horizon
    c, a, ego, gaps = synthetic.exit_scene(200, 0)
    for gi, g in enumerate(gaps[:2]):
        add("libm_cfg3_exit200_gap%d" % gi, scene_spec(c, a), ego, "right", g, 0.3, 50, 10, 3, 6)
    # configs[4]: 256 vehicles x 100 steps, ego keep_lane, minSteps = steps (the throughput workload)
    c, a, ego = synthetic.lane_change_scene(256, 4, seed=0)
    add("libm_cfg4_lc256", scene_spec(c, a), ego, "keep_lane", -2, 0.3, 100, 100, 2026, 4)
    for c in cases:
        with gzip.open(os.path.join(HERE, c["tag"] + ".json.gz"), "wt") as f:
            json.dump(c, f, separators=(",", ":"))
        print(c["tag"], [(r.get("ok"), r.get("finishStep"), r.get("draws")) for r in c["rollouts"]])
    print("wrote %d fixtures" % len(cases))


if __name__ == "__main__":
    main()

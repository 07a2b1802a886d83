#!/usr/bin/env python
"""Known-answer vectors for the learned decision models, produced by IMPORTING THE REFERENCE'S learned_model.py
(pure numpy; runs in the build container where /root/reference exists):  python tests/golden/make_learned_model_golden.py
Inputs: the three feature matrices of the reference's own __main__ (learned_model.py:79-99) plus seeded random ones."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import learned_model as ref  # noqa: E402


def main():
    rng = random.Random(11)
    cases = []
    lc = ref.LearnedLaneChangeModel()
    mains = [
        (["left", "keep_lane", "dcc", "acc"],
         [[1.0, 0.0, 0.8347861709085435, 0.6743563285927423, 0.24833333333333338, 0.8191169496505016, 0.9227193316538841],
          [1.0, 0.0, 0.8853556042443036, 0.8439449702745763, 0.09999999999999999, 0.6586241190819563, 0.6891999923860896],
          [1.0, 0.0, 0.6976766932260383, 0.63619700339087, 0.09999999999999999, 0.8208634255022598, 0.9273373686498663],
          [1.0, 0.0, 0.7808685861030975, 0.9227713968177605, 0.09999999999999999, 0.6802358091597496, 0.7512569752282063]]),
        (["keep_lane", "acc", "dcc"],
         [[1.0, 0.0, 0.9289858483420468, 0.8855733633991669, 0.1, 1.0, 1.0],
          [1.0, 0.0, 0.8842244456976619, 0.9954408577571839, 0.1, 1.0, 1.0],
          [1.0, 0.0, 0.861946719555924, 0.7104609461924002, 0.1, 1.0, 1.0]]),
        (["left", "keep_lane", "dcc", "acc"],
         [[1.0, 0.0, 0.8619296717349689, 0.9142026620520505, 0.125, 0.9982689235522592, 0.9692335473951416],
          [1.0, 0.0, 0.7196382886956361, 0.9370796297108704, 0.09999999999999999, 0.9979878665033657, 0.969206543425642],
          [1.0, 0.0, 0.6978860317014086, 0.8001376554275782, 0.09999999999999999, 0.9981125958871385, 0.9695845002591762],
          [1.0, 0.0, 0.5856435906807819, 0.9828949735363534, 0.09999999999999999, 0.9983211486397954, 0.9697824316858275]]),
    ]
    for commands, feats in mains:
        q, d = lc.inference(feats, commands)
        cases.append({"model": "lane_change", "commands": commands, "features": feats, "last": "initial",
                      "q": [[c, float(v)] for c, v in q], "decision": d})
    all_cmds = ["keep_lane", "dcc", "acc", "left", "right"]
    for _ in range(40):
        commands = all_cmds[:rng.choice([3, 4, 5])]
        feats = [[rng.random(), rng.random() * 0.3] + [rng.uniform(0.3, 1.0) for _ in range(5)] for _ in commands]
        last = rng.choice(["initial"] + commands)
        q, d = lc.inference(feats, commands, last)
        cases.append({"model": "lane_change", "commands": commands, "features": feats, "last": last,
                      "q": [[c, float(v)] for c, v in q], "decision": d})
    mm = ref.LearnedMergingModel()
    for _ in range(40):
        n = rng.choice([1, 2, 3, 4])
        feats = [[rng.choice([0.0, 0.1, 0.5, 1.0]) * rng.random(), rng.random()] + [rng.uniform(0.0, 1.0) for _ in range(5)]
                 for _ in range(n)]
        idx, q = mm.inference(feats)
        cases.append({"model": "merging", "features": feats, "index": idx, "q": {str(k): float(v) for k, v in q.items()}})
    with open(os.path.join(HERE, "learned_model.json"), "w") as f:
        json.dump({"tag": "learned_model", "kind": "learned_model", "cases": cases}, f, separators=(",", ":"))
    print("wrote %d learned-model cases" % len(cases))


if __name__ == "__main__":
    main()

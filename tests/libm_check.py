"""Comparison of a rollout result with the STOCK reference (platform libm; tests/golden/libm_*.json.gz): discrete outcomes
must be identical, continuous ones within the stated per-step tolerance.

Discrete = rollout accepted (ok), finish, fallBack, finishStep, number of random draws consumed (every MOBIL decision and
yielding intention that fires consumes one, so an extra or missing decision shows here), number of recorded states, and
for every vehicle at every step the lane whose lateral interval holds it and the direction of its lateral motion (a
lane-change / merging decision taken differently moves a vehicle by 0.3 m per step — far outside any tolerance).
Continuous = x, y, v, a of every vehicle at every step, and the four outcome statistics."""
import glob
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# Stated tolerance of the product against the reference as shipped: absolute, per vehicle and step.
# Measured on the committed records: 3.9e-16 (states), 1.2e-16 (statistics) — one or two units in the last place.
STATE_TOL = 1e-12       # m, m/s, m/s^2
STAT_TOL = 1e-12        # utility / comfort statistics (dimensionless, O(1))
DISCRETE_KEYS = ("ok", "finish", "fallBack", "finishStep", "draws")
STAT_KEYS = ("utility", "comfort", "utilityObjects", "comfortObjects")


def load_cases():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "libm_*.json.gz"))):
        with gzip.open(p, "rt") as f:
            out.append(json.load(f))
    return out


def lane_of(corridors, y):
    for c in corridors:
        if c["left_y"] >= y >= c["right_y"]:
            return c["id"]
    return None


def lateral_direction(y0, y1):
    d = y1 - y0
    return 0 if abs(d) < 1e-6 else (1 if d > 0 else -1)


def compare(case, statuses, histories):
    """statuses: list of dicts (one per rollout of the case); histories: {rollout index: {agent id: [(x, y, v, a), ...]}}.
    Returns (discrete_mismatches, max_state_deviation, max_stat_deviation)."""
    mismatches = []
    dev_state = dev_stat = 0.0
    corridors = case["scene"]["corridors"]
    for i, want in enumerate(case["rollouts"]):
        got = statuses[i]
        for k in DISCRETE_KEYS:
            if int(got[k]) != int(want[k]):
                mismatches.append((case["tag"], "rollout", i, k, got[k], want[k]))
        for k in STAT_KEYS:
            dev_stat = max(dev_stat, abs(got[k] - want[k]))
        if "hist" not in want or i not in histories:
            continue
        for aid, hw in want["hist"].items():
            hg = histories[i][int(aid)]
            if len(hg) != len(hw):
                mismatches.append((case["tag"], "rollout", i, "agent", aid, "states", len(hg), len(hw)))
                continue
            for t, (g, w) in enumerate(zip(hg, hw)):
                dev_state = max(dev_state, max(abs(a - b) for a, b in zip(g, w)))
                if lane_of(corridors, g[1]) != lane_of(corridors, w[1]):
                    mismatches.append((case["tag"], "rollout", i, "agent", aid, "t", t, "lane"))
                if t and lateral_direction(hg[t - 1][1], g[1]) != lateral_direction(hw[t - 1][1], w[1]):
                    mismatches.append((case["tag"], "rollout", i, "agent", aid, "t", t, "lateral direction"))
    return mismatches, dev_state, dev_stat

"""Host-side logic of libmcsim_b200.so that needs no GPU: the fixed-order two-level feature reduction
(runMTimes so:0xc2a6e-0xc2b1f + runMultiThreads so:0xbca08-0xbca92) against the oracle's restatement."""
import ctypes as C
import random

from interactive_highway_simulation_b200 import _abi as abi


def random_statuses(rng, n, ok_groups, m):
    st = (abi.Status * n)()
    for i, s in enumerate(st):
        s.utility = rng.uniform(0.3, 1.1); s.comfort = rng.uniform(0.2, 1.0)
        s.utilityObjects = rng.uniform(0.5, 1.0); s.comfortObjects = rng.uniform(0.5, 1.0)
        s.finish = rng.random() < 0.7; s.fallBack = rng.random() < 0.2
        s.finishStep = rng.randrange(3, 40); s.ok = 1 if ok_groups[i // m] else 0
    return st


def test_reduce_features_matches_oracle(oracle, backend_lib):
    rng = random.Random(4)
    for trial in range(20):
        groups, m, steps = 8, rng.choice([1, 3, 20, 30]), rng.choice([25, 40, 100])
        ok_groups = [rng.random() < 0.9 for _ in range(groups)]
        if trial == 0:
            ok_groups = [True] * groups
        st = random_statuses(rng, groups * m, ok_groups, m)
        got = backend_lib.reduce_features(st, 1, groups, m, steps)[0]
        want = oracle.reduce(st, groups, m, steps)
        for f, _ in abi.Feature._fields_:
            a, b = getattr(got, f), getattr(want, f)
            assert a == b or (a != a and b != b), (trial, f, a, b)


def test_reduce_features_is_mean_of_group_means(backend_lib):
    rng = random.Random(9)
    groups, m = 8, 5
    st = random_statuses(rng, groups * m, [True] * groups, m)
    f = backend_lib.reduce_features(st, 1, groups, m, 40)[0]
    acc = 0.0
    for g in range(groups):
        u = 0.0
        for i in range(m):
            u = u + st[g * m + i].utility
        acc = acc + u / m
    assert f.utility == acc / groups
    assert f.fromNSims == groups * m


def test_reduce_features_rejects_bad_arguments(backend_lib):
    st = (abi.Status * 8)()
    feats = (abi.Feature * 1)()
    assert backend_lib.lib().mcs_reduce_features(st, 1, 0, 1, 40, feats) == abi.MCS_ERR_ARG
    assert backend_lib.lib().mcs_reduce_features(None, 1, 8, 1, 40, feats) == abi.MCS_ERR_ARG


# ---- scene preparation on the host (MapMultiLane ctor so:0x109f00, matchAgentsOnCorridor so:0x10b810) ----
import math

import pytest

import golden_io
import scenes

ORDER_CASES = golden_io.load_all("order")
NONE = -2 ** 31


@pytest.mark.parametrize("case", ORDER_CASES, ids=golden_io.ids(ORDER_CASES))
def test_prepare_host_orders_match_reference_binary(backend_lib, case):
    """Vehicle visit order and corridor iteration order, as recorded from the reference binary's unordered_maps."""
    s = golden_io.scene_of(case)
    got = backend_lib.prepare_host(s.c_corridors(), s.c_agents())
    assert [s.agents[i]["id"] for i in got["agent_order"]] == case["agents"]
    assert [s.corridors[i]["id"] for i in got["corridor_order"]] == case["corridors"]


def test_prepare_host_matches_oracle_on_random_scenes(oracle, backend_lib):
    """Orders and neighbour lanes against the oracle (sizes cross several unordered_map rehash thresholds, ids shuffled and
    sparse); first lane match and lane-change prior against their definition."""
    rng = random.Random(77)
    for n_agents in (1, 2, 5, 11, 12, 13, 14, 29, 30, 31, 59, 60, 61, 62, 97, 127, 128, 200, 256):
        n_lanes = rng.choice([1, 2, 3, 4, 6, 8])
        s, _ = scenes.random_scene(rng, n_lanes=n_lanes, n_agents=n_agents)
        for a in s.agents:                                   # sparse / large ids exercise the bucket arithmetic
            a["id"] = a["id"] * rng.choice([1, 7, 13]) + rng.choice([0, 1000, 100003])
        if len({a["id"] for a in s.agents}) != n_agents:
            continue
        got = backend_lib.prepare_host(s.c_corridors(), s.c_agents())
        ao, co, li, ri = oracle.orders(s)
        assert [s.agents[i]["id"] for i in got["agent_order"]] == ao
        assert [s.corridors[i]["id"] for i in got["corridor_order"]] == co
        assert got["left_ids"] == li and got["right_ids"] == ri
        cs = s.c_corridors()
        box = {c.id: (max(c.left[1], c.left[3]), min(c.right[1], c.right[3]), c.center[1]) for c in cs}
        for i, a in enumerate(s.agents):
            first = next((cid for cid in co if box[cid][0] >= a["y"] >= box[cid][1]), NONE)
            assert got["first_lane_ids"][i] == first, (n_agents, i)
            if first == NONE:
                continue
            d = ((a["y"] - box[first][2]) * 0.333 + 0.6805 * a.get("vy", 0.0)) - 0.01
            g = lambda t, w: math.exp(-t * t / w)
            pl = 0.8 * g(d - 1.04, 0.055) + 0.8 * g(d - 0.35, 0.055) + g(d - 0.7, 0.07) * 0.65
            pr = 0.75 * g(1.04 + d, 0.055) + 0.75 * g(0.35 + d, 0.055) + g(0.7 + d, 0.07) * 0.6
            pk = g(d, 0.055) * 2.5
            want = [pl / (pl + pk + pr), pk / (pl + pk + pr), pr / (pl + pk + pr)]
            assert got["priors"][i] == pytest.approx(want, rel=1e-12, abs=1e-300)
            assert sum(got["priors"][i]) == pytest.approx(1.0, abs=1e-15)


def test_prepare_host_rejects_capacity(backend_lib):
    rng = random.Random(3)
    s, _ = scenes.random_scene(rng, n_lanes=2, n_agents=4)
    cs, ag = s.c_corridors(), s.c_agents()
    big = (abi.Agent * 257)(*[ag[0]] * 257)
    for k in range(257):
        big[k].id = k
    with pytest.raises(Exception):
        backend_lib.prepare_host(cs, big)

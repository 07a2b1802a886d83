"""Host-side logic of libmcsim_b200.so that needs no GPU: the fixed-order two-level feature reduction
(runMTimes so:0xc2a6e-0xc2b1f + runMultiThreads so:0xbca08-0xbca92) against the oracle's restatement."""
import ctypes as C
import os
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


def test_a_marked_rollout_anywhere_fails_the_reduction(backend_lib):
    """mcs_status.overflow of ANY rollout (not only a candidate's first) travels into its group's partial and makes the
    combine fail on whichever rank runs it (the mark used to be read from rollout 0 only)."""
    rng = random.Random(2)
    groups, m = 8, 4
    for bad, code in ((17, abi.MCS_ERR_SCENE), (31, abi.MCS_ERR_CAPACITY)):
        st = random_statuses(rng, groups * m, [True] * groups, m)
        st[bad].overflow = 2 if code == abi.MCS_ERR_SCENE else 1
        feats = (abi.Feature * 1)()
        assert backend_lib.lib().mcs_reduce_features(st, 1, groups, m, 40, feats) == code
        parts = backend_lib.group_partials(st, 1, groups, m, 40)
        assert [p.reserved for p in parts] == [st[bad].overflow if g == bad // m else 0 for g in range(groups)]
        assert backend_lib.lib().mcs_combine_groups(parts, 1, groups, feats) == code
        st[bad].overflow = 0
        assert backend_lib.lib().mcs_reduce_features(st, 1, groups, m, 40, feats) == abi.MCS_OK


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


def test_utility_scalars_equal_the_reference_formulas():
    """utility.py mirrors: hand-evaluated IDM / RSS values (utility.py:23-80,129-141) and the line helpers"""
    import math
    import random
    from types import SimpleNamespace as NS
    from interactive_highway_simulation_b200 import utility as u
    p = NS(acc_max=1.4, acc_min=-4.0, min_dis=2.0, acc_com=-2.0, thw=1.5, vd=30.0)
    f = NS(v=8.0, dis=20.0)
    b = NS(v=12.0, dis=-15.0)
    free = 1.4 * (1 - (10.0 / 27.5) ** 4)
    dd = 0.7 * (2.0 + 10.0 * 1.5) + 10.0 * (10.0 - 8.0) / (2 * math.sqrt(2.8))
    want = free - 1.4 * (dd / 20.0) ** 2
    assert u.idm_acc_f_b([f], None, 10.0, 25.0, p) == max(min(1.4, want), -4.0)
    db = 0.7 * (2.0 + 12.0 * 1.5) + 12.0 * (12.0 - 10.0) / (2 * math.sqrt(2.8))
    assert u.idm_acc_f_b([f], b, 10.0, 25.0, p) == max(min(1.4, want + 1.4 * (db / 15.0) ** 2), -4.0)
    assert u.idm_acc_fs_bs([f], [b], 10.0, 25.0, p) == u.idm_acc_f_b([f], b, 10.0, 25.0, p)
    assert u.idm_acc_f_b([], None, 10.0, 25.0, p, obey_v_limit=False) == 1.4 * (1 - (10.0 / 30.0) ** 4)
    assert u.rss_dis(10.0, 12.0, 0.4, -8.0, -10.0) == 12.0 * 0.4 + 144.0 / 16.0 - 100.0 / 20.0
    assert u.rss_dis(30.0, 1.0, 0.4, -8.0, -10.0) == 0.0
    assert u.thw_and_ttc(10.0, 0.0, 5.0) == (10.0 / 1e-10, -5.0 / 1e-3)
    q = u.idm_p_for_merging_exit_agents(17, p, False)
    random.seed(17)
    r = random.random()
    assert (q.acc_max, q.acc_min, p.acc_max) == (1.4 - (0.6 + 0.4 * r), -4.0 + (0.8 + 0.4 * r), 1.4)
    assert u.idm_p_for_merging_exit_agents(17, p, True).acc_max == 1.4
    assert u.extend_line([[0, 0], [3, 4]], 5)[-1] == [6.0, 8.0]
    assert u.extend_line_both_sides([[0, 0], [3, 4]], 5) == [[-3.0, -4.0], [0, 0], [3, 4], [6.0, 8.0]]
    assert u.vehicle_pose_to_rect(1.0, 2.0, 0.0, 4.0, 2.0) == [[3.0, 3.0], [3.0, 1.0], [-1.0, 1.0], [-1.0, 3.0]]


def test_yielding_intention_follows_the_global_random_stream():
    """type.py:70-90: a truthy seed re-seeds `random`; seed 0 / None draws from the running stream"""
    import random
    from interactive_highway_simulation_b200.type import YieldingIntention
    random.seed(5)
    want = random.random()
    a = YieldingIntention(3, 0.5, random_intention_seed=5)
    assert a.yielding == (want <= 0.5)
    random.seed(9)
    first = random.random()
    random.seed(9)
    b = YieldingIntention(3, first, random_intention_seed=0)        # not re-seeded: consumes the first draw of the stream
    assert b.yielding
    y = YieldingIntention(1, 0.5, random_intention_seed=5)
    y.yielding = True
    y.update_intention(0.3126)
    assert y.yielding
    y.update_intention(0.3125)                                      # <= threshold * initial probability: gives up yielding
    assert not y.yielding
    y.update_intention(0.6874)
    assert not y.yielding
    y.update_intention(0.6875)                                      # (1 - p) <= threshold * (1 - initial): yields again
    assert y.yielding


def test_corridor_host_values_and_statistics_trimming():
    """the cheap host values of the reference's Corridor (type.py:157-169, 208-215) and remove_invalid_data
    (agent.py:325-331) as user scripts touch them"""
    from interactive_highway_simulation_b200 import agent, outer_env as oe
    c = oe.Corridor(1, "main", [[0, 3.5], [50, 3.5], [100, 4.5]], [[0, 0], [100, 1.0]], [[0, 1.75], [100, 2.75]], 30)
    assert c.border == [[0, 3.5], [50, 3.5], [100, 4.5], [100, 1.0], [0, 0]]
    assert (c.x_min, c.x_max, c.y_min, c.y_max) == (0, 100, 0, 4.5) == c.get_x_y_limit()
    assert abs(c.width - 3.0 * 100 / math.hypot(100, 1)) < 1e-12            # left's middle vertex (50, 3.5) against the slanted right border
    assert c.contains([10, 1]) and not c.contains([10, 3.5]) and not c.contains([10, 5]) and not c.contains([-1, 1])
    st = agent.MergingAndExitStatistics()
    st.utility_ego, st.comfort_ego, st.utility_other, st.comfort_other = [1, 2, 3, 4], [5, 6, 7, 8], [1, 1, 1, 1], [2, 2, 2, 2]
    st.total_t, st.epoch_and_ego_id_history = [9, 9, 9, 9], [[0, 1], [0, 2], [0, 3], [0, 4]]
    st.remove_invalid_data([1])
    # the reference ranges every later list over the ALREADY shortened utility_ego: one sample removed, one more cut off the end
    assert st.utility_ego == [1, 3, 4] and st.comfort_ego == [5, 7] and st.epoch_and_ego_id_history == [[0, 1], [0, 3]]


def test_hash_iteration_order_restatement_equals_the_container(backend_lib):
    """csrc/walk_kernel.cuh restates the list / bucket discipline of libstdc++'s std::unordered_map<int, T> for the device
    (visit order of the inner simulator's agents and corridors); scene.cpp uses the container itself.  Same order for random
    id sets of every size up to 64 (bucket growth 13 -> 29 -> 59 -> 127), colliding and negative ids included."""
    from interactive_highway_simulation_b200 import backend
    lib = backend_lib.lib()
    lib.mce_hash_iteration_order.restype = C.c_int
    rng = random.Random(11)
    cs = (abi.Corridor * 1)()
    cs[0].id = 1; cs[0].type = 0
    cs[0].left[:] = [-1000, 3.5, 5000, 3.5]; cs[0].right[:] = [-1000, 0, 5000, 0]; cs[0].center[:] = [-1000, 1.75, 5000, 1.75]
    for n in list(range(1, 65)) + [13, 14, 29, 30, 59, 60]:
        for pool in (range(0, 70), range(0, 400), range(-200, 2000), [13 * k for k in range(200)]):
            ids = rng.sample(list(pool), n)
            arr = (C.c_int32 * n)(*ids)
            got = (C.c_int32 * n)()
            assert lib.mce_hash_iteration_order(arr, n, got) == 0
            agents = (abi.Agent * n)()
            for k, a in enumerate(agents):
                a.id = ids[k]; a.x = 10.0 * k; a.y = 1.0; a.l = 4.8; a.w = 2.0; a.v = 10.0; a.vd = 12.0
                a.idm[:] = (1.6, 2.0, -3.0, -2.0, -8.0, 1.2); a.rss[:] = (0.7, 0.4, -6.0, -6.0, 1.8, 1.2)
            want = backend.prepare_host(cs, agents)["agent_order"]
            assert list(got) == want, (n, ids)


def test_unit_draw_without_division(tmp_path):
    """mcs_unit31 (csrc/rollout_kernel.cuh) turns a 31-bit draw into r / 2147483647.0 with one multiplication and two fused
    multiply-adds instead of the division sequence: identical to the IEEE quotient for EVERY r in [0, 2^31) — all of them
    are tried here (tests/unit31_check.c, ~3 s with hardware FMA)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unit31_check.c")
    exe = str(tmp_path / "unit31_check")
    with open("/proc/cpuinfo") as f:
        has_fma = " fma " in f.read()
    if not has_fma:
        pytest.skip("no hardware FMA on this host: the exhaustive loop would run on libm's software fma for minutes")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout.strip()
    assert out == "0", out

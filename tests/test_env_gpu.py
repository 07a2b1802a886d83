"""The outer-environment kernels (include/mcsim_env.h) through the C ABI: bit-identical to oracle/env_oracle.py on the
golden states and on random batched scenes over bent polylines, and equal to the records of the reference's own
environment.py / mcs_wrapper.py (ids exact, coordinates within 1e-9)."""
import math
import random

import numpy as np
import pytest

from env_cases import BORDER, STATES, TOL, DuckAgent, corridor_dicts, prepared, slots
from oracle import env_oracle as eo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oe(backend_lib):
    from interactive_highway_simulation_b200 import outer_env
    return outer_env


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | ((a != a) & (b != b))))


def oracle_tables(pm, xs, ys):
    lane, arc, order, start = eo.match(pm, xs, ys)
    n = len(xs)
    idx, dis = np.full((10, n), -1, np.int32), np.zeros((10, n))
    for i in range(n):
        a, b = eo.neighbors(pm, lane, arc, order, start, i, xs[i], ys[i])
        idx[:, i], dis[:, i] = a, b
    return lane, arc, order, start, idx, dis


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_match_tables_bit_identical_to_the_oracle(oe, rec):
    pm, ids, xs, ys = prepared(rec)
    dm = oe.DeviceMap(rec["corridors"])
    t = dm.match(xs, ys, 1, arc_on_lanes=True)
    lane, arc, order, start, idx, dis = oracle_tables(pm, xs, ys)
    assert same(t.lane[0], lane) and same(t.arc[0], arc) and same(t.order[0], order) and same(t.start[0], start)
    assert same(t.nb_index[0], idx) and same(t.nb_dis[0], dis)
    assert same(t.arc_on_lane[0], eo.arc_on_lanes(pm, xs, ys))
    # moved vehicles against the stale vectors
    slot = slots(ids)
    q = [slot[m["id"]] for m in rec["moved"] if m["id"] in slot]
    qx, qy = [m["x"] for m in rec["moved"]], [m["y"] for m in rec["moved"]]
    gi, gd = dm.neighbors([0] * len(q), q, qx, qy)
    for k, i in enumerate(q):
        a, b = eo.neighbors(pm, lane, arc, order, start, i, qx[k], qy[k])
        assert same(gi[:, k], a) and same(gd[:, k], b)
    # Frenet, centre projection, step
    for c in range(pm.n):
        s, d = dm.frenet(c, xs, ys)
        want = [pm.ref[c].to_frenet_frame(x, y) for x, y in zip(xs, ys)]
        assert same(s, [w[0] for w in want]) and same(d, [w[1] for w in want])
        s, d = dm.project_center(c, xs, ys)
        want = [eo.project_and_distance(pm.center[c], x, y) for x, y in zip(xs, ys)]
        assert same(s, [w[0] for w in want]) and same(d, [w[1] for w in want])
    cpos = {c["id"]: k for k, c in enumerate(rec["corridors"])}
    by_id = {a["id"]: a for a in rec["agents"]}
    st = rec["steps"]
    for dt in sorted({s["dt"] for s in st}):
        sel = [s for s in st if s["dt"] == dt]
        ag = [by_id[s["id"]] for s in sel]
        out = dm.step_along_path([cpos[s["lane"]] for s in sel], [a["x"] for a in ag], [a["y"] for a in ag], [a["v"] for a in ag],
                                 [s["ax"] for s in sel], [s["vy"] for s in sel], dt)
        for k, (s, a) in enumerate(zip(sel, ag)):
            want = eo.step_along_path(pm, cpos[s["lane"]], a["x"], a["y"], a["v"], s["ax"], s["vy"], dt)
            assert tuple(float(o[k]) for o in out) == want
            assert all(abs(float(o[k]) - w) <= TOL for o, w in zip(out, s["out"]))      # the reference's own answer
    dm.close()


def build_env(oe, rec):
    cs = []
    for c in rec["corridors"]:
        k = oe.Corridor(c["id"], c["type"], c["left"], c["right"], c["center"], c["v_limit"])
        k.set_left_and_right_corridor_id(c["left_id"], c["right_id"])
        cs.append(k)
    return oe.Environment(oe.Map(cs), [DuckAgent(a) for a in rec["agents"]])


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_environment_mirror_answers_like_the_reference(oe, rec):
    env = build_env(oe, rec)
    for a in rec["agents"]:
        want = rec["lane"][str(a["id"])]
        got = env.matched_agents[a["id"]].corridor_id if a["id"] in env.matched_agents else None
        assert got == want
        assert (a["id"] in env.removed_agents) == (want is None)
    for c, want in rec["lane_vectors"].items():
        got = env.matched_agents_sorted_by_arc_length[int(c)]
        assert [g[1].id for g in got] == [w[1] for w in want]
        assert all(abs(g[0] - w[0]) <= TOL for g, w in zip(got, want))
    names = ("left_ff", "left_f", "left_b", "left_bb", "f", "b", "right_ff", "right_f", "right_b", "right_bb")

    def check(table):
        for a, want in table.items():
            nb = env.neighbor_around_agent(int(a))
            for k, name in enumerate(names):
                g = getattr(nb, name)
                if want[k] is None:
                    assert g is None
                else:
                    assert g is not None and g.id == want[k][0] and abs(g.dis - want[k][1]) <= TOL
            f, b = env.front_agent(int(a)), env.following_agent(int(a))
            assert (f.id if f else None) == (want[4][0] if want[4] else None)
            assert (b.id if b else None) == (want[5][0] if want[5] else None)

    check(rec["neighbors"])
    for w in rec["within_range"]:
        got = env.sorted_agents_on_corridor_within_range_of_vehicle(w["corridor"], w["ego"], w["max_dis"])
        assert [a.id for a in got] == w["ids"]
    for m in rec["moved"]:
        env.agents[m["id"]].x, env.agents[m["id"]].y = m["x"], m["y"]
    check(rec["neighbors_after_move"])


@pytest.mark.parametrize("rec", [r for r in STATES if r["builders"]], ids=[r["tag"] for r in STATES if r["builders"]])
def test_marshalling_builders_equal_the_reference(oe, rec):
    """mcs_wrapper.py's three builders: same corridors, vehicles (order, ids, behaviours, parameters, vd noise) and gap
    candidates as the reference's own functions on the same environment and numpy seed."""
    from interactive_highway_simulation_b200 import mcs_wrapper as mw
    env = build_env(oe, rec)
    for b in rec["builders"]:
        np.random.seed(b["np_seed"])
        ego = env.agents[b["ego"]]
        got = getattr(mw, b["builder"])(env, ego, env.corridor_for_agent(ego.id))
        corridors, agents = got[0], got[1]
        assert (list(got[2]) if len(got) > 2 else None) == b["possible_actions"], b["builder"]
        assert len(corridors) == len(b["corridors"]) and len(agents) == len(b["agents"]), (b["builder"], b["ego"])
        for c, w in zip(corridors, b["corridors"]):
            g = [c.id, c.type, c.left.p1.x, c.left.p1.y, c.left.p2.x, c.left.p2.y, c.right.p1.x, c.right.p1.y, c.right.p2.x,
                 c.right.p2.y, c.center.p1.x, c.center.p1.y, c.center.p2.x, c.center.p2.y, c.vLimit]
            assert g[:2] == w[:2]
            assert all(abs(x - y) <= TOL for x, y in zip(g[2:], w[2:])), (b["builder"], b["ego"], g, w)
        for a, w in zip(agents, b["agents"]):
            g = [a.id, a.x, a.y, a.vx, a.vy, a.vd, a.a, a.l, a.w] + list(a.idm._memory_order()) + list(a.rss._memory_order()) + \
                list(a.yp._memory_order()) + [a.behavior]
            assert g[0] == w[0] and g[-1] == w[-1], (b["builder"], b["ego"], g[0], w[0])
            assert all(abs(x - y) <= TOL for x, y in zip(g[1:-1], w[1:-1])), (b["builder"], b["ego"], g, w)


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_device_builder_bit_identical_to_the_oracle(oe, rec):
    """mce_build_rollout_scene against oracle build_scene (itself pinned on the reference's builders, test_env_oracle.py):
    all three builders for every vehicle that has the lanes the builder needs, at the matched positions and again after some
    vehicles have moved (stale lane vectors, current positions) — every field of every record identical."""
    from interactive_highway_simulation_b200 import _abi as abi
    from interactive_highway_simulation_b200 import outer_env
    pm, ids, xs, ys = prepared(rec)
    tables = eo.match(pm, xs, ys)
    slot = slots(ids)
    cds = corridor_dicts(rec)
    info = [{"id": c["id"], "type": abi.LANE_TYPES.get(c["type"], abi.LANE_OTHER), "v_limit": c["v_limit"]} for c in rec["corridors"]]
    dm = oe.DeviceMap(rec["corridors"])
    dm.set_lane_info([c["id"] for c in info], [c["type"] for c in info], [c["v_limit"] for c in info])
    dm.match(xs, ys, 1)
    agents = rec["agents"]
    n = len(agents)
    rng = random.Random(5)
    checked = 0
    for moved in (False, True):
        qx, qy = list(xs), list(ys)
        if moved:
            for m in rec["moved"]:
                if m["id"] in slot:
                    qx[slot[m["id"]]], qy[slot[m["id"]]] = m["x"], m["y"]
        attrs = np.array([qx, qy, [a["v"] for a in agents], [a["vy"] for a in agents], [a["a"] for a in agents],
                          [a["l"] for a in agents], [a["w"] for a in agents], [a["idm"][0] for a in agents]], np.float64)
        per = [dict(v=a["v"], vy=a["vy"], a=a["a"], l=a["l"], w=a["w"], vd=a["idm"][0]) for a in agents]
        idarr = np.array(ids, np.int32)
        for e in range(n):
            lane = tables[0][e]
            if lane < 0:
                continue
            for kind in (0, 1, 2):
                if kind == 0 and pm.left[lane] < 0:
                    continue
                if kind == 1:
                    r, ex = pm.right[lane], -1
                    while r >= 0 and ex < 0:
                        ex, r = (r if info[r]["type"] == abi.LANE_TYPES["exit"] else -1), pm.right[r]
                    if ex < 0:
                        continue
                a = agents[e]
                i, s_, m_, y = a["idm"], a["rss"], a["mobil"], a["yielding"]
                rq = outer_env.BuildRequest()
                rq.kind, rq.scene, rq.ego_slot = kind, 0, e
                rq.ego_behavior = rng.choice([0, 2, 3])
                rq.others_merging = rng.choice([0, 1])
                rq.ego_commit, rq.ego_keep_step, rq.ego_target_lane = rng.choice([0, 1, 2, 3]), rng.randrange(12), rng.choice([-1, 1, 2])
                rq.ego_vd = i[0]
                rq.ego_idm[:] = (i[1], i[3], i[2], i[4], -8.0, i[5])
                rq.ego_rss[:] = tuple(s_)
                rq.ego_yp[:] = (y[0], y[1], y[2], y[3], m_[0], m_[1], m_[2])
                noise = np.array([rng.gauss(0, 1) for _ in range(48)])
                cs, nc, ag, na, actions, _ = dm.build_scene(rq, idarr, attrs, noise, records=True, scene=False)
                ego = dict(behavior=rq.ego_behavior, vd=rq.ego_vd, idm=tuple(rq.ego_idm), rss=tuple(rq.ego_rss), yp=tuple(rq.ego_yp),
                           commit=rq.ego_commit, keep=rq.ego_keep_step, target=rq.ego_target_lane)
                wc, wa, wact = eo.build_scene(pm, info, tables, qx, qy, per, ids, kind, e, ego, bool(rq.others_merging), list(noise))
                assert actions == wact and nc == len(wc) and na == len(wa), (kind, e, actions, wact, nc, na)
                for c, w in zip(cs[:nc], wc):
                    assert (c.id, c.type, list(c.left), list(c.right), list(c.center), c.v_limit) == (w[0], w[1], w[2], w[3], w[4], w[5]), (kind, e)
                for g, w in zip(ag[:na], wa):
                    got = (g.id, g.behavior, g.x, g.y, g.v, g.vy, g.vd, g.a, g.l, g.w, tuple(g.idm), tuple(g.rss), tuple(g.yp), g.commit,
                           g.commit_keep_lane_step, g.target_lane_id)
                    want = (w["id"], w["behavior"], w["x"], w["y"], w["v"], w["vy"], w["vd"], w["a"], w["l"], w["w"], w["idm"], w["rss"],
                            w["yp"], w["commit"], w["keep"], w["target"])
                    assert got == want, (kind, e, got, want)
                checked += 1
    assert checked > 0
    dm.close()


def bent_map():
    """three lanes following a bent centre line (4-point polylines), lane 0 right-most"""
    base = [(0.0, 0.0), (300.0, 0.0), (600.0, 40.0), (1000.0, 40.0)]

    def offset(d):
        pts = []
        for k, (x, y) in enumerate(base):
            a, b = base[max(k - 1, 0)], base[min(k + 1, len(base) - 1)]
            yaw = math.atan2(b[1] - a[1], b[0] - a[0])
            pts.append([x - d * math.sin(yaw), y + d * math.cos(yaw)])
        return pts

    cs = []
    for k in range(3):
        cs.append({"id": 10 + k, "type": "main", "right": offset(3.5 * k), "left": offset(3.5 * (k + 1)), "center": offset(3.5 * k + 1.75),
                   "v_limit": 30.0, "left_id": 11 + k if k < 2 else None, "right_id": 9 + k if k > 0 else None})
    return cs


def test_batched_random_scenes_on_bent_lanes(oe):
    cs = bent_map()
    pos = {c["id"]: k for k, c in enumerate(cs)}
    cd = [dict(c, left_index=pos[c["left_id"]] if c["left_id"] is not None else -1,
               right_index=pos[c["right_id"]] if c["right_id"] is not None else -1) for c in cs]
    pm = eo.PreparedMap(cd)
    dm = oe.DeviceMap(cs)
    rng = random.Random(9)
    n_scenes, n = 24, 300
    xs, ys = np.empty((n_scenes, n)), np.empty((n_scenes, n))
    for s in range(n_scenes):
        for i in range(n):
            xs[s, i] = rng.uniform(-30, 1030) if rng.random() > 0.05 else float("nan")
            ys[s, i] = rng.uniform(-3, 54)
            if rng.random() < 0.1:
                xs[s, i] = float(round(xs[s, i] / 50) * 50) if xs[s, i] == xs[s, i] else xs[s, i]      # ties
    t = dm.match(xs, ys, n_scenes, arc_on_lanes=True)
    for s in range(n_scenes):
        lane, arc, order, start, idx, dis = oracle_tables(pm, list(xs[s]), list(ys[s]))
        assert same(t.lane[s], lane) and same(t.arc[s], arc) and same(t.order[s], order) and same(t.start[s], start), s
        assert same(t.nb_index[s], idx) and same(t.nb_dis[s], dis), s
        # properties: every lane vector is sorted, the vectors partition the matched slots
        members = [j for j in t.order[s] if j >= 0]
        assert sorted(members) == [i for i in range(n) if lane[i] >= 0]
        for c in range(3):
            keys = [t.arc[s, j] for j in t.order[s, t.start[s, c]:t.start[s, c + 1]]]
            assert keys == sorted(keys)
    # Frenet on the bent reference line and stepping along it
    s0, d0 = dm.frenet(1, np.nan_to_num(xs[0], nan=5.0), ys[0])
    want = [pm.ref[1].to_frenet_frame(x, y) for x, y in zip(np.nan_to_num(xs[0], nan=5.0), ys[0])]
    assert same(s0, [w[0] for w in want]) and same(d0, [w[1] for w in want])
    good = [i for i in range(n) if t.lane[0, i] >= 0]
    v = [rng.uniform(0, 35) for _ in good]
    ax = [rng.uniform(-8, 2) for _ in good]
    vy = [rng.choice([0.0, 1.0, -1.0]) for _ in good]
    out = dm.step_along_path([int(t.lane[0, i]) for i in good], xs[0, good], ys[0, good], v, ax, vy, 0.2)
    for k, i in enumerate(good):
        want = eo.step_along_path(pm, int(t.lane[0, i]), float(xs[0, i]), float(ys[0, i]), v[k], ax[k], vy[k], 0.2)
        assert tuple(float(o[k]) for o in out) == want
    ms = dm.last_kernel_ms()
    assert ms[0] > 0 and ms[1] > 0
    dm.close()


def test_capacity_and_argument_errors(oe):
    dm = oe.DeviceMap(bent_map())
    with pytest.raises(oe.EnvError) as e:
        dm.neighbors([0], [0], [0.0], [0.0])                 # before any match
    assert e.value.code == -1
    with pytest.raises(oe.EnvError) as e:
        dm.match(np.zeros(1025), np.zeros(1025), 1)
    assert e.value.code == -3
    t = dm.match(np.array([float("nan")] * 4), np.zeros(4), 1)     # a scene of empty slots
    assert list(t.lane[0]) == [-1] * 4 and list(t.order[0]) == [-1] * 4 and list(t.start[0]) == [0] * 9
    with pytest.raises(oe.EnvError):
        dm.neighbors([0], [7], [0.0], [0.0])                 # slot outside the matched batch
    with pytest.raises(oe.EnvError):
        dm.frenet(5, [0.0], [0.0])
    dm.close()


@pytest.mark.parametrize("rec", [r for r in STATES if r["decisions"]], ids=[r["tag"] for r in STATES if r["decisions"]])
def test_learned_and_closest_gap_behaviours_equal_the_reference(oe, rec):
    """behaviors.py:89-176 end to end — outer environment, builder, all candidates in one rollout launch, learned model,
    single-step behaviour — against the reference's own behaviors.py running on the reference's machine code with the
    same seeds: the same chosen command and bit-identical (ax, vy)."""
    from interactive_highway_simulation_b200 import behaviors as bh, learned_model as lm, mc_sim_highway as ms
    env = build_env(oe, rec)
    for d in rec["decisions"]:
        a = env.agents[d["ego"]]
        fresh = DuckAgent(d["agent"])
        keep = (a.behavior_model, a.rss_param)
        a.behavior_model, a.rss_param = fresh.behavior_model, fresh.rss_param
        a.learned_merging_model, a.learned_lane_change_model = lm.LearnedMergingModel(), lm.LearnedLaneChangeModel()
        a.current_decision = "Initial"
        np.random.seed(d["np_seed"]); ms.seed(d["backend_seed"])
        act = getattr(bh, d["fn"])(env, a)
        assert act.ref_path.id == d["ref_lane"]
        if d["decision"] is not None and "closest" not in d["fn"]:
            assert a.current_decision == d["decision"], (d["fn"], d["ego"], a.current_decision, d["decision"])
        assert (act.ax, act.vy) == (d["ax"], d["vy"]), (d["fn"], d["ego"], act.ax, d["ax"], act.vy, d["vy"])
        a.behavior_model, a.rss_param = keep


@pytest.mark.parametrize("rec", BORDER, ids=[r["tag"] for r in BORDER])
def test_border_and_centre_distance_kernels(oe, rec):
    """mce_center_distance / mce_border_distance: bit-identical to the oracle, and equal (1e-9) to what the reference's own
    Corridor.distance_to_border / distance_to_corridor_end / centerSimplified.signed_distance answered."""
    pm = eo.PreparedMap(corridor_dicts(rec))
    dm = oe.DeviceMap([{**c} for c in rec["corridors"]])
    pos = {c["id"]: k for k, c in enumerate(rec["corridors"])}
    S = rec["samples"]
    lane = [pos[s["corridor"]] for s in S]
    d, e = dm.center_distance(lane, [s["x"] for s in S], [s["y"] for s in S])
    want = [eo.center_distance(pm, l, s["x"], s["y"]) for l, s in zip(lane, S)]
    assert same(d, [w[0] for w in want]) and same(e, [w[1] for w in want])
    assert all(abs(g - s["signed"]) <= TOL for g, s in zip(d, S)) and all(abs(g - s["to_end"]) <= TOL for g, s in zip(e, S))
    for side, key in ((0, "left"), (1, "right")):
        got = dm.border_distance(lane, [side] * len(S), [s["hull"] for s in S])
        assert same(got, [eo.border_distance(pm, l, side, s["hull"]) for l, s in zip(lane, S)]), key
        assert all(abs(g - s[key]) <= TOL for g, s in zip(got, S)), key
    with pytest.raises(oe.EnvError):
        dm.border_distance([0], [2], [S[0]["hull"]])
    with pytest.raises(oe.EnvError):
        dm.center_distance([len(pos)], [0.0], [0.0])
    dm.close()

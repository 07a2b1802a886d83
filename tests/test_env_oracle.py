"""oracle/env_oracle.py against the records written by the reference's own environment.py / type.py / utility.py
(tests/golden/make_env_golden.py): lanes, per-lane orders and neighbour ids exact, arc lengths / distances / Frenet
coordinates / stepped positions within 1e-9 (the reference evaluates hypot and arctan2/sin, the oracle sqrt and a cross
product)."""
import math

import pytest

from env_cases import BORDER, STATES, corridor_dicts, prepared, slots, TOL
from oracle import env_oracle as eo


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_match_and_lane_vectors(rec):
    pm, ids, xs, ys = prepared(rec)
    lane, arc, order, start = eo.match(pm, xs, ys)
    cid = [c["id"] for c in rec["corridors"]]
    for i, a in enumerate(ids):
        want = rec["lane"][str(a)]
        assert (cid[lane[i]] if lane[i] >= 0 else None) == want, (a, lane[i], want)
    for c, k in enumerate(cid):
        want = rec["lane_vectors"][str(k)]
        got = order[start[c]:start[c + 1]]
        assert [ids[j] for j in got] == [w[1] for w in want], k
        for j, w in zip(got, want):
            assert abs(arc[j] - w[0]) <= TOL


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_neighbors_at_match_time_and_after_moves(rec):
    pm, ids, xs, ys = prepared(rec)
    lane, arc, order, start = eo.match(pm, xs, ys)
    slot = slots(ids)

    def check(table, qx, qy):
        for a, want in table.items():
            i = slot[int(a)]
            idx, dis = eo.neighbors(pm, lane, arc, order, start, i, qx[i], qy[i])
            for k in range(10):
                if want[k] is None:
                    assert idx[k] == -1, (a, k, ids[idx[k]])
                else:
                    assert idx[k] >= 0 and ids[idx[k]] == want[k][0], (a, k, idx[k], want[k])
                    assert abs(dis[k] - want[k][1]) <= TOL

    check(rec["neighbors"], xs, ys)
    qx, qy = list(xs), list(ys)
    for m in rec["moved"]:
        qx[slot[m["id"]]], qy[slot[m["id"]]] = m["x"], m["y"]
    check(rec["neighbors_after_move"], qx, qy)


@pytest.mark.parametrize("rec", STATES, ids=[r["tag"] for r in STATES])
def test_within_range_frenet_and_step(rec):
    pm, ids, xs, ys = prepared(rec)
    lane, arc, order, start = eo.match(pm, xs, ys)
    slot = slots(ids)
    cpos = {c["id"]: k for k, c in enumerate(rec["corridors"])}
    for w in rec["within_range"]:
        got = eo.within_range(pm, lane, xs, ys, cpos[w["corridor"]], slot[w["ego"]], w["max_dis"])
        assert [ids[j] for j in got] == w["ids"]
    for c, rows in rec["frenet"].items():
        for a, (s, d) in zip(rec["frenet_ids"], rows):
            gs, gd = pm.ref[cpos[int(c)]].to_frenet_frame(xs[slot[a]], ys[slot[a]])
            assert abs(gs - s) <= TOL and abs(gd - d) <= TOL, (c, a, gs, s, gd, d)
    by_id = {a["id"]: a for a in rec["agents"]}
    for st in rec["steps"]:
        a = by_id[st["id"]]
        x, y, v, yaw = eo.step_along_path(pm, cpos[st["lane"]], a["x"], a["y"], a["v"], st["ax"], st["vy"], st["dt"])
        for g, w in zip((x, y, v, yaw), st["out"]):
            assert abs(g - w) <= TOL


def test_cross_product_sign_equals_the_reference_angle_test():
    """type.py:297-305 decides the side with sin(arctan2(p - foot) - yaw) > 0; the oracle and the kernels with a cross
    product.  Same answer on a bent polyline for points off the line."""
    import random
    rng = random.Random(5)
    path = [(0.0, 0.0), (50.0, 2.0), (120.0, -6.0), (200.0, 30.0)]
    ls = eo.LineSimplified(path)
    for _ in range(4000):
        px, py = rng.uniform(-40, 240), rng.uniform(-40, 60)
        s, dis = eo.project_and_distance(ls.line, px, py)
        if dis < 1e-6:
            continue
        fx, fy = ls.get_point_at_length(s)
        yaw = ls.yaw_list[ls.get_yaw_index(s)]
        ref = 1 if math.sin(math.atan2(py - fy, px - fx) - yaw) > 0 else -1
        assert ls.side_of(s, px, py) == ref


@pytest.mark.parametrize("rec", BORDER, ids=[r["tag"] for r in BORDER])
def test_border_and_centre_distances(rec):
    """Corridor.distance_to_border / distance_to_corridor_end / centerSimplified.signed_distance as the reference's own
    type.py answered (tests/golden/make_border_golden.py)"""
    pm = eo.PreparedMap(corridor_dicts(rec))
    pos = {c["id"]: k for k, c in enumerate(rec["corridors"])}
    for s in rec["samples"]:
        lane = pos[s["corridor"]]
        d, to_end = eo.center_distance(pm, lane, s["x"], s["y"])
        assert abs(d - s["signed"]) <= TOL and abs(to_end - s["to_end"]) <= TOL, s
        for side, key in ((0, "left"), (1, "right")):
            got = eo.border_distance(pm, lane, side, s["hull"])
            assert abs(got - s[key]) <= TOL, (s, key, got)


BUILDER_RECS = [r for r in STATES if r["builders"]]


@pytest.mark.parametrize("rec", BUILDER_RECS, ids=[r["tag"] for r in BUILDER_RECS])
def test_builders_equal_the_reference(rec):
    """oracle build_scene (the restatement of mcs_wrapper.py:24-283 the device builder is checked against) on the records
    the reference's own three builders wrote: same corridors, vehicles (order, ids, behaviours, parameters, vd noise) and
    candidate gaps."""
    import numpy as np
    from interactive_highway_simulation_b200 import _abi as abi
    pm, ids, xs, ys = prepared(rec)
    tables = eo.match(pm, xs, ys)
    slot = slots(ids)
    info = [{"id": c["id"], "type": abi.LANE_TYPES.get(c["type"], abi.LANE_OTHER), "v_limit": c["v_limit"]} for c in rec["corridors"]]
    attrs = [dict(v=a["v"], vy=a["vy"], a=a["a"], l=a["l"], w=a["w"], vd=a["idm"][0]) for a in rec["agents"]]
    kinds = {"merging_mcs_sim_from_env": 0, "exit_mcs_sim_from_env": 1, "lane_change_mcs_sim_from_env": 2}
    lane_names = {v: k for k, v in abi.LANE_TYPES.items()}
    beh_names = {v: k for k, v in abi.BEHAVIORS.items()}
    for b in rec["builders"]:
        a = next(x for x in rec["agents"] if x["id"] == b["ego"])
        kind = kinds[b["builder"]]
        merging_model = "merging" in a["behavior_model"]
        i, s, m, y = a["idm"], a["rss"], a["mobil"], a["yielding"]
        ego = dict(behavior=abi.BEHAVIORS["exit" if kind == 1 else ("merging" if merging_model else "idm")], vd=i[0],
                   idm=(i[1], i[3], i[2], i[4], -8.0, i[5]), rss=tuple(s) if kind == 2 else (0.7, 0.4, -8.0, -10.0, 1.8, 1.2),
                   yp=(y[0], y[1], y[2], y[3], m[0], m[1], m[2]), commit=0, keep=0, target=-1)
        np.random.seed(b["np_seed"])
        noise = list(np.random.normal(0., 1, 48))
        cs, ag, actions = eo.build_scene(pm, info, tables, xs, ys, attrs, ids, kind, slot[b["ego"]], ego, merging_model, noise)
        assert (actions if kind != 2 else None) == b["possible_actions"], b["builder"]
        assert len(cs) == len(b["corridors"]) and len(ag) == len(b["agents"]), (b["builder"], b["ego"])
        for c, w in zip(cs, b["corridors"]):
            assert [c[0], lane_names[c[1]]] == w[:2]
            g = c[2] + c[3] + c[4] + [c[5]]
            assert all(abs(p - q) <= TOL for p, q in zip(g, w[2:])), (b["builder"], b["ego"], g, w)
        for r, w in zip(ag, b["agents"]):
            g = [r["id"], r["x"], r["y"], r["v"], r["vy"], r["vd"], r["a"], r["l"], r["w"]] + list(r["idm"]) + list(r["rss"]) + \
                list(r["yp"]) + [beh_names[r["behavior"]]]
            assert g[0] == w[0] and g[-1] == w[-1], (b["builder"], b["ego"], g[0], w[0])
            assert all(abs(p - q) <= TOL for p, q in zip(g[1:-1], w[1:-1])), (b["builder"], b["ego"], g, w)

"""Synthetic rollout scenes in the shape the reference hands to its inner simulator.

Follows the reference's own scene generator and marshalling:
  * per-lane spawn by time headway, speeds ~ N(v_limit + dv, v_var), lateral offset ~ N(0, 0.2), truck ratio on the
    truck lane, vehicle shapes       — random_traffic_generator.py:77-136 (RandomTrafficGenerator.random_agents_on_lane)
  * ego keeps its own IDM / yielding / MOBIL parameters, every other vehicle gets the generic
    idm_param(l) / rss_param(False) / YieldingParam() and vd + N(0, 1)
                                       — mcs_wrapper.py:8-21, 193-283 (lane_change_mcs_sim_from_env)
  * maps: straight lanes in the Frenet frame of the ego lane, 3.5 m wide, speed limits
    100/3.6 (+20/3.6, +30/3.6 on the lanes to the left)          — evaluation_lane_change.py:18-44
Unlike mcs_wrapper the 80 m neighbourhood cut is not applied: all vehicles of the scene are simulated
(BASELINE.json's scaled-up configurations: 40 / 100 / 200 / 256 vehicles per rollout).
"""
import numpy as np

from . import _abi as abi

SHAPES = [(4.5, 1.9), (4.8, 2.0), (4.9, 2.1), (5.2, 2.2), (5.9, 2.4)]
TRUCK_SHAPES = [(11.5, 2.5), (12.0, 2.5), (10.0, 2.5)]
LANE_WIDTH = 3.5
GENERIC_YP = (-0.5, 1.81, -4.8, -1.1, 0.9, 0.5, 0.51)          # YieldingParam() defaults (mc_sim_highway.so 0x12ebe0)
GENERIC_RSS = (0.7, 0.4, -6.0, -6.0, 1.8, 1.2)                 # rss_param(False)      mcs_wrapper.py:16
EGO_RSS = (0.7, 0.4, -8.0, -10.0, 1.8, 1.2)                    # RSSParameter() defaults, type.py:36-43


def generic_idm(l):
    """idm_param(l) of mcs_wrapper.py:8-13 in the binary's memory order (accMax minDis accMin accCom aE thw)."""
    return (0.8, 2.0, -1.5, -2.0, -8.0, 1.2) if l > 7 else (1.6, 2.0, -3.0, -2.0, -8.0, 1.2)


def lane_change_scene(n_vehicles, n_lanes=4, seed=0, truck_ratio=0.4, thw_mean=1.8, thw_var=0.4, v_var=2.0,
                      ego_behavior="idm", merging_lane=False):
    """Return (corridors, agents, ego_id) as ctypes arrays of abi.Corridor / abi.Agent.

    n_vehicles are spread evenly over the lanes; lanes are as long as needed (the reference's 500 m
    s_range would cap a lane at ≈ 10 vehicles).
    """
    rs = np.random.RandomState(seed)
    v_limits = [100 / 3.6, 100 / 3.6, 100 / 3.6 + 20 / 3.6, 100 / 3.6 + 30 / 3.6]
    while len(v_limits) < n_lanes:
        v_limits.append(v_limits[-1])
    lanes = []
    for k in range(n_lanes):
        typ = "merging" if (merging_lane and k == 0) else "main"
        lanes.append(dict(id=k, type=typ, right_y=k * LANE_WIDTH, left_y=(k + 1) * LANE_WIDTH,
                          center_y=(k + 0.5) * LANE_WIDTH, v_limit=v_limits[k]))
    per_lane = [n_vehicles // n_lanes + (1 if k < n_vehicles % n_lanes else 0) for k in range(n_lanes)]
    main_limits = [l["v_limit"] for l in lanes if l["type"] == "main"]
    truck_lane = min((l for l in lanes if l["type"] == "main"), key=lambda l: l["center_y"])["id"]
    records = []
    x_max = 0.0
    for lane, count in zip(lanes, per_lane):
        x = rs.uniform(0.0, 0.3 * lane["v_limit"] * thw_mean)
        for _ in range(count):
            y_off = rs.normal(0, 0.2)
            v = rs.normal(lane["v_limit"], v_var)
            truck = lane["id"] == truck_lane and rs.uniform(0, 1) < truck_ratio
            l, w = (TRUCK_SHAPES if truck else SHAPES)[rs.randint(0, 3 if truck else 5)]
            vd_base = min(main_limits) if truck else main_limits[rs.randint(0, len(main_limits))]
            vd = vd_base + rs.normal(0.0, 1.0 if truck else 0.5)
            records.append(dict(lane=lane["id"], x=x, y=lane["center_y"] + y_off, v=max(v, 1.0), vd=vd, l=l, w=w))
            x_max = max(x_max, x)
            x += rs.uniform(lane["v_limit"] * (thw_mean - thw_var), lane["v_limit"] * (thw_mean + thw_var)) + l
    # ego: a vehicle near the middle of a middle main lane
    mains = [l["id"] for l in lanes if l["type"] == "main"]
    ego_lane = mains[min(1, len(mains) - 1)]
    in_lane = [i for i, r in enumerate(records) if r["lane"] == ego_lane]
    ego_rec = in_lane[len(in_lane) // 2]
    corridors = (abi.Corridor * n_lanes)()
    x0, x1 = -1000.0, x_max + 6000.0
    for o, lane in zip(corridors, lanes):
        o.id = lane["id"]; o.type = abi.LANE_TYPES[lane["type"]]
        o.left[:] = [x0, lane["left_y"], x1, lane["left_y"]]
        o.right[:] = [x0, lane["right_y"], x1, lane["right_y"]]
        o.center[:] = [x0, lane["center_y"], x1, lane["center_y"]]
        o.v_limit = lane["v_limit"]
    agents = (abi.Agent * len(records))()
    order = [ego_rec] + [i for i in range(len(records)) if i != ego_rec]   # mcs_wrapper puts the ego first
    for slot, i in enumerate(order):
        r = records[i]
        o = agents[slot]
        o.id = i
        o.x, o.y, o.v, o.vy, o.a, o.l, o.w = r["x"], r["y"], r["v"], 0.0, 0.0, r["l"], r["w"]
        if i == ego_rec:
            o.behavior = abi.BEHAVIORS[ego_behavior]
            o.vd = r["vd"]
            # random_idm_p (random_traffic_generator.py:8-15), aEmergency fixed to -8 by mcs_wrapper.py:195-199
            o.idm[:] = [rs.uniform(1.8, 2.2), max(2.0, rs.normal(4, 1)), rs.uniform(-4.0, -2.0), rs.normal(-2.0, 0.1), -8.0,
                        max(1.0, rs.normal(1.5, 0.1))]
            o.rss[:] = EGO_RSS
            delta_a = rs.normal(0.5, 0.2)
            o.yp[:] = [1.70968573, 0.00582201, 0.40742229, 1.02390871, rs.uniform(0, 1), delta_a,
                       delta_a + rs.normal(0.1, 0.05)]
        else:
            o.behavior = abi.BEHAVIORS["merging" if lanes[r["lane"]]["type"] == "merging" else "idm_lc"]
            o.vd = r["vd"] + rs.normal(0.0, 1.0)
            o.idm[:] = generic_idm(r["l"])
            o.rss[:] = GENERIC_RSS
            o.yp[:] = GENERIC_YP
        o.commit = abi.COMMITS["not_decided"]; o.commit_keep_lane_step = 0; o.target_lane_id = -1
    return corridors, agents, ego_rec


def merging_scene(n_vehicles=40, seed=0, n_main=3):
    """BASELINE.json configs[1] shape: a merging lane (id 0) + n_main main lanes, ego `merging` on the merging lane, the
    other merging-lane vehicles `merging`, main-lane vehicles `idm_lc` with the generic parameters and rss_param(True)
    (mcs_wrapper.py:24-75).  Returns (corridors, agents, ego_id, gap_ids) with gap_ids = [-1] + up to three vehicles of
    the adjacent main lane around the ego (the reference's `possible_actions`, mcs_wrapper.py:40,57-59)."""
    corridors, agents, _ = lane_change_scene(n_vehicles, n_main + 1, seed=seed, merging_lane=True)
    on_merge = [a for a in agents if 0.0 <= a.y < LANE_WIDTH]
    if not on_merge:
        raise ValueError("no vehicle on the merging lane")
    ego = on_merge[len(on_merge) // 2]
    rs = np.random.RandomState(seed + 1000)
    for a in agents:
        if a.id == ego.id:
            a.behavior = abi.BEHAVIORS["merging"]
            a.idm[:] = [rs.uniform(1.8, 2.2), max(2.0, rs.normal(4, 1)), rs.uniform(-4.0, -2.0), rs.normal(-2.0, 0.1), -8.0,
                        max(1.0, rs.normal(1.5, 0.1))]
            a.rss[:] = EGO_RSS
            a.yp[:] = [1.70968573, 0.00582201, 0.40742229, 1.02390871, rs.uniform(0, 1), 0.5, 0.6]
        else:
            a.idm[:] = generic_idm(a.l)
            a.yp[:] = GENERIC_YP
            a.rss[:] = EGO_RSS if a.behavior == abi.BEHAVIORS["merging"] else GENERIC_RSS
    left = sorted((a for a in agents if LANE_WIDTH <= a.y < 2 * LANE_WIDTH), key=lambda a: abs(a.x - ego.x))[:3]
    gaps = [-1] + [a.id for a in sorted(left, key=lambda a: -a.x)]
    return corridors, agents, ego.id, gaps


def exit_scene(n_vehicles=200, seed=0, n_main=3):
    """BASELINE.json configs[3] shape: an exit lane (id 0) + n_main main lanes, ego `exit` on the main lane next to it,
    every other vehicle `idm_lc` with the generic parameters (mcs_wrapper.py:78-190).  Returns (corridors, agents, ego_id,
    gap_ids) with gap_ids = [-1] + up to three vehicles of the exit lane around the ego (mcs_wrapper.py:155-157)."""
    corridors, agents, _ = lane_change_scene(n_vehicles, n_main + 1, seed=seed)
    corridors[0].type = abi.LANE_TYPES["exit"]
    on_main = sorted((a for a in agents if LANE_WIDTH <= a.y < 2 * LANE_WIDTH), key=lambda a: a.x)
    ego = on_main[len(on_main) // 2]
    rs = np.random.RandomState(seed + 2000)
    for a in agents:
        if a.id == ego.id:
            a.behavior = abi.BEHAVIORS["exit"]
            a.idm[:] = [rs.uniform(1.8, 2.2), max(2.0, rs.normal(4, 1)), rs.uniform(-4.0, -2.0), rs.normal(-2.0, 0.1), -8.0,
                        max(1.0, rs.normal(1.5, 0.1))]
            a.rss[:] = EGO_RSS
            a.yp[:] = [1.70968573, 0.00582201, 0.40742229, 1.02390871, rs.uniform(0, 1), 0.5, 0.6]
        else:
            a.behavior = abi.BEHAVIORS["idm_lc"]
            a.idm[:] = generic_idm(a.l)
            a.yp[:] = GENERIC_YP
            a.rss[:] = GENERIC_RSS
    right = sorted((a for a in agents if 0.0 <= a.y < LANE_WIDTH), key=lambda a: abs(a.x - ego.x))[:3]
    gaps = [-1] + [a.id for a in sorted(right, key=lambda a: -a.x)]
    return corridors, agents, ego.id, gaps


def generated_scene(seed, n_vehicles=256, n_lanes=4):
    """The throughput workload (BASELINE.json configs[4]) on a scene of the reference's own generator, as BASELINE.md §3 /
    SURVEY.md §8(d) prescribe: `np.random.seed(seed); random.seed(seed)`, then RandomTrafficGenerator.random_agents_on_map
    (traffic_generator.py, pinned on the reference's random_traffic_generator.py) on the lanes of
    evaluation_lane_change.py:18-44 — speed limits 100/3.6, 100/3.6, +20/3.6, +30/3.6, headway parameters (1.8, 0.4) on
    the two right lanes and (2.0, 1.0) on the two left ones, truck ratio 0.4 on the rightmost lane — made long enough for
    n_vehicles / n_lanes vehicles per lane (the shipped 500 m hold about ten) and with the ramp replaced by a fourth main
    lane (configs[4] is IDM / MOBIL only).  Marshalled like lane_change_mcs_sim_from_env (mcs_wrapper.py:193-283) without
    its 80 m cut: Frenet frame of the ego lane's centre line extended by 1000 m, the ego keeps its own parameters, every
    other vehicle becomes `idm_lc` with idm_param(l) / rss_param(False) / YieldingParam() and vd + N(0, 1) drawn in
    list order.  Returns (corridors, agents, ego_id) as ctypes arrays."""
    import random
    from . import traffic_generator as tg
    np.random.seed(seed)
    random.seed(seed)
    base = 100 / 3.6
    limits = [base, base, base + 20 / 3.6, base + 30 / 3.6]
    headway = [(1.8, 0.4), (1.8, 0.4), (2.0, 1.0), (2.0, 1.0)]
    while len(limits) < n_lanes:
        limits.append(limits[-1]); headway.append(headway[-1])
    per_lane = [n_vehicles // n_lanes + (1 if k < n_vehicles % n_lanes else 0) for k in range(n_lanes)]
    x0, x1 = -200.0, 200.0 + max(per_lane) * 120.0
    lanes, params = {}, {}
    for k in range(n_lanes):
        lo, hi, mid = k * LANE_WIDTH, (k + 1) * LANE_WIDTH, (k + 0.5) * LANE_WIDTH
        lanes[k] = tg.StraightLane(k, "main", [[x0, hi], [x1, hi]], [[x0, lo], [x1, lo]], [[x0, mid], [x1, mid]], limits[k],
                                   k + 1 if k + 1 < n_lanes else None, k - 1 if k > 0 else None)
        params[k] = tg.RandomDensityAndVelocityParameter(headway[k][0], headway[k][1], 0, 2, [0, x1 - x0], per_lane[k])
    _, vehicles = tg.RandomTrafficGenerator().random_agents_on_map(lanes, params, unsafe_lc_ratio=0., truck_ratio=0.4)
    if len(vehicles) != n_vehicles:
        raise ValueError("generator produced %d vehicles, wanted %d" % (len(vehicles), n_vehicles))
    ego_lane = min(1, n_lanes - 1)
    mid_y = (ego_lane + 0.5) * LANE_WIDTH
    in_lane = sorted((v for v in vehicles if ego_lane * LANE_WIDTH <= v["y"] < (ego_lane + 1) * LANE_WIDTH), key=lambda v: v["x"])
    ego = in_lane[len(in_lane) // 2]
    s_of = lambda x: x - (x0 - 1000.0)            # arc length on the ego lane's centre line extended by 1000 m
    corridors = (abi.Corridor * n_lanes)()
    for o, k in zip(corridors, range(n_lanes)):
        o.id = k; o.type = abi.LANE_TYPES["main"]
        o.left[:] = [s_of(x0), (k + 1) * LANE_WIDTH - mid_y, s_of(x1), (k + 1) * LANE_WIDTH - mid_y]
        o.right[:] = [s_of(x0), k * LANE_WIDTH - mid_y, s_of(x1), k * LANE_WIDTH - mid_y]
        o.center[:] = [s_of(x0), (k + 0.5) * LANE_WIDTH - mid_y, s_of(x1), (k + 0.5) * LANE_WIDTH - mid_y]
        o.v_limit = limits[k]
    agents = (abi.Agent * n_vehicles)()
    order = [ego] + [v for v in vehicles if v is not ego]         # the builders put the ego first
    for o, v in zip(agents, order):
        o.id = v["id"]
        o.x, o.y, o.v, o.vy, o.a, o.l, o.w = s_of(v["x"]), v["y"] - mid_y, v["v"], v["vy"], v["a"], v["l"], v["w"]
        o.commit = abi.COMMITS["not_decided"]; o.commit_keep_lane_step = 0; o.target_lane_id = -1
        if v is ego:
            ip, mp, yp = v["idm_p"], v["mobil_param"], v["yielding_param"]
            o.behavior = abi.BEHAVIORS["idm"]
            o.vd = ip["vd"]
            o.idm[:] = [ip["acc_max"], ip["min_dis"], ip["acc_min"], ip["acc_com"], -8.0, ip["thw"]]
            o.rss[:] = EGO_RSS
            o.yp[:] = [yp[0], yp[1], yp[2], yp[3], mp["politeness_factor"], mp["delta_a"], mp["bias_a"]]
        else:
            o.behavior = abi.BEHAVIORS["idm_lc"]
            o.vd = v["idm_p"]["vd"] + np.random.normal(0., 1)
            o.idm[:] = generic_idm(v["l"])
            o.rss[:] = GENERIC_RSS
            o.yp[:] = GENERIC_YP
    return corridors, agents, ego["id"]

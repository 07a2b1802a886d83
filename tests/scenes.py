"""Scenes shared by the parity tests (SceneSpec = plain description accepted by the oracle, the reference
harness and the product)."""
import random

from oracle.pyoracle import SceneSpec

ACTIONS = ("keep_lane", "left", "right", "acc", "dcc")


def probe_scene():
    """the scene whose known-answer vectors were captured from the reference binary."""
    s = SceneSpec()
    s.add_corridor(1, "main", 3.5, 0.0, 1.75, 27.78)
    s.add_corridor(2, "main", 7.0, 3.5, 5.25, 33.3)
    s.add_corridor(3, "main", 10.5, 7.0, 8.75, 36.1)
    s.add_agent(0, 100, 5.25, 25, 30, "idm")
    s.add_agent(1, 140, 5.2, 22, 24, "idm_lc")
    s.add_agent(2, 60, 5.3, 27, 29, "idm_lc")
    s.add_agent(3, 120, 1.7, 20, 22, "idm_lc")
    s.add_agent(4, 80, 1.8, 21, 23, "idm_lc")
    s.add_agent(5, 115, 8.7, 30, 33, "idm_lc")
    s.add_agent(6, 70, 8.8, 31, 34, "idm_lc")
    s.add_agent(7, 180, 5.25, 21, 23, "idm_lc")
    return s


def random_scene(rng, n_lanes=3, n_agents=12, trucks=0.3, lane_w=3.5, lane_types=None, ids_shuffled=True,
                 ego_behavior="idm", others=("idm_lc", "idm_lc", "idm_lc", "idm"), ego_lane=None, spacing=(3, 60)):
    """Random straight-lane scene.  lane_types: list of corridor types bottom (right-most) to top, default all main.
    Vehicles on a `merging` lane get behaviour `merging`.  Only the ego may have behaviour `exit`: for any other
    vehicle the reference pushes no action and then reads past its action vector (so:0x10ed15), and mcs_wrapper.py
    never builds such a scene."""
    s = SceneSpec()
    lane_types = lane_types or ["main"] * n_lanes
    n_lanes = len(lane_types)
    for k in range(n_lanes):
        s.add_corridor(k + 1, lane_types[k], (k + 1) * lane_w, k * lane_w, (k + 0.5) * lane_w,
                       rng.choice([22.2, 27.78, 33.3, 36.1]))
    ids = list(range(n_agents))
    if ids_shuffled:
        ids = rng.sample(range(0, 400), n_agents)
    per_lane_x = {k: rng.uniform(0, 40) for k in range(n_lanes)}
    ego_idx = rng.randrange(n_agents)
    for i in range(n_agents):
        k = rng.randrange(n_lanes)
        if i == ego_idx and ego_lane is not None:
            k = ego_lane
        elif lane_types[k] != "main" and rng.random() < 0.6:
            k = rng.choice([j for j in range(n_lanes) if lane_types[j] == "main"])
        truck = rng.random() < trucks
        l = rng.uniform(8, 14) if truck else rng.uniform(4, 5.5)
        x = per_lane_x[k] + l + rng.uniform(*spacing)
        per_lane_x[k] = x
        v = rng.uniform(12, 34)
        idm = (rng.uniform(0.8, 2.0), rng.uniform(1.5, 3), rng.uniform(-4, -2), rng.uniform(-2.5, -1.5), -8.0,
               rng.uniform(0.8, 1.8))
        rss = (0.7, 0.4, rng.choice([-6.0, -8.0]), rng.choice([-6.0, -10.0]), 1.8, 1.2)
        yp = (rng.uniform(-1, 0), rng.uniform(1, 2.5), rng.uniform(-6, -3), rng.uniform(-2, 0), rng.uniform(0, 1),
              rng.uniform(0.1, 1), rng.uniform(0, 1))
        if i == ego_idx:
            beh = ego_behavior
        elif lane_types[k] == "merging":
            beh = "merging"
        else:
            beh = rng.choice(others)
        s.add_agent(ids[i], x, (k + 0.5) * lane_w + rng.gauss(0, 0.25), v, v + rng.uniform(-3, 6), beh,
                    vy=rng.gauss(0, 0.2), a=rng.uniform(-1, 1), l=l, w=2.0, idm=idm, rss=rss, yp=yp)
    return s, ids[ego_idx]


def lane_change_cases(n_scenes, seed0=1000, sizes=(5, 9, 14, 20, 31, 33, 64, 100)):
    """(tag, scene, ego, dt, steps, min_steps, seed) for all-main-lane scenes."""
    out = []
    for sc in range(n_scenes):
        rng = random.Random(seed0 + sc)
        n_lanes = rng.choice([2, 3, 4, 5])
        n_agents = rng.choice(list(sizes))
        s, ego = random_scene(rng, n_lanes, n_agents)
        out.append(("scene%d_n%d_L%d" % (sc, n_agents, n_lanes), s, ego, rng.choice([0.2, 0.3]),
                    rng.choice([25, 40, 60]), rng.choice([10, 30]), 55 + sc))
    return out


STATUS_KEYS = ("ok", "finish", "fallBack", "finishStep", "utility", "comfort", "utilityObjects", "comfortObjects", "draws")


def same(a, b):
    return a == b or (isinstance(a, float) and isinstance(b, float) and a != a and b != b)


def merge_exit_cases(n_scenes, seed0=900, sizes=(4, 7, 12, 20, 30)):
    """(tag, scene, ego, direction, gap ids) — ego `merging` on a merging lane (direction left) or ego `exit` next to an
    exit lane (direction right); main-lane vehicles beside the merging lane exercise the yielding model, vehicles on
    the merging lane run mergingBehaviorClosestGap."""
    out = []
    for sc in range(n_scenes):
        rng = random.Random(seed0 + sc)
        kind = rng.choice(["merge", "exit"])
        na = rng.choice(list(sizes))
        spacing = rng.choice([(3, 60), (2, 25), (10, 90)])
        if kind == "merge":
            lt = ["merging", "main"] + ["main"] * rng.choice([0, 1, 2])
            s, ego = random_scene(rng, n_agents=na, lane_types=lt, ego_behavior="merging", ego_lane=0, spacing=spacing)
            d = "left"
        else:
            lt = ["exit", "main"] + ["main"] * rng.choice([0, 1])
            s, ego = random_scene(rng, n_agents=na, lane_types=lt, ego_behavior="exit", ego_lane=1, spacing=spacing)
            d = "right"
        ids = [a["id"] for a in s.agents]
        gaps = [-1] + rng.sample(ids, min(2, len(ids)))
        out.append(("%s%d_n%d_L%d" % (kind, sc, na, len(lt)), s, ego, d, gaps, 40 + sc))
    return out


def has_side_lane(scene, agent, direction, lane_w=3.5):
    k = int(agent["y"] // lane_w)
    n = len(scene.corridors)
    return 0 <= k < n and ((direction == "left" and k + 1 < n) or (direction == "right" and k - 1 >= 0))

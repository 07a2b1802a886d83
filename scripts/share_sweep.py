"""Own vs shared thread blocks at the launch-sized thread count, several launch sizes (diagnostic for MCS_SHARE_MIN_PER_SM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
def sweep(name, sc, cands, sp, counts, seed):
    for R in counts:
        row = []
        for share in (None, 0, 1, 2, 4, 1000):
            if share is None: os.environ.pop("MCS_SHARE_MIN_PER_SM", None)
            else: os.environ["MCS_SHARE_MIN_PER_SM"] = str(share)
            best = 1e9
            for i in range(4):
                ms, vs = sc.rollouts_device(cands, sp, seed, 0, R); best = min(best, ms)
            row.append("%s %.3f" % ("auto" if share is None else share, best))
        print("%s x %4d rollouts: share_min %s" % (name, R, " | ".join(row)))
acts = ["keep_lane", "dcc", "acc", "left", "right"]
for nveh in (20, 40):
    corridors, agents, ego, gaps = synthetic.merging_scene(nveh, 0)
    sc = backend.DeviceScene(corridors, agents)
    cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
    sweep("merging %d veh, %d cand" % (nveh, len(gaps)), sc, cands, abi.SimParam(0.3, 40, 10 if nveh == 40 else 30), (16, 32, 64, 120, 160, 200, 256, 512), 1000)
    sc.close()
for nveh in (24, 100):
    corridors, agents, ego = synthetic.lane_change_scene(nveh, 4, seed=3)
    sc = backend.DeviceScene(corridors, agents)
    cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[a], -2, 0) for a in acts])
    sweep("lane change %d veh, 5 cand" % nveh, sc, cands, abi.SimParam(0.3, 40, 30), (16, 32, 64, 160, 320, 640, 1280), 77)
    sc.close()

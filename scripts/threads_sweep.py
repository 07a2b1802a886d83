"""Kernel time of a few launch sizes at every threads-per-rollout setting (diagnostic for capi.cu rollout_threads)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
def sweep(name, sc, cands, sp, counts, seed):
    for R in counts:
        row = []
        for T in (None, 32, 64, 128, 256):
            if T is None: os.environ.pop("MCS_ROLLOUT_THREADS", None)
            else: os.environ["MCS_ROLLOUT_THREADS"] = str(T)
            best = 1e9
            for i in range(4):
                ms, vs = sc.rollouts_device(cands, sp, seed, 0, R); best = min(best, ms)
            row.append("%s %.3f" % ("auto" if T is None else T, best))
        print("%s x %4d rollouts: %s" % (name, R, " | ".join(row)))
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sweep("merging 40 veh, 4 cand", sc, cands, abi.SimParam(0.3, 40, 10), (32, 64, 128, 256, 512), 1000)
sc.close()
corridors, agents, ego = synthetic.lane_change_scene(24, 4, seed=3)
sc = backend.DeviceScene(corridors, agents)
acts = ["keep_lane", "dcc", "acc", "left", "right"]
cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[a], -2, 0) for a in acts])
sweep("lane change 24 veh, 5 cand", sc, cands, abi.SimParam(0.3, 40, 30), (32, 64, 160, 320, 640, 1280), 77)
sc.close()
corridors, agents, ego = synthetic.lane_change_scene(100, 4, seed=1)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[a], -2, 0) for a in acts])
sweep("lane change 100 veh, 5 cand", sc, cands, abi.SimParam(0.3, 40, 30), (32, 160, 640), 78)
sc.close()

"""Tiny run of every kernel for compute-sanitizer (racecheck / memcheck): throughput scene (few rollouts), merging scene,
trajectory variant, the four single decisions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
corridors, agents, ego = synthetic.lane_change_scene(256, 4, seed=0)
sc = backend.DeviceScene(corridors, agents)
cand = (abi.Candidate * 2)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0), abi.Candidate(ego, abi.ACTIONS["left"], -2, 0))
sp = abi.SimParam(0.3, 30, 10)
feats, _ = sc.rollouts(cand, sp, max(1, R // 8), 5)
print("lane change", feats[0].utility, feats[1].successRate)
buf, ns, st = sc.trajectories(cand[1], sp, 5, 0)
print("traj states", ns)
o = sc.decide_mobil(agents[3].id, 0.3, 0.0, True, 9); print("mobil", o.ax, o.commit)
o = sc.decide_fixed_direction(ego, abi.ACTIONS["left"]); print("fixed dir", o.ax, o.vy)
sc.close()
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
feats, _ = sc.rollouts(cands, abi.SimParam(0.3, 40, 10), max(1, R // 8), 7)
print("merging", [f.successRate for f in feats])
o = sc.decide_fixed_gap(ego, abi.ACTIONS["left"], gaps[1]); print("fixed gap", o.ax, o.vy)
o = sc.decide_closest_gap(ego, abi.ACTIONS["left"]); print("closest gap", o.ax, o.vy)
sc.close()
# exit scene of 200 vehicles: the helper-warp instantiation, first in blocks of its own, then three rollouts per 768-thread block
c, a, ego, gaps = synthetic.exit_scene(200, 0)
sc = backend.DeviceScene(c, a)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps])
feats, _ = sc.rollouts(cands, abi.SimParam(0.3, 12, 4), 1, 3)
print("exit, own blocks", [f.successRate for f in feats])
ms, vs = sc.rollouts_device(cands, abi.SimParam(0.3, 8, 4), 3, 0, 120)
print("exit, shared blocks", ms, vs)
sc.close()

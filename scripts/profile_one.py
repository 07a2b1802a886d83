"""One launch of the throughput configuration for ncu (diagnostic). usage: profile_one.py [rollouts] [lib]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 2:
    backend.LIB_PATH = os.path.abspath(sys.argv[2])
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
corridors, agents, ego = synthetic.lane_change_scene(256, 4, seed=0)
sc = backend.DeviceScene(corridors, agents)
cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
sp = abi.SimParam(0.3, 100, 100)
for i in range(3):
    ms, vs = sc.rollouts_device(cand, sp, 5, 0, R)
print("%d rollouts %.2f ms -> %.3e vehicle-steps/s" % (R, ms, vs / ms * 1e3))
sc.close()

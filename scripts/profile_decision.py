"""One launch of the decision configuration for ncu (diagnostic).  usage: profile_decision.py [rollouts per candidate]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for i in range(3):
    ms, vs = sc.rollouts_device(cands, sp, 1000, 0, R)
print("decision config %.3f ms %d vehicle-steps" % (ms, vs))
sc.close()

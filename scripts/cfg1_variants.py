import sys, os
sys.path.insert(0, os.getcwd())
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
for R in [int(x) for x in os.environ.get("CFG1_R", "512,2048").split(",")]:
    best = 1e9
    for i in range(5):
        ms, vs = sc.rollouts_device(cands, sp, 1000, 0, R); best = min(best, ms)
    print(os.environ.get("MCS_KERNEL"), os.environ.get("MCS_ROLLOUT_THREADS"), os.environ.get("MCS_NO_HELPER"), "4 x %d: %.3f ms" % (R, best))

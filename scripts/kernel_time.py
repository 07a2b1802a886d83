"""Kernel time of the two headline configurations (diagnostic; bench.py is the measurement of record).
usage: kernel_time.py [path/to/libmcsim_b200.so]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 1:
    backend.LIB_PATH = os.path.abspath(sys.argv[1])
corridors, agents, ego = synthetic.lane_change_scene(256, 4, seed=0)
sc = backend.DeviceScene(corridors, agents)
cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
sp = abi.SimParam(0.3, 100, 100)
R = 16384
best = 1e9
for i in range(4):
    ms, vs = sc.rollouts_device(cand, sp, 5, 0, R); best = min(best, ms)
print("throughput config: %d rollouts %.2f ms -> %.3e vehicle-steps/s" % (R, best, vs / best * 1e3))
chk = sc.rollouts_range(cand, sp, 5, 0, 64)
import hashlib
print("  checksum", hashlib.md5(bytes(chk)).hexdigest())
sc.close()
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
best = 1e9
for i in range(4):
    ms, vs = sc.rollouts_device(cands, sp, 1000, 0, 512); best = min(best, ms)
print("decision config: 4 x 512 rollouts x 40 vehicles %.3f ms (%d vehicle-steps)" % (best, vs))
chk = sc.rollouts_range(cands, sp, 1000, 0, 64)
print("  checksum", hashlib.md5(bytes(chk)).hexdigest())
sc.close()
corridors, agents, ego = synthetic.lane_change_scene(24, 4, seed=3)
sc = backend.DeviceScene(corridors, agents)
acts = ["keep_lane", "dcc", "acc", "left", "right"]
cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[a], -2, 0) for a in acts])
sp = abi.SimParam(0.3, 40, 30)
best = 1e9
for i in range(4):
    ms, vs = sc.rollouts_device(cands, sp, 77, 0, 160); best = min(best, ms)
print("reference-size lane-change decision: 5 x 160 rollouts x 24 vehicles %.3f ms (%d vehicle-steps)" % (best, vs))
best = 1e9
for i in range(4):
    ms, vs = sc.rollouts_device(cands, sp, 77, 0, 4096); best = min(best, ms)
print("lane-change decision x 4096 rollouts: 5 x 4096 x 24 vehicles %.3f ms (%d vehicle-steps)" % (best, vs))
sc.close()
corridors, agents, ego, gaps = synthetic.merging_scene(20, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 30)
best = 1e9
for i in range(4):
    ms, vs = sc.rollouts_device(cands, sp, 55, 0, 160); best = min(best, ms)
print("reference-size merging decision: %d x 160 rollouts x 20 vehicles %.3f ms (%d vehicle-steps)" % (len(gaps), best, vs))
sc.close()
print("sweep of decision sizes (merging scene, 4 candidates):")
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
for R in (32, 64, 128, 256, 512, 1024):
    best = 1e9
    for i in range(4):
        ms, vs = sc.rollouts_device(cands, sp, 1000, 0, R); best = min(best, ms)
    print("  4 x %4d rollouts: %.3f ms" % (R, best))
sc.close()
print("sweep of decision sizes (24-vehicle lane change, 5 candidates):")
corridors, agents, ego = synthetic.lane_change_scene(24, 4, seed=3)
sc = backend.DeviceScene(corridors, agents)
cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[a], -2, 0) for a in acts])
sp = abi.SimParam(0.3, 40, 30)
for R in (32, 64, 160, 320, 640, 1280):
    best = 1e9
    for i in range(4):
        ms, vs = sc.rollouts_device(cands, sp, 77, 0, R); best = min(best, ms)
    print("  5 x %4d rollouts: %.3f ms" % (R, best))
sc.close()

"""Where does one configs[1] decision spend its time? (diagnostic)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interactive_highway_simulation_b200 import backend, synthetic, sharding, learned_model, _abi as abi
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
model = learned_model.LearnedMergingModel()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for i in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); scene = backend.DeviceScene(corridors, agents); t1 = time.perf_counter()
    feats = sharding.simulation_multi_thread_sharded(scene, cands, sp, 64, 1000 + i); t2 = time.perf_counter()
    idx, _ = model.inference([learned_model.feature_vector(f) for f in feats]); t3 = time.perf_counter()
    act = scene.decide_fixed_gap(ego, abi.ACTIONS["left"], gaps[idx]); t4 = time.perf_counter()
    ms, vs = scene.rollouts_device(cands, sp, 1000 + i, 0, 512); t5 = time.perf_counter()
    scene.close()
    print("create %.3f  sharded %.3f  model %.3f  decide %.3f | rollouts_device %.3f (kernel %.3f ms, %d vehicle-steps)" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), ms, vs))
st = None

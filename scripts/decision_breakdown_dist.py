"""Where does one sharded configs[1] decision spend its time on N ranks? (diagnostic; run under torchrun)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from interactive_highway_simulation_b200 import backend, synthetic, sharding, _abi as abi
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
corridors, agents, ego, gaps = synthetic.merging_scene(40, 0)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sp = abi.SimParam(0.3, 40, 10)
nc = len(cands); first, per = sharding.group_range(rank, world)
dev = torch.device("cuda", local)
rows = []
for i in range(30):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene = backend.DeviceScene(corridors, agents, device=local); t1 = time.perf_counter()
    mine = torch.empty((nc * per, 8), dtype=torch.float64, device=dev); t2 = time.perf_counter()
    scene.rollouts_groups(cands, sp, 64, 1000 + i, first, per, d_partials_ptr=mine.data_ptr()); t3 = time.perf_counter()
    if world > 1:
        gathered = torch.empty((world * nc * per, 8), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
    else:
        gathered = mine
    t4 = time.perf_counter()
    ordered = sharding.interleave_gathered(gathered, nc, world); t5 = time.perf_counter()
    host = ordered.cpu(); t6 = time.perf_counter()
    feats = sharding.features_from_gathered(host, nc); t7 = time.perf_counter()
    scene.close(); t8 = time.perf_counter()
    rows.append([1e3 * (b - a) for a, b in zip((t0, t1, t2, t3, t4, t5, t6, t7), (t1, t2, t3, t4, t5, t6, t7, t8))])
if rank == 0:
    rows = rows[5:]
    med = [sorted(c)[len(c) // 2] for c in zip(*rows)]
    print("world %d  median ms: create %.3f  empty %.3f  rollouts_groups %.3f  all_gather(enqueue) %.3f  interleave %.3f  .cpu() %.3f  combine %.3f  close %.3f  | sum %.3f"
          % ((world,) + tuple(med) + (sum(med),)))
if world > 1: dist.destroy_process_group()

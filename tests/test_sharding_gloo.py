"""N>1 host logic on CPU: world_size-2 (and 4) `gloo` process groups exchange the per-group partial features and every
rank combines them in fixed order — bit-identical to the single-rank reduction.  The partials here come from oracle
statuses through the library's host function mcs_group_partials (no GPU); on a B200 box the same exchange runs over
NCCL on device tensors (tests/test_gpu_parity.py::test_sharded_groups_equal_single_gpu, bench.py --gpus N)."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interactive_highway_simulation_b200 import _abi as abi
from interactive_highway_simulation_b200 import sharding


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def synthetic_statuses(n_cand, m, seed):
    import random
    rng = random.Random(seed)
    st = (abi.Status * (n_cand * 8 * m))()
    for i, s in enumerate(st):
        s.utility = rng.uniform(0.3, 1.1); s.comfort = rng.uniform(0.2, 1.0)
        s.utilityObjects = rng.uniform(0.5, 1.0); s.comfortObjects = rng.uniform(0.5, 1.0)
        s.finish = rng.random() < 0.7; s.fallBack = rng.random() < 0.2
        s.finishStep = rng.randrange(3, 40)
        s.ok = 0 if (i // (8 * m)) == 1 else 1          # candidate 1: setEgo failed -> fromNSims 0
    return st


def worker(rank, world, port, n_cand, m, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from interactive_highway_simulation_b200 import backend
        st = synthetic_statuses(n_cand, m, 5)                       # every rank can build all statuses; it uses only its own
        first, per = sharding.group_range(rank, world)
        mine_st = (abi.Status * (n_cand * per * m))()
        for c in range(n_cand):
            for i in range(per * m):
                mine_st[c * per * m + i] = st[c * 8 * m + first * m + i]
        part = backend.group_partials(mine_st, n_cand, per, m, steps)
        mine = torch.frombuffer(bytearray(bytes(part)), dtype=torch.float64).view(n_cand * per, 8).clone()
        gathered = torch.empty((world * n_cand * per, 8), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, mine)
        ordered = sharding.interleave_gathered(gathered, n_cand, world)
        feats = sharding.features_from_gathered(ordered, n_cand)
        want = backend.reduce_features(st, n_cand, 8, m, steps)
        ok = all(bytes(feats[c]) == bytes(want[c]) for c in range(n_cand))
        with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
            f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_group_partials_all_gather_matches_single_rank(tmp_path, backend_lib, world):
    port = free_port()
    mp.spawn(worker, args=(world, port, 3, 5, 40, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("rank%d" % r)).read_text() == "ok"


def test_ranges():
    assert [sharding.group_range(r, 4) for r in range(4)] == [(0, 2), (2, 2), (4, 2), (6, 2)]
    assert sharding.group_range(0, 1) == (0, 8)
    with pytest.raises(ValueError):
        sharding.group_range(0, 3)
    covered = []
    for r in range(3):
        f, c = sharding.rollout_range(1000, r, 3)
        covered += list(range(f, f + c))
    assert covered == list(range(1000))


def test_interleave_order():
    world, nc = 4, 2
    per = 8 // world
    # record (rank, cand, g) encoded in the first double
    g = torch.zeros((world * nc * per, 8), dtype=torch.float64)
    for r in range(world):
        for c in range(nc):
            for q in range(per):
                g[(r * nc + c) * per + q, 0] = c * 100 + r * per + q
    o = sharding.interleave_gathered(g, nc, world)
    assert [int(v) for v in o[:, 0]] == [c * 100 + k for c in range(nc) for k in range(8)]

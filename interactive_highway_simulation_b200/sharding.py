"""Multi-GPU form of simulationMultiThread: one process per GPU, rollouts sharded by runMultiThreads' own unit — the 8
groups (`std::thread`s, mc_sim_highway.so 0xbb952) — and ONE exchange step: an all-gather of the per-group partial
features (64 B per candidate and group; < 3 KB per decision), after which every rank combines them in group order.
Because the combination order is fixed and independent of the rank count, the features (and the learned model's discrete
decision on top of them) are bit-identical for 1, 2, 4 and 8 GPUs.

torch.distributed is plumbing only: NCCL all_gather_into_tensor on a device tensor the kernels wrote (no host staging),
or gloo on CPU tensors for the host-logic tests.
"""
import ctypes as C

from . import _abi as abi
from . import backend

GROUPS = abi.MCS_GROUPS


def group_range(rank, world):
    """Groups [first, first+count) owned by `rank`; world must divide the reference's 8 groups."""
    if world < 1 or GROUPS % world != 0:
        raise ValueError("world size %d does not divide the %d rollout groups of runMultiThreads" % (world, GROUPS))
    per = GROUPS // world
    return rank * per, per


def rollout_range(total, rank, world):
    """Contiguous slice [first, first+count) of `total` independent rollouts for `rank` (throughput sweeps)."""
    first = total * rank // world
    return first, total * (rank + 1) // world - first


def interleave_gathered(gathered, n_candidates, world):
    """all_gather output [world][n_candidates][groups/world] (64-byte records as rows of 8 doubles) -> the
    [n_candidates][groups] order mcs_combine_groups expects."""
    per = GROUPS // world
    return gathered.view(world, n_candidates, per, 8).permute(1, 0, 2, 3).reshape(n_candidates * GROUPS, 8).contiguous()


def features_from_gathered(gathered_cpu, n_candidates):
    """[n_candidates*groups][8] float64 CPU tensor of partials -> list of abi.Feature (fixed group order)."""
    buf = (abi.GroupPartial * (n_candidates * GROUPS))()
    C.memmove(buf, gathered_cpu.data_ptr(), C.sizeof(buf))
    return backend.combine_groups(buf, n_candidates, GROUPS)


def simulation_multi_thread_sharded(scene, candidates, sim_param, n_per_group, seed, group=None):
    """Every rank calls this with the same arguments and its own DeviceScene; returns the same features on every rank,
    bit-identical to scene.rollouts(...) on one GPU."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nc = len(candidates)
    first, per = group_range(rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.empty((nc * per, 8), dtype=torch.float64, device=dev)
    try:
        scene.rollouts_groups(candidates, sim_param, n_per_group, seed, first, per, d_partials_ptr=mine.data_ptr())
    except backend.McsimError as e:
        # A rollout of THIS rank's groups hit a case the reference leaves undefined.  The exchange below is a collective
        # every rank has to enter, and the mark travels inside the partials (GroupPartial.reserved): mcs_combine_groups
        # then refuses them on every rank alike.  Anything else (CUDA failure, bad arguments) is raised right away.
        if e.code not in (abi.MCS_ERR_SCENE, abi.MCS_ERR_CAPACITY):
            raise
    if world > 1:
        gathered = torch.empty((world * nc * per, 8), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    ordered = interleave_gathered(gathered, nc, world)
    return features_from_gathered(ordered.cpu(), nc)

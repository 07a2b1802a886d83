"""Thin ctypes layer over libmcsim_b200.so (C ABI in include/mcsim.h).

There is no CPU implementation behind this module: if the CUDA library is missing or no B200 is
visible, every compute entry point raises.  (The CPU restatement under oracle/ is test infrastructure
and is never imported here.)
"""
import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcsim_b200.so")

EXPORTS = [
    "mcs_device_count", "mcs_last_error", "mcs_version", "mcs_scene_create", "mcs_scene_destroy",
    "mcs_scene_visit_order", "mcs_scene_prepare_host", "mcs_rollouts", "mcs_rollouts_range", "mcs_reduce_features", "mcs_rollouts_device",
    "mcs_rollouts_groups", "mcs_rollouts_groups_device", "mcs_scene_sync", "mcs_group_partials", "mcs_combine_groups",
    "mcs_trajectories", "mcs_decide_mobil", "mcs_decide_fixed_direction", "mcs_decide_fixed_gap",
    "mcs_decide_closest_gap", "mcs_fp64_peak",
]


class McsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmcsim_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libmcsim_b200.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McsimError(abi.MCS_ERR_CUDA, "%s not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                                            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.mcs_device_count.restype = C.c_int
    L.mcs_last_error.restype = C.c_char_p
    L.mcs_version.restype = C.c_char_p
    L.mcs_scene_create.restype = C.c_int
    L.mcs_scene_create.argtypes = [P(abi.Corridor), C.c_int, P(abi.Agent), C.c_int, C.c_int, P(C.c_void_p)]
    L.mcs_scene_destroy.restype = C.c_int
    L.mcs_scene_destroy.argtypes = [C.c_void_p]
    L.mcs_scene_visit_order.restype = C.c_int
    L.mcs_scene_visit_order.argtypes = [C.c_void_p, P(C.c_int32), C.c_int]
    L.mcs_scene_prepare_host.restype = C.c_int
    L.mcs_scene_prepare_host.argtypes = [P(abi.Corridor), C.c_int, P(abi.Agent), C.c_int] + [P(C.c_int32)] * 5 + [P(C.c_double)]
    L.mcs_rollouts.restype = C.c_int
    L.mcs_rollouts.argtypes = [C.c_void_p, P(abi.Candidate), C.c_int, P(abi.SimParam), C.c_int, C.c_uint64,
                               P(abi.Feature), P(abi.Status)]
    L.mcs_rollouts_range.restype = C.c_int
    L.mcs_rollouts_range.argtypes = [C.c_void_p, P(abi.Candidate), C.c_int, P(abi.SimParam), C.c_uint64, C.c_int64,
                                     C.c_int64, P(abi.Status)]
    L.mcs_reduce_features.restype = C.c_int
    L.mcs_reduce_features.argtypes = [P(abi.Status), C.c_int, C.c_int, C.c_int, C.c_int, P(abi.Feature)]
    L.mcs_rollouts_device.restype = C.c_int
    L.mcs_rollouts_device.argtypes = [C.c_void_p, P(abi.Candidate), C.c_int, P(abi.SimParam), C.c_uint64, C.c_int64,
                                      C.c_int64, C.c_void_p, C.c_void_p, P(C.c_float), P(C.c_uint64)]
    L.mcs_rollouts_groups.restype = C.c_int
    L.mcs_rollouts_groups.argtypes = [C.c_void_p, P(abi.Candidate), C.c_int, P(abi.SimParam), C.c_int, C.c_uint64, C.c_int,
                                      C.c_int, P(abi.GroupPartial)]
    L.mcs_rollouts_groups_device.restype = C.c_int
    L.mcs_rollouts_groups_device.argtypes = [C.c_void_p, P(abi.Candidate), C.c_int, P(abi.SimParam), C.c_int, C.c_uint64,
                                             C.c_int, C.c_int, C.c_void_p]
    L.mcs_scene_sync.restype = C.c_int
    L.mcs_scene_sync.argtypes = [C.c_void_p]
    L.mcs_group_partials.restype = C.c_int
    L.mcs_group_partials.argtypes = [P(abi.Status), C.c_int, C.c_int, C.c_int, C.c_int, P(abi.GroupPartial)]
    L.mcs_combine_groups.restype = C.c_int
    L.mcs_combine_groups.argtypes = [P(abi.GroupPartial), C.c_int, C.c_int, P(abi.Feature)]
    L.mcs_trajectories.restype = C.c_int
    L.mcs_trajectories.argtypes = [C.c_void_p, P(abi.Candidate), P(abi.SimParam), C.c_uint64, C.c_int64,
                                   P(C.c_double), P(C.c_int32), P(abi.Status)]
    L.mcs_decide_mobil.restype = C.c_int
    L.mcs_decide_mobil.argtypes = [C.c_void_p, C.c_int32, P(abi.Action), C.c_int, C.c_uint64, P(abi.ActionWithType)]
    L.mcs_decide_fixed_direction.restype = C.c_int
    L.mcs_decide_fixed_direction.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(abi.Action)]
    L.mcs_decide_fixed_gap.restype = C.c_int
    L.mcs_decide_fixed_gap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, P(abi.Action)]
    L.mcs_decide_closest_gap.restype = C.c_int
    L.mcs_decide_closest_gap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(abi.Action)]
    L.mcs_fp64_peak.restype = C.c_int
    L.mcs_fp64_peak.argtypes = [C.c_int, C.c_int, P(C.c_double), P(C.c_float)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise McsimError(rc, lib().mcs_last_error().decode("utf-8", "replace"))


class DeviceScene:
    """A MapMultiLane + list[Agent] resident on one GPU (mcs_scene_create)."""

    def __init__(self, corridors, agents, device=0):
        self._h = C.c_void_p()
        self.n_agents = len(agents)
        self.n_corridors = len(corridors)
        self._corridors = corridors   # keep the ctypes arrays alive
        self._agents = agents
        check(lib().mcs_scene_create(corridors, len(corridors), agents, len(agents), device, C.byref(self._h)))

    @classmethod
    def adopt(cls, handle, n_agents, n_corridors):
        """Wrap an mcs_scene* another entry point created (mce_build_rollout_scene); this object now owns it."""
        self = cls.__new__(cls)
        self._h, self.n_agents, self.n_corridors, self._corridors, self._agents = handle, n_agents, n_corridors, None, None
        return self

    def close(self):
        if self._h:
            lib().mcs_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def visit_order(self):
        out = (C.c_int32 * self.n_agents)()
        check(lib().mcs_scene_visit_order(self._h, out, self.n_agents))
        return list(out)

    def rollouts(self, candidates, sim_param, n_per_group, seed, want_status=False):
        nc = len(candidates)
        feats = (abi.Feature * nc)()
        st = None
        if want_status:
            st = (abi.Status * (nc * abi.MCS_GROUPS * n_per_group))()
        check(lib().mcs_rollouts(self._h, candidates, nc, C.byref(sim_param), n_per_group, seed, feats, st))
        return feats, st

    def rollouts_range(self, candidates, sim_param, seed, first, count):
        nc = len(candidates)
        st = (abi.Status * (nc * count))()
        check(lib().mcs_rollouts_range(self._h, candidates, nc, C.byref(sim_param), seed, first, count, st))
        return st

    def rollouts_groups(self, candidates, sim_param, n_per_group, seed, first_group, n_groups, d_partials_ptr=None):
        """Groups [first_group, first_group+n_groups) of every candidate -> one GroupPartial per (candidate, group); with
        d_partials_ptr (a device pointer, e.g. tensor.data_ptr()) the partials stay on the GPU for the NCCL exchange."""
        nc = len(candidates)
        if d_partials_ptr is not None:
            check(lib().mcs_rollouts_groups_device(self._h, candidates, nc, C.byref(sim_param), n_per_group, seed, first_group,
                                                   n_groups, d_partials_ptr))
            return None
        out = (abi.GroupPartial * (nc * n_groups))()
        check(lib().mcs_rollouts_groups(self._h, candidates, nc, C.byref(sim_param), n_per_group, seed, first_group, n_groups, out))
        return out

    def rollouts_device(self, candidates, sim_param, seed, first, count, d_status_ptr=None):
        ms = C.c_float()
        steps = C.c_uint64()
        check(lib().mcs_rollouts_device(self._h, candidates, len(candidates), C.byref(sim_param), seed, first, count,
                                        d_status_ptr, None, C.byref(ms), C.byref(steps)))
        return ms.value, steps.value

    def trajectories(self, candidate, sim_param, seed, rollout):
        n = self.n_agents
        buf = (C.c_double * (n * (sim_param.steps + 1) * 4))()
        ns = C.c_int32()
        st = abi.Status()
        check(lib().mcs_trajectories(self._h, C.byref(candidate), C.byref(sim_param), seed, rollout, buf, C.byref(ns),
                                     C.byref(st)))
        return buf, ns.value, st


    # ---- the reference's per-step Python-facing wrappers (behaviors.py:81,114,135,144,152,174)
    def decide_mobil(self, agent_id, ax, vy, require_rss, seed):
        out = abi.ActionWithType()
        act = abi.Action(ax, vy)
        check(lib().mcs_decide_mobil(self._h, agent_id, C.byref(act), 1 if require_rss else 0, seed, C.byref(out)))
        return out

    def decide_fixed_direction(self, agent_id, action):
        out = abi.Action()
        check(lib().mcs_decide_fixed_direction(self._h, agent_id, action, C.byref(out)))
        return out

    def decide_fixed_gap(self, agent_id, direction, gap_id):
        out = abi.Action()
        check(lib().mcs_decide_fixed_gap(self._h, agent_id, direction, gap_id, C.byref(out)))
        return out

    def decide_closest_gap(self, agent_id, direction):
        out = abi.Action()
        check(lib().mcs_decide_closest_gap(self._h, agent_id, direction, C.byref(out)))
        return out


def fp64_peak_tflops(device=0, iters=20000):
    """Measured FP64 multiply-add rate of the device (TFLOP/s, kernel ms): the roofline denominator of the rollout kernels."""
    tf, ms = C.c_double(), C.c_float()
    check(lib().mcs_fp64_peak(device, iters, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def prepare_host(corridors, agents):
    """Host-side scene preparation only (no device): visit order, corridor order, neighbour lane ids, first lane match and
    lane-change prior of every agent."""
    nc, na = len(corridors), len(agents)
    ao, co = (C.c_int32 * na)(), (C.c_int32 * nc)()
    li, ri, fl = (C.c_int32 * nc)(), (C.c_int32 * nc)(), (C.c_int32 * na)()
    pr = (C.c_double * (3 * na))()
    check(lib().mcs_scene_prepare_host(corridors, nc, agents, na, ao, co, li, ri, fl, pr))
    return dict(agent_order=list(ao), corridor_order=list(co), left_ids=list(li), right_ids=list(ri), first_lane_ids=list(fl),
                priors=[tuple(pr[3 * i:3 * i + 3]) for i in range(na)])


def group_partials(statuses, n_candidates, groups, n_per_group, steps):
    out = (abi.GroupPartial * (n_candidates * groups))()
    check(lib().mcs_group_partials(statuses, n_candidates, groups, n_per_group, steps, out))
    return out


def combine_groups(partials, n_candidates, groups):
    feats = (abi.Feature * n_candidates)()
    check(lib().mcs_combine_groups(partials, n_candidates, groups, feats))
    return feats


def reduce_features(statuses, n_candidates, groups, n_per_group, steps):
    feats = (abi.Feature * n_candidates)()
    check(lib().mcs_reduce_features(statuses, n_candidates, groups, n_per_group, steps, feats))
    return feats

"""ctypes mirror of include/mcsim.h (the C ABI of libmcsim_b200.so).

Field order and widths follow the header exactly; the header in turn follows the memory
layout of the reference's value types.
"""
import ctypes as C

# enumerations (reference passes std::string; binding strings in mc_sim_highway.so .rodata)
LANE_TYPES = {"main": 0, "merging": 1, "exit": 2}
LANE_OTHER = 3
BEHAVIORS = {"idm": 0, "idm_lc": 1, "merging": 2, "exit": 3}
BEH_OTHER = 4
ACTIONS = {"keep_lane": 0, "acc": 1, "dcc": 2, "left": 3, "right": 4}
ACT_INVALID = 5
COMMITS = {"not_decided": 0, "keep_lane": 1, "left": 2, "right": 3}
COMMIT_NAMES = {v: k for k, v in COMMITS.items()}

MCS_OK, MCS_ERR_ARG, MCS_ERR_CUDA, MCS_ERR_CAPACITY, MCS_ERR_SCENE = 0, -1, -2, -3, -4
MCS_MAX_AGENTS = 256
MCS_MAX_LANES = 8
MCS_GROUPS = 8


class Corridor(C.Structure):
    _fields_ = [("id", C.c_int32), ("type", C.c_int32), ("left", C.c_double * 4), ("right", C.c_double * 4),
                ("center", C.c_double * 4), ("v_limit", C.c_double)]


class Agent(C.Structure):
    _fields_ = [("id", C.c_int32), ("behavior", C.c_int32),
                ("x", C.c_double), ("y", C.c_double), ("v", C.c_double), ("vy", C.c_double),
                ("vd", C.c_double), ("a", C.c_double), ("l", C.c_double), ("w", C.c_double),
                ("idm", C.c_double * 6), ("rss", C.c_double * 6), ("yp", C.c_double * 7),
                ("commit", C.c_int32), ("commit_keep_lane_step", C.c_int32), ("target_lane_id", C.c_int32),
                ("reserved", C.c_int32)]


class SimParam(C.Structure):
    _fields_ = [("dt", C.c_double), ("steps", C.c_int32), ("min_steps", C.c_int32)]


class Candidate(C.Structure):
    _fields_ = [("ego_id", C.c_int32), ("action", C.c_int32), ("gap_id", C.c_int32), ("reserved", C.c_int32)]


class Feature(C.Structure):
    _fields_ = [("fallBackRate", C.c_double), ("successRate", C.c_double), ("utility", C.c_double),
                ("comfort", C.c_double), ("expectedStepRatio", C.c_double), ("utilityObjects", C.c_double),
                ("comfortObjects", C.c_double), ("fromNSims", C.c_int32), ("reserved", C.c_int32)]


class Status(C.Structure):
    _fields_ = [("utility", C.c_double), ("comfort", C.c_double), ("utilityObjects", C.c_double),
                ("comfortObjects", C.c_double), ("finishStep", C.c_int32), ("executedSteps", C.c_int32),
                ("draws", C.c_uint32), ("finish", C.c_uint8), ("fallBack", C.c_uint8), ("ok", C.c_uint8),
                ("overflow", C.c_uint8), ("reserved", C.c_int32 * 4)]


class GroupPartial(C.Structure):
    _fields_ = [("f", C.c_double * 7), ("fromNSims", C.c_int32), ("reserved", C.c_int32)]


class Action(C.Structure):
    _fields_ = [("ax", C.c_double), ("vy", C.c_double)]


class ActionWithType(C.Structure):
    _fields_ = [("ax", C.c_double), ("vy", C.c_double), ("commit", C.c_int32),
                ("commit_keep_lane_step", C.c_int32), ("target_lane_id", C.c_int32), ("reserved", C.c_int32)]



assert C.sizeof(GroupPartial) == 64
assert C.sizeof(Status) == 64 and C.sizeof(Feature) == 64 and C.sizeof(Agent) == 240 and C.sizeof(Corridor) == 112

/* mcsim.h — C ABI of libmcsim_b200.so, the B200 drop-in for the rollout path of
 * einsteinguang/interactive_highway_simulation's precompiled `mc_sim_highway` module.
 *
 * Every entry point replaces one Python-visible function of that module (binding table
 * in mc_sim_highway.so `init_module_mc_sim_highway` 0xc2bf0; callers in the reference's
 * behaviors.py).  Plain pointers and sizes only; the caller owns every buffer; return
 * value 0 = ok, negative = error (mcs_last_error() gives the text).  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with MCS_ERR_CUDA.
 *
 *   reference (Python name → C++ symbol @ address)                          → here
 *   simulationMultiThread → python_converter::simulationMultiThreadWrapper
 *       0xbddd0 → runMultiThreads 0xbb910 → runMTimes 0xc28c0 → Simulation::run 0x10e790
 *                                                                           → mcs_rollouts*
 *   Simulation(...).run() / agentsStatesHistory()  0x10e790 / 0xef540       → mcs_trajectories
 *   laneChangeMobilBehavior → laneChangeBehaviorMOBILWrapper 0xb9960        → mcs_decide_mobil
 *   fixedDirectionLaneChangeBehavior → customizedLaneChangeBehaviorWrapper  → mcs_decide_fixed_direction
 *   fixedGapMergingBehavior → customizedMergingBehaviorWrapper              → mcs_decide_fixed_gap
 *   closestGapMergingBehavior → closestGapMergingBehaviorWrapper            → mcs_decide_closest_gap
 *   MapMultiLane(list[Corridor]) 0x109f00 + Environment(map, list[Agent]) 0x10c6c0
 *                                                                           → mcs_scene_create
 */
#ifndef MCSIM_H
#define MCSIM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enumerations (the reference passes these as std::string) ---- */
enum { MCS_LANE_MAIN = 0, MCS_LANE_MERGING = 1, MCS_LANE_EXIT = 2, MCS_LANE_OTHER = 3 };
enum { MCS_BEH_IDM = 0, MCS_BEH_IDM_LC = 1, MCS_BEH_MERGING = 2, MCS_BEH_EXIT = 3, MCS_BEH_OTHER = 4 };
/* ego command: "keep_lane" "acc" "dcc" "left" "right"; anything else = INVALID */
enum { MCS_ACT_KEEP_LANE = 0, MCS_ACT_ACC = 1, MCS_ACT_DCC = 2, MCS_ACT_LEFT = 3, MCS_ACT_RIGHT = 4, MCS_ACT_INVALID = 5 };
/* Agent::commitToDecision: "not_decided" "keep_lane" "left" "right" */
enum { MCS_COMMIT_NOT_DECIDED = 0, MCS_COMMIT_KEEP_LANE = 1, MCS_COMMIT_LEFT = 2, MCS_COMMIT_RIGHT = 3 };

enum {
  MCS_OK = 0,
  MCS_ERR_ARG = -1,      /* bad pointer / size / enum */
  MCS_ERR_CUDA = -2,     /* no device, launch or copy failure */
  MCS_ERR_CAPACITY = -3, /* scene larger than the compiled limits (MCS_MAX_*) */
  MCS_ERR_SCENE = -4     /* an agent's lane is unknown where the reference would crash */
};

#define MCS_MAX_AGENTS 256   /* vehicles per rollout  */
#define MCS_MAX_LANES 8      /* corridors per scene   */
#define MCS_GROUPS 8         /* runMultiThreads' hard-coded thread count (0xbb952) */

/* ---- value types (layout = the reference's memory order) ---- */
typedef struct mcs_corridor {
  int32_t id;
  int32_t type;      /* MCS_LANE_* */
  double left[4];    /* Line left:   p1.x p1.y p2.x p2.y  (Corridor+0x38) */
  double right[4];   /* Line right                         (Corridor+0x58) */
  double center[4];  /* Line center                        (Corridor+0x78) */
  double v_limit;    /* Corridor+0x98 */
} mcs_corridor;

typedef struct mcs_agent {
  int32_t id;
  int32_t behavior;              /* MCS_BEH_* */
  double x, y, v, vy, vd, a, l, w; /* Agent ctor order: x y vx vy vd a l w */
  double idm[6];                 /* accMax minDis accMin accCom aEmergency thw (memory order) */
  double rss[6];                 /* rTimeOther rTimeEgo aMinBrake aMaxBrake aSoft dSoft */
  double yp[7];                  /* bias a b c politenessFactor deltaA biasA */
  int32_t commit;                /* MCS_COMMIT_*; Python may set Agent.commitToDecision (behaviors.py:77) */
  int32_t commit_keep_lane_step; /* Agent.commitKeepLaneStep */
  int32_t target_lane_id;        /* Agent.targetLaneId (-1 = none) */
  int32_t reserved;
} mcs_agent;

typedef struct mcs_sim_param { /* SimParameter(dt, steps, minSteps) */
  double dt;
  int32_t steps;
  int32_t min_steps;
} mcs_sim_param;

typedef struct mcs_candidate { /* one simulationMultiThread call: (egoId, action, gapId) */
  int32_t ego_id;
  int32_t action;  /* MCS_ACT_* */
  int32_t gap_id;  /* -2 = n/a, -1 = last gap, else vehicle id */
  int32_t reserved;
} mcs_candidate;

typedef struct mcs_feature { /* BehaviorFeature, same field order as the reference struct */
  double fallBackRate, successRate, utility, comfort, expectedStepRatio, utilityObjects, comfortObjects;
  int32_t fromNSims;
  int32_t reserved;
} mcs_feature;

typedef struct mcs_status { /* StatusOneSim of one rollout + bookkeeping; 64 bytes */
  double utility, comfort, utilityObjects, comfortObjects;
  int32_t finishStep;
  int32_t executedSteps;  /* steps actually simulated (≤ steps) */
  uint32_t draws;         /* random draws consumed, incl. the N aggressive-flag draws */
  uint8_t finish, fallBack, ok, overflow;  /* overflow: 0 fine, 1 yielding table full, 2 reference undefined here.  Every
                                              entry point returns the error code if ANY rollout of its launch is marked
                                              (a launch-wide word the kernel ORs into), whichever rollout it was. */
  int32_t reserved[4];
} mcs_status;

typedef struct mcs_group_partial { /* one runMTimes group = one thread's BehaviorFeature inside runMultiThreads; 64 bytes */
  double f[7];       /* group means: fallBackRate successRate utility comfort expectedStepRatio utilityObjects comfortObjects */
  int32_t fromNSims; /* n_per_group, or 0 when setEgoAgentIdAndDrivingBehavior failed (group skipped by the mean) */
  int32_t reserved;  /* max of mcs_status.overflow over the group's rollouts: a scene-error mark survives the all-gather and
                        makes mcs_combine_groups fail (MCS_ERR_SCENE / MCS_ERR_CAPACITY) on every rank */
} mcs_group_partial;

typedef struct mcs_action { double ax, vy; } mcs_action;

typedef struct mcs_action_with_type { /* ActionWithType */
  double ax, vy;
  int32_t commit;                /* MCS_COMMIT_* */
  int32_t commit_keep_lane_step;
  int32_t target_lane_id;
  int32_t reserved;
} mcs_action_with_type;

typedef struct mcs_scene mcs_scene; /* opaque: map + agents resident in HBM */

/* ---- lifetime ---- */
int mcs_device_count(void);
const char* mcs_last_error(void);
const char* mcs_version(void);

/* MapMultiLane(corridors) + list[Agent] → scene bound to CUDA device `device` (structure-of-arrays).  Fails without a
 * usable device (there is no CPU path).  The single host→device copy of the scene is made at its first use, when the
 * launch has fixed the row length (threads per rollout); the scene then stays resident in HBM. */
int mcs_scene_create(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                     int device, mcs_scene** out);
int mcs_scene_destroy(mcs_scene* scene);
/* agent visit order (std::unordered_map iteration order the reference walks) and neighbour lane ids, for tests */
int mcs_scene_visit_order(const mcs_scene* scene, int32_t* order_out, int n);
/* Host-side half of mcs_scene_create alone (no device needed): what the reference derives once per call —
 * agent visit order (input indices), corridor iteration order (input indices), per input corridor the id of its left /
 * right neighbour (INT32_MIN = none; MapMultiLane ctor 0x109f00), per input agent the id of the lane of its first match
 * (INT32_MIN = none; matchAgentsOnCorridor 0x10b810) and its lane-change prior {left, keep, right} (0x10bb97). */
int mcs_scene_prepare_host(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                           int32_t* agent_order_out, int32_t* corridor_order_out, int32_t* left_ids_out,
                           int32_t* right_ids_out, int32_t* first_lane_ids_out, double* priors_out /* [n_agents][3] */);

/* ---- simulationMultiThread, batched over candidates ----
 * For each candidate c: MCS_GROUPS groups × n_per_group rollouts; rollout r (0-based over the
 * candidate's groups*n_per_group rollouts, offset by first_rollout) draws from the stream
 * mcs_rollout_key(seed + c, r).  features_out[c] = mean over groups of the per-group means,
 * exactly as runMTimes/runMultiThreads combine them.  Host buffers in and out. */
int mcs_rollouts(mcs_scene* scene, const mcs_candidate* candidates, int n_candidates, const mcs_sim_param* param,
                 int n_per_group, uint64_t seed, mcs_feature* features_out, mcs_status* status_out /* may be NULL;
                 else n_candidates*MCS_GROUPS*n_per_group records */);

/* Sharded form: simulate only rollouts [first, first+count) of every candidate and return their
 * statuses (count per candidate); the caller (one rank per GPU) gathers and calls mcs_reduce_features. */
int mcs_rollouts_range(mcs_scene* scene, const mcs_candidate* candidates, int n_candidates, const mcs_sim_param* param,
                       uint64_t seed, int64_t first, int64_t count, mcs_status* status_out);
/* Fixed-order two-level mean (host, pure function): statuses = n_candidates × (groups*n_per_group). */
int mcs_reduce_features(const mcs_status* statuses, int n_candidates, int groups, int n_per_group, int steps,
                        mcs_feature* features_out);

/* Multi-GPU form of simulationMultiThread: the 8 groups of runMultiThreads are the shard unit.  A rank simulates groups
 * [first_group, first_group+n_groups) of every candidate (rollouts first_group*n_per_group ...) and gets one partial
 * per (candidate, group): partials_out[c*n_groups + g].  The ranks all-gather their partials (64 B each — the only
 * data that crosses NVLink) and every rank calls mcs_combine_groups, which adds them in group order exactly like
 * runMultiThreads 0xbca08-0xbca92: the result is bit-identical to the single-GPU mcs_rollouts.
 * mcs_rollouts_groups_device leaves the partials in device memory (d_partials_out, e.g. a torch tensor that is then
 * handed to NCCL); it returns after the device work has finished, so the buffer may be used from any stream. */
int mcs_rollouts_groups(mcs_scene* scene, const mcs_candidate* candidates, int n_candidates, const mcs_sim_param* param,
                        int n_per_group, uint64_t seed, int first_group, int n_groups, mcs_group_partial* partials_out);
int mcs_rollouts_groups_device(mcs_scene* scene, const mcs_candidate* candidates, int n_candidates,
                               const mcs_sim_param* param, int n_per_group, uint64_t seed, int first_group, int n_groups,
                               void* d_partials_out);
int mcs_scene_sync(mcs_scene* scene);
/* host, pure: statuses[n_candidates][groups*n_per_group] -> partials[n_candidates][groups] (runMTimes 0xc2a6e-0xc2b1f) */
int mcs_group_partials(const mcs_status* statuses, int n_candidates, int groups, int n_per_group, int steps,
                       mcs_group_partial* partials_out);
/* host, pure: partials[n_candidates][groups] -> features (runMultiThreads 0xbca08-0xbca92, fixed group order) */
int mcs_combine_groups(const mcs_group_partial* partials, int n_candidates, int groups, mcs_feature* features_out);

/* Device-resident variant used by the throughput benchmark: results stay in HBM
 * (d_status_out = device pointer or NULL to keep only the per-block vehicle-step counters).
 * Returns total executed vehicle-steps through *vehicle_steps_out after synchronising; *kernel_ms_out = the rollout
 * kernel's duration from CUDA events on the stream it was launched on.  `cuda_stream` is reserved (pass NULL): all work
 * of a device goes to the library's own stream. */
int mcs_rollouts_device(mcs_scene* scene, const mcs_candidate* candidates, int n_candidates, const mcs_sim_param* param,
                        uint64_t seed, int64_t first, int64_t count, void* d_status_out, void* cuda_stream,
                        float* kernel_ms_out, uint64_t* vehicle_steps_out);

/* Measurement helper (no reference counterpart): the FP64 rate of device `device` on a stream of independent
 * fused multiply-adds — what a CUDA-core FP64 kernel like the rollout kernel can reach at best; the denominator of
 * bench.py's `roofline` (FP64 pipe, not HBM, bounds this path).  iters x 64 DFMA per thread. */
int mcs_fp64_peak(int device, int iters, double* tflops_out, float* ms_out);

/* ---- Simulation.run() + agentsStatesHistory() for one rollout ----
 * states_out: [n_agents][steps+1][4] (x y v a), agent order = input order; n_states_out = states per agent */
int mcs_trajectories(mcs_scene* scene, const mcs_candidate* candidate, const mcs_sim_param* param, uint64_t seed,
                     int64_t rollout, double* states_out, int32_t* n_states_out, mcs_status* status_out);

/* ---- single decisions (the reference's per-step Python-facing wrappers) ---- */
int mcs_decide_mobil(mcs_scene* scene, int32_t agent_id, const mcs_action* idm_action, int require_rss,
                     uint64_t seed, mcs_action_with_type* out);
int mcs_decide_fixed_direction(mcs_scene* scene, int32_t agent_id, int32_t action, mcs_action* out);
int mcs_decide_fixed_gap(mcs_scene* scene, int32_t agent_id, int32_t direction, int32_t gap_id, mcs_action* out);
int mcs_decide_closest_gap(mcs_scene* scene, int32_t agent_id, int32_t direction, mcs_action* out);

#ifdef __cplusplus
}
#endif
#endif /* MCSIM_H */

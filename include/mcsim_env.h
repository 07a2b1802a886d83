/* mcsim_env.h — C ABI of the OUTER environment's lane match, per-lane ordering, leader / follower / side-neighbour search
 * and Frenet conversion (libmcsim_b200.so, csrc/env_capi.cu + env_kernels.cuh).
 *
 * Replaces, for all vehicles of a scene (and for a batch of scenes) in one launch, what the reference's Python does
 * per vehicle with shapely (SURVEY.md §8 rows a9/a10 Python twins, a13, a14):
 *   environment.py:40-44    Map.corridor_match            first corridor (dict order) whose polygon holds the point
 *   environment.py:85-101   match_agents_on_corridor      lane of every vehicle; unmatched vehicles leave the scene
 *   environment.py:135-143  agents_in_corridor_by_arc_length   per-lane stable sort on the centre-line arc length
 *   environment.py:166-195  front_agent / following_agent
 *   environment.py:197-251  neighbor_around_agent         the 8 side neighbours, on the +100 m extended centre line
 *   environment.py:145-156  sorted_agents_on_corridor_within_range_of_vehicle (from arc_on_lane + lane_order)
 *   type.py:297-317         LineSimplified.signed_distance / to_frenet_frame on extend_line_both_sides(center, 1000)
 *   utility.py:88-98        step_along_path                (mce_step_along_path)
 *   type.py:180-206,307-315 distance_to_corridor_end / distance_to_border / centerSimplified.signed_distance
 *                           (mce_center_distance, mce_border_distance: the geometry idm_behavior and the yielding model read)
 *
 * Plain pointers and sizes; caller owns all buffers; HOST buffers in, HOST buffers out (copies inside; every query entry
 * stages its arguments through one pinned buffer of the handle: one H2D and one D2H copy per call).  Calls on one handle
 * must not run concurrently (use one handle per thread); different handles are independent.  No CPU
 * fallback: without a CUDA device every compute entry fails with MCE_ERR_CUDA.
 * The keys of the per-lane vectors stay resident in the handle after mce_env_match, so that later queries for vehicles
 * that have moved since (the reference plans and moves its vehicles one after the other against the arc lengths of the
 * last match, environment.py:316-323) are answered against the same stale keys (mce_env_neighbors).
 */
#ifndef MCSIM_ENV_H
#define MCSIM_ENV_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { MCE_OK = 0, MCE_ERR_ARG = -1, MCE_ERR_CUDA = -2, MCE_ERR_CAPACITY = -3 };

#define MCE_MAX_LANES 8
#define MCE_MAX_POINTS 16   /* points per polyline of a corridor          */
#define MCE_MAX_AGENTS 1024 /* vehicles per scene                         */
#define MCE_NEIGHBORS 10    /* Neighbor(...) ctor order, agent.py:203-213 */
#define MCE_BUILD_MAX_AGENTS 48 /* vehicles one builder call can marshal (the reference's selection gives <= ~25) */

/* index into the neighbour tables, the order of the reference's Neighbor constructor */
enum { MCE_LEFT_FF = 0, MCE_LEFT_F, MCE_LEFT_B, MCE_LEFT_BB, MCE_FRONT, MCE_BACK, MCE_RIGHT_FF, MCE_RIGHT_F, MCE_RIGHT_B,
       MCE_RIGHT_BB };

/* type.py:140-169 Corridor(idx, t, left, right, center, v_limit) + set_left_and_right_corridor_id.
 * left_index / right_index: position of the neighbour corridor in the array passed to mce_map_create, -1 = None. */
typedef struct mce_corridor {
  int32_t id, left_index, right_index;
  int32_t n_left, n_right, n_center;
  double left[MCE_MAX_POINTS][2], right[MCE_MAX_POINTS][2], center[MCE_MAX_POINTS][2];
} mce_corridor;

typedef struct mce_map mce_map; /* opaque: corridors of one map resident on a device + per-device work buffers */

const char* mce_last_error(void);

/* Map ctor environment.py:21-32 (corridor order = dict insertion order = array order) plus what the Corridor ctor
 * derives: polygon border (left + reversed right, type.py:160-164), centerExtended (extend_line(center, 100),
 * type.py:154, utility.py:111-115), and the Frenet reference line of mcs_wrapper.py:42,109,211
 * (extend_line_both_sides(center, 1000) utility.py:118-126, simplified with tolerance 0.05, type.py:257-272). */
int mce_map_create(const mce_corridor* corridors, int n_corridors, int device, mce_map** out);
int mce_map_destroy(mce_map* map);

/* match_agents_on_corridor + every neighbour query of one environment state, for n_scenes scenes of n_agents slots
 * each on the same map (x, y: [n_scenes][n_agents]; a slot with NaN x is empty).
 *   lane_out      [n_scenes][n_agents]                 corridor index, -1 = not on any corridor (vehicle removed)
 *   arc_out       [n_scenes][n_agents]                 arc length on the own lane's centre line
 *   lane_order_out[n_scenes][n_agents]                 slots grouped by lane (lane 0 first), each lane sorted by arc
 *                                                       length, ties in slot order (stable sort); -1 padding at the end
 *   lane_start_out[n_scenes][MCE_MAX_LANES + 1]        offsets of the lanes inside lane_order_out
 *   arc_on_lane_out [n_scenes][n_corridors][n_agents]  centre-line arc length of every vehicle on every corridor (may be NULL)
 *   nb_index_out  [n_scenes][MCE_NEIGHBORS][n_agents]  slot of the neighbour, -1 = None
 *   nb_dis_out    [n_scenes][MCE_NEIGHBORS][n_agents]  IdmMatch.dis (key - own arc on that lane's line)
 * Any output pointer may be NULL. */
int mce_env_match(mce_map* map, const double* x, const double* y, int n_agents, int n_scenes, int32_t* lane_out,
                  double* arc_out, int32_t* lane_order_out, int32_t* lane_start_out, double* arc_on_lane_out,
                  int32_t* nb_index_out, double* nb_dis_out);

/* front_agent / following_agent / neighbor_around_agent for n_queries vehicles (scene[i], slot[i]) at their CURRENT
 * positions (qx, qy) against the lanes and keys of the last mce_env_match on this map handle.
 *   nb_index_out [MCE_NEIGHBORS][n_queries], nb_dis_out likewise. */
int mce_env_neighbors(mce_map* map, const int32_t* scene, const int32_t* slot, const double* qx, const double* qy,
                      int n_queries, int32_t* nb_index_out, double* nb_dis_out);

/* LineSimplified(extend_line_both_sides(corridors[ref_lane].center, 1000)).to_frenet_frame([x, y]) for n points. */
int mce_frenet(mce_map* map, int ref_lane, const double* x, const double* y, int n, double* s_out, double* d_out);

/* corridors[lane].centerLine.project(Point([x, y])) for n points (environment.py:139,149-152,160-161): arc length and
 * distance to the (unextended, unsimplified) centre line. */
int mce_project_center(mce_map* map, int lane, const double* x, const double* y, int n, double* s_out, double* dist_out);

/* utility.py:88-98 step_along_path on the centre line of corridors[ref_lane[i]] (DecoupledAction.ref_path is always a
 * corridor's centerSimplified, behaviors.py): new x, y, v, yaw of n vehicles.  cos / sin of the segment yaws are
 * evaluated once on the host at mce_map_create. */
int mce_step_along_path(mce_map* map, const int32_t* ref_lane, const double* x, const double* y, const double* v,
                        const double* ax, const double* vy, double dt, int n, double* x_out, double* y_out,
                        double* v_out, double* yaw_out);

/* corridors[lane[i]].centerSimplified.signed_distance([x, y]) (type.py:307-315; the lateral "attach to the centre
 * line" term of idm_behavior, behaviors.py:20-22) and Corridor.distance_to_corridor_end([x, y]) (type.py:180-181; the
 * yielding features d_to_merging_lane_end, type.py:126-127) for n points, each on its own corridor. */
int mce_center_distance(mce_map* map, const int32_t* lane, const double* x, const double* y, int n, double* signed_out,
                        double* to_end_out);

/* Environment.dis_to_lane_border (environment.py:113-115) = Corridor.distance_to_border(hull, direction)
 * (type.py:183-206) for n vehicle hulls (hull: [n][4][2], the corner order of vehicle_pose_to_rect utility.py:12-20):
 * minus the largest distance by which a corner has passed the border on that side, else the distance of the hull polygon
 * to the border line (the yielding feature d_0_lat, type.py:123). */
enum { MCE_SIDE_LEFT = 0, MCE_SIDE_RIGHT = 1 };
int mce_border_distance(mce_map* map, const int32_t* lane, const int32_t* side, const double* hull, int n, double* out);

/* ---- mcs_wrapper.py:24-283 on the device: the three builders of the inner simulator's scene ------------------------
 * Corridor.id / .type (MCS_LANE_*) / .v_limit of every corridor of the map, in the order of mce_map_create — the values
 * the builders copy into the CorridorMS records and the exit builder's lane arithmetic (mcs_wrapper.py:96-118). */
struct mcs_corridor;
struct mcs_agent;
struct mcs_scene;
int mce_map_set_lane_info(mce_map* map, const int32_t* ids, const int32_t* types, const double* v_limits, int n_corridors);

typedef struct mce_build_request {
  int32_t kind;            /* 0 merging_mcs_sim_from_env, 1 exit_mcs_sim_from_env, 2 lane_change_mcs_sim_from_env */
  int32_t scene, ego_slot; /* against the tables of the last mce_env_match on this handle */
  int32_t ego_behavior;    /* MCS_BEH_* of the ego's AgentMS ("merging" / "idm" / "exit", mcs_wrapper.py:45,187,277) */
  int32_t others_merging;  /* merging builder: behaviour of the collected main-lane vehicles, 1 "merging" / 0 "idm_lc" (:50) */
  int32_t ego_commit, ego_keep_step, ego_target_lane; /* what behaviors.py:75-78 pokes into the ego's AgentMS */
  double ego_vd;
  double ego_idm[6], ego_rss[6], ego_yp[7];           /* memory order of mcs_agent */
} mce_build_request;

/* One builder call on the device: neighbour selection (FF / F / B / BB per side, front, front of the front, follower,
 * +-80 m on the second lanes) over the resident match tables at the vehicles' CURRENT positions, XY -> Frenet of the
 * builder's reference line, parameter assignment, candidate gap ids — all in one launch; the records come back in the
 * reference's append order, and (scene_out != NULL) the marshalled scene is created on `device` right away
 * (mcs_scene_create), ready for mcs_rollouts / mcs_decide_*.
 *   ids    [n_agents]        Agent.id of every slot of the matched scene
 *   attrs  [8][n_agents]     x y v vy a l w vd (idm_param.vd) of every slot NOW (vehicles moved since the match included)
 *   noise  [n_noise]         np.random.normal(0., 1) draws added to the vd of the non-ego vehicles, in append order; NULL:
 *                            the records carry the plain vd and the caller, who only now knows how many vehicles were
 *                            collected, draws *n_agents_out - 1 values and adds them itself (what the Python layer does:
 *                            the reference draws one value per appended vehicle from numpy's global stream)
 *   corridors_out [MCE_MAX_LANES], agents_out [MCE_BUILD_MAX_AGENTS], actions_out [4]   (host; any may be NULL) */
int mce_build_rollout_scene(mce_map* map, const mce_build_request* request, const int32_t* ids, const double* attrs,
                            const double* noise, int n_noise, struct mcs_corridor* corridors_out, int32_t* n_corridors_out,
                            struct mcs_agent* agents_out, int32_t* n_agents_out, int32_t* actions_out, int32_t* n_actions_out,
                            int device, struct mcs_scene** scene_out);

/* ---- environment.py:316-323 on the device: the sequential plan-and-move walk over the vehicles of a scene ------------
 * One launch walks slots [first, last) in dict order against the tables of the last mce_env_match (csrc/walk_kernel.cuh):
 * idm_behavior (behaviors.py:10-67, yielding intentions included), the MOBIL / closest-gap behaviours through the device
 * builder and the single-step decision, Agent.step.  It stops at the first vehicle that needs the host (behaviour
 * MCE_WB_HOST — the learned behaviours — or a yielding seed of 0) and says where; the host handles that vehicle and calls
 * again behind it. */
enum { MCE_WB_HOST = 0, MCE_WB_IDM = 1, MCE_WB_MOBIL = 2, MCE_WB_MOBIL_SAFE = 3, MCE_WB_MERGE_CLOSEST = 4, MCE_WB_EXIT_CLOSEST = 5 };
enum { MCE_WF_MERGING_MODEL = 1, MCE_WF_EXIT_MODEL = 2 }; /* "merging" / "exit" in Agent.behavior_model */
#define MCE_YIELD_CAP 12
typedef struct mce_walk_vehicle { /* agent.py Agent as far as plan / step read and write it */
  int32_t id, behavior /* MCE_WB_* */, flags /* MCE_WF_* */, seed /* random_yielding_seed */;
  int32_t commit /* MCS_COMMIT_* */, keep_step, target_lane, yi_n;
  double x, y, v, vy, a, yaw, l, w;
  double hull[8];   /* vehicle_pose_to_rect corner order */
  double idm[6];    /* acc_max acc_min min_dis acc_com thw vd */
  double rss[6];    /* r_t_ego r_t_other a_min a_max soft_acc soft_dcc */
  double mobil[3];  /* politeness_factor delta_a bias_a */
  double yp[13];    /* yielding model weights */
  double threshold; /* change_intention_threshold */
  double ax_last;   /* out: the acceleration planned in the last walk */
  int32_t yi_id[MCE_YIELD_CAP];       /* yielding_intention_to_others: other vehicle's id, */
  int32_t yi_yielding[MCE_YIELD_CAP]; /* current intention, */
  double yi_p0[MCE_YIELD_CAP];        /* first probability */
} mce_walk_vehicle;
typedef struct mce_walk_result {
  int32_t stopped_at;  /* slot that needs the host, -1 = reached `last` */
  int32_t status;      /* 0 ok; 100 + builder status, 200 + decision status, 300 intention table full */
  int32_t noise_used;  /* np.random.normal(0., 1) draws consumed from `noise` */
  int32_t mobil_calls; /* laneChangeMobilBehavior calls made: drop-in seeds seed_base .. seed_base + mobil_calls - 1 */
  int32_t random_used, random_seed; /* Python `random` consumed; every consumer is random.seed(s); random.random(): last s */
  int32_t moved, reserved;
} mce_walk_result;
/* vehicles [n_agents of the last match]: in / out (host).  upload != 0: the array goes to the device first (first call of a
 * step, or after the host changed a vehicle); otherwise the walk continues on the device's copy.  noise: the block of
 * np.random.normal(0., 1) draws the builders consume in order; random_first [100]: random.seed(s); random.random() for
 * s = 0..99.  ids [n_agents]. */
int mce_walk_step(mce_map* map, mce_walk_vehicle* vehicles, int upload, const int32_t* ids, const double* noise, int n_noise,
                  const double* random_first, int first, int last, double dt, uint64_t seed_base, mce_walk_result* result);
/* std::unordered_map<int, T> iteration order after emplacing ids[0..n) in order (n <= 64), as csrc/walk_kernel.cuh restates
 * it for the device (scene.cpp uses the container itself); order_out[k] = input index of the k-th node.  For tests. */
int mce_hash_iteration_order(const int32_t* ids, int n, int32_t* order_out);

/* average duration (ms, CUDA events on the launching stream) of the kernels of the last mce_env_match call:
 * out[0] = match + order kernel, out[1] = neighbour kernel */
int mce_last_kernel_ms(mce_map* map, float* out2);

#ifdef __cplusplus
}
#endif
#endif

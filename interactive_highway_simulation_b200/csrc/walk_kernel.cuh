// The outer simulator's sequential step on the device (environment.py:316-323): ONE block walks the vehicles of a scene in
// dict order — plan (behaviors.py), move (agent.py:193-208), against the lane vectors of the last match — instead of the
// host issuing three launch + synchronise round trips per vehicle.
//
// Per vehicle, exactly what the reference's Python does one after the other:
//   idm_behavior (behaviors.py:10-67)      IDM to the leader with the Python formula (utility.py:23-49), lateral attach to
//                                          the centre line, yielding intentions (type.py:70-137: 12-feature logistic,
//                                          hysteresis, Python's `random` re-seeded per vehicle) towards ramp vehicles on the
//                                          right and exiting vehicles on the left, softened IDM to those it yields to, RSS brake
//   lane_change_behavior_mobil (:70-86)    + mcs_wrapper's builder (env_kernels.cuh build_scene_block) -> the inner simulator's
//   merging_behavior_closest_gap (:140-145)  constructor bookkeeping (visit order, first lane match, lane vectors: scene.cpp
//   exit_behavior_closest_gap (:148-153)     restated below) -> the single-step decision (decide_kernel.cuh decide_eval)
//   Agent.step (agent.py:193-208)          step_along_path on the corridor's centre line, hull
// Vehicles whose behaviour needs rollouts (the learned ones), whose yielding seed is 0 (their intentions would draw from the
// un-re-seeded global `random` state) or whose behaviour the kernel does not know end the walk: the host handles that
// vehicle with the Python path and continues the walk behind it.
//
// Randomness: the reference consumes three host streams here, and the walk consumes the same values in the same order —
//   numpy's global stream   one normal(0, 1) per marshalled neighbour: a block drawn ahead by the host, a cursor here; the host
//                           rewinds its stream to the number consumed
//   Python's `random`       every consumer re-seeds with the vehicle's yielding seed and draws once (type.py:75-77,
//                           utility.py:136-138): a 100-entry table of random.seed(s); random.random(); the last seed used
//                           goes back so that the host leaves its stream where the reference would
//   the drop-in's seeds     one per laneChangeMobilBehavior call (mc_sim_highway._next_seed): base + cursor
#pragma once
#include "decide_kernel.cuh"
#include "env_kernels.cuh"

namespace mcw {

enum { WB_HOST = 0, WB_IDM = 1, WB_MOBIL = 2, WB_MOBIL_SAFE = 3, WB_MERGE_CLOSEST = 4, WB_EXIT_CLOSEST = 5 };
enum { WF_MERGING_MODEL = 1, WF_EXIT_MODEL = 2 };
constexpr int kYieldCap = 12;

struct WalkVehicle {               // one vehicle of the outer scene; resident on the device between steps
  int32_t id, behavior, flags, seed;
  int32_t commit, keep_step, target_lane, yi_n;
  double x, y, v, vy, a, yaw, l, w;
  double hull[8];                  // vehicle_pose_to_rect corner order: head-left, head-right, tail-right, tail-left
  double idm[6];                   // acc_max acc_min min_dis acc_com thw vd
  double rss[6];                   // r_t_ego r_t_other a_min a_max soft_acc soft_dcc
  double mobil[3];                 // politeness_factor delta_a bias_a
  double yp[13];                   // yielding model weights
  double threshold;                // change_intention_threshold
  double ax_last;                  // the planned acceleration of the last step (diagnostics)
  int32_t yi_id[kYieldCap];        // yielding intentions: id of the other vehicle, first probability, current intention
  int32_t yi_yielding[kYieldCap];
  double yi_p0[kYieldCap];
};

struct WalkParams {
  int first, last;                 // slots [first, last) in dict order
  int scene;                       // scene of the matched batch
  double dt;
  uint64_t seed_base;              // drop-in seed of the next laneChangeMobilBehavior call
  int n_noise;
};
struct WalkResult {
  int32_t stopped_at;              // slot that needs the host (-1: the walk reached `last`)
  int32_t status;                  // 0 ok, else 100 + builder status / 200 + decision status / 300 capacity of the intention table
  int32_t noise_used, mobil_calls; // cursors of the numpy block and of the drop-in's seed sequence
  int32_t random_used, random_seed;   // whether Python's `random` was consumed, and the seed of its last consumer
  int32_t moved;                   // vehicles moved
  int32_t pad;
};

// ---- iteration order of a libstdc++ std::unordered_map<int, T> after emplace(ids[0]), emplace(ids[1]), ... (unique keys).
// The reference walks its agents (and corridors) in that order; scene.cpp replays it with the container itself, here the
// container's list / bucket discipline is restated: a node whose bucket is empty goes to the head of the list, otherwise
// right behind its bucket's "before" node (the head of the bucket's run); growth 13 -> 29 -> 59 -> 127 buckets re-links
// every node in list order by the same rule (_M_rehash_aux, unique keys).  order[k] = input index of the k-th node.
__host__ __device__ inline void hash_iteration_order(const int32_t* ids, int n, int* order) {
  // singly linked list over input indices; head = first node; bucket_before[b]: -2 = empty, -1 = list head sentinel, else node
  int next[64];
  int bucket_before[127];
  int head = -1, n_bkt = 13, bbegin_bkt = -1;
  // std::hash<int> is the identity on the sign-extended value; ids are non-negative in every scene of the reference, and a
  // 32-bit remainder by one of the four bucket counts is a multiply and a shift (the 64-bit remainder is a ~100-instruction
  // routine on the device, once per emplace and re-link of the walk's serial path)
  auto bkt_of = [&](int idx, int nb) {
    const int id = ids[idx];
    if (id < 0) return (int)((unsigned long long)(long long)id % (unsigned long long)nb);
    const unsigned u = (unsigned)id;
    return (int)(nb == 13 ? u % 13u : nb == 29 ? u % 29u : nb == 59 ? u % 59u : u % 127u);
  };
  for (int b = 0; b < 127; ++b) bucket_before[b] = -2;
  auto link = [&](int node, int nb) {
    const int b = bkt_of(node, nb);
    if (bucket_before[b] == -2) {
      next[node] = head;
      head = node;
      if (next[node] >= 0) bucket_before[bbegin_bkt] = node;    // the old first node's bucket now starts behind `node`
      bucket_before[b] = -1;
      bbegin_bkt = b;
    } else {
      const int before = bucket_before[b];
      if (before == -1) { next[node] = head; head = node; }
      else { next[node] = next[before]; next[before] = node; }
    }
  };
  for (int k = 0; k < n; ++k) {
    if (k + 1 > n_bkt) {                                        // _M_need_rehash: grow before inserting element k + 1
      const int nb = n_bkt == 13 ? 29 : (n_bkt == 29 ? 59 : 127);
      int cur = head;
      for (int b = 0; b < 127; ++b) bucket_before[b] = -2;
      head = -1; bbegin_bkt = -1;
      while (cur >= 0) { const int nx = next[cur]; link(cur, nb); cur = nx; }
      n_bkt = nb;
    }
    link(k, n_bkt);
  }
  int cur = head, q = 0;
  while (cur >= 0) { order[q++] = cur; cur = next[cur]; }
}

// ---- what the inner simulator's constructors derive from the marshalled records (scene.cpp build_scene_host: MapMultiLane
// 0x109f00, Environment 0x10c6c0 with its first matchAgentsOnCorridor, Agent 0xf8380) for a single-step decision: lanes in
// hash order with their neighbours, vehicles in visit order, first lane match with the behaviour switch, softened IDM,
// per-lane vectors sorted by x.  The lane-change prior is not needed (the Python-facing wrappers never blend it in).
// Three phases of a block (the caller synchronises between them); F / I: row stride 64.
// phase 1 (one thread): lanes and the two iteration orders
__device__ inline void prep_orders(const mce::BuildOut& b, double* F, int32_t* I, uint8_t* lane_vec, mcs::SceneDev& sc, int* visit) {
  const int nc = b.n_corridors, na = b.n_agents;
  int lane_ids[MCS_MAX_LANES], lorder[MCS_MAX_LANES];
  for (int k = 0; k < nc; ++k) lane_ids[k] = b.corridors[k].id;
  hash_iteration_order(lane_ids, nc, lorder);
  sc.n = na; sc.n_pad = 64; sc.n_lanes = nc; sc.has_exit_lanes = 0;
  sc.f = F; sc.i = I; sc.lane_vec = lane_vec; sc.thread_slot = lane_vec;
  for (int k = 0; k < 64; ++k) lane_vec[k] = 0;
  int leftmost = -1;
  for (int k = 0; k < nc; ++k) {
    const mcs_corridor& c = b.corridors[lorder[k]];
    mcs::LaneDev& L = sc.lanes[k];
    L.leftY = c.left[1]; L.rightY = c.right[1]; L.centerY = c.center[1]; L.vLimit = c.v_limit; L.endX = c.center[2]; L.startX = c.center[0];
    L.id = c.id; L.type = c.type; L.leftIdx = -1; L.rightIdx = -1; L.leftmost = 0; L.nb = 0;
    // neighbours by centre y: the nearest lane above / below
    double best_up = 0.0, best_dn = 0.0;
    for (int j = 0; j < nc; ++j) {
      const double cy = b.corridors[lorder[j]].center[1];
      if (cy > c.center[1] && (L.leftIdx < 0 || cy < best_up)) { L.leftIdx = j; best_up = cy; }
      if (c.center[1] > cy && (L.rightIdx < 0 || cy > best_dn)) { L.rightIdx = j; best_dn = cy; }
    }
    if (c.type == MCS_LANE_EXIT) sc.has_exit_lanes = 1;
    if (L.leftIdx < 0) leftmost = k;
  }
  if (leftmost >= 0) sc.lanes[leftmost].leftmost = 1;
  mcs::lane_neighbour_bits(sc.lanes, nc);
  int32_t ids[64];
  for (int k = 0; k < na; ++k) ids[k] = b.agents[k].id;
  hash_iteration_order(ids, na, visit);
}
// phase 2 (thread k < 64): the vehicle of slot k
__device__ inline void prep_vehicle(const mce::BuildOut& b, double* F, int32_t* I, const mcs::SceneDev& sc, const int* visit, int k) {
  constexpr int np = 64;
  if (k >= b.n_agents) return;
  const mcs_agent& a = b.agents[visit[k]];
  F[mcs::F_X * np + k] = a.x; F[mcs::F_Y * np + k] = a.y; F[mcs::F_V * np + k] = a.v; F[mcs::F_VY * np + k] = a.vy; F[mcs::F_A * np + k] = a.a;
  F[mcs::F_L * np + k] = a.l; F[mcs::F_W * np + k] = a.w; F[mcs::F_VD * np + k] = a.vd;
  F[mcs::F_IDM_ACCMAX * np + k] = a.idm[0]; F[mcs::F_IDM_MINDIS * np + k] = a.idm[1]; F[mcs::F_IDM_ACCMIN * np + k] = a.idm[2];
  F[mcs::F_IDM_ACCCOM * np + k] = a.idm[3]; F[mcs::F_IDM_THW * np + k] = a.idm[5];
  F[mcs::F_IDMY_ACCMAX * np + k] = (7.0 > a.l) ? a.idm[0] - 0.8 : a.idm[0];
  F[mcs::F_IDMY_ACCMIN * np + k] = (7.0 > a.l) ? 1.0 + a.idm[2] : a.idm[2];
  F[mcs::F_RSS_RTOTHER * np + k] = a.rss[0]; F[mcs::F_RSS_RTEGO * np + k] = a.rss[1]; F[mcs::F_RSS_AMIN * np + k] = a.rss[2];
  F[mcs::F_RSS_AMAX * np + k] = a.rss[3]; F[mcs::F_RSS_DSOFT * np + k] = a.rss[5];
  F[mcs::F_YP_BIAS * np + k] = a.yp[0]; F[mcs::F_YP_A * np + k] = a.yp[1]; F[mcs::F_YP_B * np + k] = a.yp[2]; F[mcs::F_YP_C * np + k] = a.yp[3];
  F[mcs::F_YP_PF * np + k] = a.yp[4]; F[mcs::F_YP_DA * np + k] = a.yp[5]; F[mcs::F_YP_BA * np + k] = a.yp[6];
  F[mcs::F_PRIOR_L * np + k] = 0.0; F[mcs::F_PRIOR_K * np + k] = 0.0; F[mcs::F_PRIOR_R * np + k] = 0.0;
  I[mcs::I_ID * np + k] = a.id; I[mcs::I_INPUT * np + k] = visit[k]; I[mcs::I_BEH * np + k] = a.behavior; I[mcs::I_BEH0 * np + k] = a.behavior;
  I[mcs::I_COMMIT * np + k] = a.commit; I[mcs::I_KEEPSTEP * np + k] = a.commit_keep_lane_step; I[mcs::I_TARGETLANE * np + k] = a.target_lane_id;
  int lane = -1;
  for (int j = 0; j < sc.n_lanes; ++j)
    if (sc.lanes[j].leftY >= a.y && a.y >= sc.lanes[j].rightY) { lane = j; break; }
  I[mcs::I_LANE * np + k] = lane;
  if (lane >= 0) {
    if (sc.lanes[lane].type == MCS_LANE_MAIN && a.behavior == MCS_BEH_MERGING) I[mcs::I_BEH * np + k] = MCS_BEH_IDM_LC;
    if (sc.lanes[lane].type == MCS_LANE_EXIT && a.behavior == MCS_BEH_EXIT) I[mcs::I_BEH * np + k] = MCS_BEH_IDM;
  }
}
// phase 3 (thread k < 64): per-lane vectors — members in visit order, stable by ascending x: position = members of the
// lanes before + members of the own lane with a smaller (x, slot)
__device__ inline void prep_lane_vectors(const mce::BuildOut& b, const double* F, const int32_t* I, uint8_t* lane_vec, mcs::SceneDev& sc, int k) {
  constexpr int np = 64;
  const int na = b.n_agents;
  if (k <= MCS_MAX_LANES) {
    int c = 0;
    for (int j = 0; j < na; ++j) { const int lj = I[mcs::I_LANE * np + j]; c += (lj >= 0 && lj < k) ? 1 : 0; }
    sc.lane_start[k] = c;
  }
  if (k >= na) return;
  const int lane = I[mcs::I_LANE * np + k];
  if (lane < 0) return;
  const double xk = F[mcs::F_X * np + k];
  int pos = 0;
  for (int j = 0; j < na; ++j) {
    const int lj = I[mcs::I_LANE * np + j];
    if (lj < 0) continue;
    if (lj < lane) { pos++; continue; }
    if (lj == lane) { const double xj = F[mcs::F_X * np + j]; pos += (xj < xk || (xj == xk && j < k)) ? 1 : 0; }
  }
  lane_vec[pos] = (uint8_t)k;
}

// utility.py:23-49 idm_acc_f_b with at most the leaders given (no follower): the outer loop's Python IDM
__device__ inline double py_idm(double acc_max, double acc_min, double min_dis, double acc_com, double thw, double vd, double v_limit,
                                double ego_v, int n_fronts, const double* f_v, const double* f_dis) {
  const double lim = 1.1 * v_limit;
  const double vdd = lim < vd ? lim : vd;                       // min(vd, 1.1 * v_limit)
  const double acc = acc_max * (1.0 - mcs_pow4(ego_v / vdd));
  double interaction = 0.0;
  if (n_fronts > 0) {
    double best = 0.0;
    for (int q = 0; q < n_fronts; ++q) {
      const double wish = 0.7 * (min_dis + ego_v * thw) + ego_v * (ego_v - f_v[q]) / (2.0 * sqrt(-acc_max * acc_com));
      const double gap = f_dis[q] > 1.0 ? f_dis[q] : 1.0;       // max(1, dis)
      const double r = wish / gap;
      const double pull = -acc_max * (r * r);
      if (q == 0 || pull < best) best = pull;                   // min(pull): the first of equal values stays
    }
    interaction = 0.0 + best;
  }
  const double s = acc + interaction;
  const double lo = acc_max < s ? acc_max : s;                  // min(acc_max, s)
  return lo > acc_min ? lo : acc_min;                           // max(lo, acc_min)
}

struct WalkShared {
  mce::EnvMapHead H;
  mce::EnvMapTail T;
  mce::EnvMapBorder B;
  mce::EnvLaneInfo info;
  WalkVehicle me;                  // the vehicle whose turn it is (copied in, worked on, copied back)
  int visit[64];
  mce::BuildScratch scratch;
  mce::BuildOut built;
  mcs::SceneDev sc;
  double F[mcs::F_COUNT * 64];
  int32_t I[mcs::I_COUNT * 64];
  uint8_t lane_vec[64];
  mce_build_request rq;
  mcs::DecideParams dp;
  mcs::DecideOut dout;
  double plan_ax, plan_vy;
  int go, kind;
  // the vehicle's yielding candidates between the two halves of its plan (the probabilities are computed one per warp)
  int nm, front;
  double front_v, front_d;
  int m_slot[kYieldCap];
  double m_dis[kYieldCap], m_p[kYieldCap];
};

// environment.py:113-115 dis_to_lane_border of a yielding candidate: the border towards the lane it is heading for
__device__ inline double yield_border(const mce::EnvMapBorder& B, int lane, int flags, const double* hull) {
  double hx[4], hy[4];
  for (int k = 0; k < 4; ++k) { hx[k] = hull[2 * k]; hy[k] = hull[2 * k + 1]; }
  return mce::border_distance_fn(B, lane, (flags & WF_EXIT_MODEL) ? MCE_SIDE_RIGHT : MCE_SIDE_LEFT, hx, hy);
}

#ifdef MCW_TIMING
#define MCW_T(k) do { if (t == 0) { const long long now_ = clock64(); tm_[k] += now_ - last_; last_ = now_; } } while (0)
#else
#define MCW_T(k) do { } while (0)
#endif
// dynamic shared memory after WalkShared: decide region (rollout layout, 64 slots, 0 steps) | tables of the scene and compact
// per-vehicle arrays: arc x y v vy l (doubles) | lane order vid vflags (ints) | lane starts
static __global__ void __launch_bounds__(128) walk_kernel(const mce::EnvMapHead* __restrict__ head, const mce::EnvMapTail* __restrict__ tail,
                                                          const mce::EnvMapBorder* __restrict__ border, const mce::EnvLaneInfo* __restrict__ g_info,
                                                          int n, const int32_t* __restrict__ g_lane, const double* __restrict__ g_arc,
                                                          const int32_t* __restrict__ g_order, const int32_t* __restrict__ g_start,
                                                          WalkVehicle* __restrict__ veh, const int32_t* __restrict__ ids, double* __restrict__ attrs,
                                                          const double* __restrict__ noise, const double* __restrict__ rnd_table,
                                                          const WalkParams wp, WalkResult* __restrict__ res, int decide_bytes) {
  extern __shared__ __align__(16) unsigned char walk_dyn[];
  WalkShared& W = *reinterpret_cast<WalkShared*>(walk_dyn);
  unsigned char* decide_raw = walk_dyn + ((sizeof(WalkShared) + 15) & ~(size_t)15);
  double* arc = reinterpret_cast<double*>(decide_raw + decide_bytes);
  double* sx = arc + n; double* sy = sx + n; double* sv = sy + n; double* svy = sv + n; double* sl = svy + n;
  int32_t* lane = reinterpret_cast<int32_t*>(sl + n); int32_t* order = lane + n; int32_t* vid = order + n; int32_t* vflags = vid + n;
  int32_t* start = vflags + n;
  // per vehicle, at its CURRENT position (computed for all before the walk, stale once the vehicle has moved — recomputed by
  // whoever needs it next): neighbour row, arc length on the own lane, signed distance to / remaining length of the own
  // centre line, distance of the hull to the border a yielding vehicle watches.  stale bits: 1 row, 2 centre + border, 4 arc
  double* nb_dis = reinterpret_cast<double*>(start + 16);
  double* arc_now = nb_dis + (size_t)n * MCE_NEIGHBORS; double* cd_sd = arc_now + n; double* cd_end = cd_sd + n; double* brd = cd_end + n;
  int32_t* nb_idx = reinterpret_cast<int32_t*>(brd + n); int32_t* stale = nb_idx + (size_t)n * MCE_NEIGHBORS;
  const int t = threadIdx.x;
#ifdef MCW_TIMING
  long long tm_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
  // the map (head: lane polygons / centre lines for the searches; tail: Frenet lines, step; border: centre / border distances)
  // and the lane info, staged once: everything the walk's single-thread chains read sits at shared-memory latency
  mce::stage(&W.H, head);
  mce::stage(&W.T, tail);
  mce::stage(&W.B, border);
  mce::stage(&W.info, g_info);
  for (int j = t; j < n; j += blockDim.x) {
    arc[j] = g_arc[(size_t)wp.scene * n + j]; lane[j] = g_lane[(size_t)wp.scene * n + j]; order[j] = g_order[(size_t)wp.scene * n + j];
    const WalkVehicle& o = veh[j];
    sx[j] = o.x; sy[j] = o.y; sv[j] = o.v; svy[j] = o.vy; sl[j] = o.l; vid[j] = o.id; vflags[j] = o.flags;
  }
  if (t <= MCE_MAX_LANES) start[t] = g_start[(size_t)wp.scene * (MCE_MAX_LANES + 1) + t];
  __shared__ WalkResult R;
  if (t == 0) { R.stopped_at = -1; R.status = 0; R.noise_used = 0; R.mobil_calls = 0; R.random_used = 0; R.random_seed = 0; R.moved = 0; R.pad = 0; }
  __syncthreads();
  const mce::EnvMapTail& T = W.T;
  const mce::EnvMapBorder& B = W.B;
  const mce::EnvLaneInfo* info = &W.info;
  WalkVehicle& me = W.me;
  mcs::Smem s;
  mcs::carve_smem(s, decide_raw, 64, 0);
  int* laneStart = s.cnt;
  {
    // one warp per quantity, one lane per vehicle
    const int job = t >> 5;
    for (int j = t & 31; j < n; j += 32) {
      const int cj = lane[j];
      if (job == 0) {
        int idx[MCE_NEIGHBORS];
        double dis[MCE_NEIGHBORS];
        mce::neighbor_row(W.H, n, lane, arc, order, start, 0, j, sx[j], sy[j], idx, dis);
        for (int k = 0; k < MCE_NEIGHBORS; ++k) { nb_idx[j * MCE_NEIGHBORS + k] = idx[k]; nb_dis[j * MCE_NEIGHBORS + k] = dis[k]; }
        arc_now[j] = arc[j];
        stale[j] = j < wp.first ? 4 : 0;      // a walk resumed behind a host-handled vehicle: the ones before have moved since the match
      } else if (job == 1) {
        double a = 0.0, b = 0.0;
        if (cj >= 0) mce::center_distance_fn(B, cj, sx[j], sy[j], a, b);
        cd_sd[j] = a; cd_end[j] = b;
      } else if (job == 2) {
        double d = 0.0;
        if (cj >= 0 && (W.info.type[cj] == MCS_LANE_MERGING || (vflags[j] & WF_EXIT_MODEL))) d = yield_border(B, cj, vflags[j], veh[j].hull);
        brd[j] = d;
      }
    }
  }
  __syncthreads();
  const mce::BuildCache cache{nb_idx, nb_dis, arc_now, stale};
  MCW_T(0);

  for (int i = wp.first; i < wp.last; ++i) {
    const int c = lane[i];
    if (c < 0) continue;                                        // left the map at the last match (block-uniform)
    {
      const double* src = reinterpret_cast<const double*>(veh + i);
      double* dst = reinterpret_cast<double*>(&W.me);
      for (int q = t; q < (int)(sizeof(WalkVehicle) / 8); q += blockDim.x) dst[q] = src[q];
    }
    __syncthreads();
    MCW_T(1);
    // ================= plan: idm_behavior — thread 0, the candidates' yielding probabilities one per warp =================
    const int* idx = nb_idx + i * MCE_NEIGHBORS;               // the vehicle has not moved yet: its row and centre distance are current
    const double* dis = nb_dis + i * MCE_NEIGHBORS;
    if (t == 0) {
      W.go = 1; W.kind = -1; W.nm = 0;
      const int beh = me.behavior;
      const bool needs_idm = beh == WB_IDM || beh == WB_MOBIL || beh == WB_MOBIL_SAFE;
      if (beh == WB_HOST || (needs_idm && me.seed == 0)) { R.stopped_at = i; W.go = 0; }
      else if (needs_idm) {
        const double x = sx[i], y = sy[i];
        const int f = idx[MCE_FRONT];
        W.front = f; W.front_v = 0.0; W.front_d = 0.0;
        if (f >= 0) { W.front_v = sv[f]; W.front_d = dis[MCE_FRONT]; }
        // candidates to yield to: ramp vehicles on the right, exiting vehicles on the left (behaviors.py:24-32)
        int nm = 0;
        const int right = W.H.right[c], left = W.H.left[c];
        if (right >= 0 && info->type[right] == MCS_LANE_MERGING) {
          const int ks[4] = {MCE_RIGHT_BB, MCE_RIGHT_B, MCE_RIGHT_F, MCE_RIGHT_FF};
          for (int q = 0; q < 4; ++q) if (idx[ks[q]] >= 0 && nm < kYieldCap) { W.m_slot[nm] = idx[ks[q]]; W.m_dis[nm] = dis[ks[q]]; nm++; }
        }
        if (left >= 0 && info->type[left] == MCS_LANE_MAIN) {
          for (int q = start[left]; q < start[left + 1]; ++q) {
            const int j = order[q];
            if (!(vflags[j] & WF_EXIT_MODEL)) continue;
            if (nm >= kYieldCap) { R.status = 300; break; }
            // agent_to_idm_match (environment.py:158-164): arc lengths of both on the other vehicle's matched lane, now
            double ae, ao, dd;
            mce::project_distance(W.H.center[lane[j]], W.H.n_center[lane[j]], x, y, ae, dd);
            if (!(stale[j] & 4)) ao = arc_now[j];
            else {
              mce::project_distance(W.H.center[lane[j]], W.H.n_center[lane[j]], sx[j], sy[j], ao, dd);
              arc_now[j] = ao; stale[j] &= ~4;
            }
            W.m_slot[nm] = j; W.m_dis[nm] = ao - ae; nm++;
          }
        }
        W.nm = nm;
      }
    }
    __syncthreads();
    if (W.nm > 0) {
      // yielding_probability (type.py:93-137), candidate q on warp q % 4
      for (int q = t >> 5; q < W.nm; q += 4) {
        if ((t & 31) != 0) continue;
        const int j = W.m_slot[q];
        const double v = me.v, el = me.l;
        const double o_l = sl[j], o_v = sv[j], o_vy = svy[j];
        double p = 0.0;
        const double half = 0.5 * (el + o_l);
        if (!(W.m_dis[q] + half < 0.0)) {
          const double gap = W.m_dis[q] - half;
          const double v_floor = v > 0.0001 ? v : 0.0001;
          if (stale[j] & 2) {
            brd[j] = yield_border(B, lane[j], vflags[j], veh[j].hull);
            mce::center_distance_fn(B, lane[j], sx[j], sy[j], cd_sd[j], cd_end[j]);
            atomicAnd(&stale[j], ~2);
          }
          const double border_d = brd[j], to_end = cd_end[j];
          const int f = W.front;
          const double fd = W.front_d, fv = W.front_v;
          const double thw = gap / (v > 1e-10 ? v : 1e-10);
          const double ttc = (v - o_v) / (v > 0.0 ? v : 1e-3);
          const double feat[12] = {f >= 0 ? fd : 200.0, gap, v, border_d, o_vy, v, f >= 0 ? fv : 20.0, thw, ttc, to_end, to_end / v_floor,
                                   to_end * (1.0 - v / v_floor)};
          double r = me.yp[0];
          for (int k = 0; k < 12; ++k) r = r + me.yp[k + 1] * feat[k];
          r = r < 100.0 ? r : 100.0;                            // min(100, r)
          r = r > -100.0 ? r : -100.0;                          // max(.., -100)
          p = 1.0 / (1.0 + mcs_exp(-r));
        }
        W.m_p[q] = p;
      }
      __syncthreads();
    }
    if (t == 0 && W.go) {
      const int beh = me.behavior;
      const bool needs_idm = beh == WB_IDM || beh == WB_MOBIL || beh == WB_MOBIL_SAFE;
      if (needs_idm) {
        const double v = me.v, el = me.l;
        const int f = W.front, nm = W.nm;
        const double fv = W.front_v, fd = W.front_d;
        const double v_limit = info->v_limit[c];
        double ax = py_idm(me.idm[0], me.idm[1], me.idm[2], me.idm[3], me.idm[4], me.idm[5], v_limit, v, f >= 0 ? 1 : 0, &fv, &fd);
        const double sd = cd_sd[i];
        const double asd = fabs(sd);
        double vy_abs = asd * 1.0 > 0.1 ? asd * 1.0 : 0.1;       // max(abs * 1.0, 0.1)
        vy_abs = vy_abs < 1.3 ? vy_abs : 1.3;                    // min(.., 1.3)
        const double vy = -sd / (asd + 1e-10) * vy_abs;
        // intentions (behaviors.py:33-44, type.py:70-92)
        bool yielding[kYieldCap];
        for (int q = 0; q < nm; ++q) {
          const int o_id = vid[W.m_slot[q]];
          const double p = W.m_p[q];
          int e = -1;
          for (int k = 0; k < me.yi_n; ++k) if (me.yi_id[k] == o_id) { e = k; break; }
          if (e < 0) {
            // YieldingIntention(...): random.seed(seed); yielding = random.random() <= p
            if (me.yi_n >= kYieldCap) { R.status = 300; yielding[q] = false; continue; }
            e = me.yi_n++;
            me.yi_id[e] = o_id; me.yi_p0[e] = p;
            me.yi_yielding[e] = rnd_table[me.seed] <= p ? 1 : 0;
            R.random_used = 1; R.random_seed = me.seed;
          } else {
            const double p0 = me.yi_p0[e], k_thr = me.threshold;
            if (me.yi_yielding[e]) me.yi_yielding[e] = !(p <= k_thr * p0) ? 1 : 0;
            else me.yi_yielding[e] = ((1.0 - p) <= k_thr * (1.0 - p0)) ? 1 : 0;
          }
          yielding[q] = me.yi_yielding[e] != 0;
        }
        // drop the intentions towards vehicles that are no candidates any more (behaviors.py:45-48)
        {
          int w_ = 0;
          for (int k = 0; k < me.yi_n; ++k) {
            bool wanted = false;
            for (int q = 0; q < nm; ++q) if (vid[W.m_slot[q]] == me.yi_id[k]) wanted = true;
            if (wanted) { me.yi_id[w_] = me.yi_id[k]; me.yi_p0[w_] = me.yi_p0[k]; me.yi_yielding[w_] = me.yi_yielding[k]; w_++; }
          }
          me.yi_n = w_;
        }
        double yv[kYieldCap], yd[kYieldCap];
        int ny = 0;
        for (int q = 0; q < nm; ++q) if (yielding[q]) { yv[ny] = sv[W.m_slot[q]]; yd[ny] = W.m_dis[q]; ny++; }
        if (ny > 0) {
          // idm_p_for_merging_exit_agents (utility.py:133-141): random.seed(seed); u = random.random(); cars soften
          const double u = rnd_table[me.seed];
          R.random_used = 1; R.random_seed = me.seed;
          double acc_max = me.idm[0], acc_min = me.idm[1];
          if (!(el > 7.0)) { acc_max = acc_max - (0.6 + 0.4 * u); acc_min = acc_min + (0.8 + 0.4 * u); }
          const double ay_ = py_idm(acc_max, acc_min, me.idm[2], me.idm[3], me.idm[4], me.idm[5], v_limit, v, ny, yv, yd);
          ax = ay_ < ax ? ay_ : ax;                               // min(ax, ay)
        }
        if (f >= 0) {
          // rss_dis(front.v, v, r_t_ego, a_min, a_max) > front.dis - 0.5 (front.l + l)  (behaviors.py:53-55)
          const double rear = v * me.rss[0] + v * v / (-2.0 * me.rss[2]);
          double need = rear - fv * fv / (-2.0 * me.rss[3]);
          need = need > 0.0 ? need : 0.0;
          if (need > fd - 0.5 * (sl[f] + el)) ax = me.rss[2];
        }
        W.plan_ax = ax; W.plan_vy = vy;
      }
      // ---- the builder request of the behaviours that go through the inner simulator
      if (W.go && beh != WB_IDM) {
        mce_build_request& rq = W.rq;
        const bool merging_model = (me.flags & WF_MERGING_MODEL) != 0;
        rq.kind = beh == WB_MERGE_CLOSEST ? 0 : (beh == WB_EXIT_CLOSEST ? 1 : 2);
        rq.scene = 0; rq.ego_slot = i;
        rq.ego_behavior = rq.kind == 1 ? MCS_BEH_EXIT : (merging_model ? MCS_BEH_MERGING : MCS_BEH_IDM);
        rq.others_merging = merging_model ? 1 : 0;
        const bool poke = rq.kind == 2;
        rq.ego_commit = poke ? me.commit : MCS_COMMIT_NOT_DECIDED; rq.ego_keep_step = poke ? me.keep_step : 0;
        rq.ego_target_lane = poke ? me.target_lane : -1;
        rq.ego_vd = me.idm[5];
        rq.ego_idm[0] = me.idm[0]; rq.ego_idm[1] = me.idm[2]; rq.ego_idm[2] = me.idm[1]; rq.ego_idm[3] = me.idm[3]; rq.ego_idm[4] = -8.0;
        rq.ego_idm[5] = me.idm[4];
        if (rq.kind == 2) {
          rq.ego_rss[0] = me.rss[1]; rq.ego_rss[1] = me.rss[0]; rq.ego_rss[2] = me.rss[2]; rq.ego_rss[3] = me.rss[3]; rq.ego_rss[4] = me.rss[4];
          rq.ego_rss[5] = me.rss[5];
        } else {
          rq.ego_rss[0] = 0.7; rq.ego_rss[1] = 0.4; rq.ego_rss[2] = -8.0; rq.ego_rss[3] = -10.0; rq.ego_rss[4] = 1.8; rq.ego_rss[5] = 1.2;
        }
        rq.ego_yp[0] = me.yp[0]; rq.ego_yp[1] = me.yp[1]; rq.ego_yp[2] = me.yp[2]; rq.ego_yp[3] = me.yp[3];
        rq.ego_yp[4] = me.mobil[0]; rq.ego_yp[5] = me.mobil[1]; rq.ego_yp[6] = me.mobil[2];
        W.kind = rq.kind;
      }
    }
    __syncthreads();
    MCW_T(2);
    if (!W.go || R.status != 0) break;
    if (W.kind >= 0) {
      // ================= mcs_wrapper's builder -> inner constructors -> the single-step decision =================
      // the builder reads the other vehicles' current v vy a l w vd from `attrs` (kept current below) and their positions from sx / sy
      mce::build_scene_block(W.H, &W.T, &W.B, info, n, lane, arc, order, start, sx, sy, W.rq, ids, attrs, noise + R.noise_used,
                             wp.n_noise - R.noise_used, W.scratch, W.built, &cache);
      __syncthreads();
      MCW_T(3);
      if (t == 0) {
        if (W.built.status != 0) R.status = 100 + W.built.status;
        else {
          R.noise_used += W.built.n_agents - 1;
          prep_orders(W.built, W.F, W.I, W.lane_vec, W.sc, W.visit);
        }
      }
      __syncthreads();
      if (R.status != 0) break;
      if (t < 64) prep_vehicle(W.built, W.F, W.I, W.sc, W.visit, t);
      __syncthreads();
      if (t < 64) prep_lane_vectors(W.built, W.F, W.I, W.lane_vec, W.sc, t);
      if (t == 64) {
        int ego_slot = -1;
        for (int k = 0; k < W.built.n_agents; ++k) if (W.built.agents[W.visit[k]].id == me.id && ego_slot < 0) ego_slot = k;
        mcs::DecideParams& dp = W.dp;
        dp.slot = ego_slot; dp.gap_slot = -1; dp.pad = 0; dp.ax = 0.0; dp.vy = 0.0; dp.key = 0; dp.require_rss = 0; dp.action = 0;
        if (W.kind == 2) {
          dp.kind = 0; dp.ax = W.plan_ax; dp.vy = W.plan_vy; dp.require_rss = me.behavior == WB_MOBIL_SAFE ? 1 : 0;
          dp.key = wp.seed_base + (uint64_t)R.mobil_calls;
          R.mobil_calls++;
        } else {
          dp.kind = 3; dp.action = W.kind == 0 ? MCS_ACT_LEFT : MCS_ACT_RIGHT;
        }
      }
      __syncthreads();
      MCW_T(4);
      if (t < 64) mcs::decide_stage(W.sc, s, laneStart, t);
      __syncthreads();
      MCW_T(5);
      if (t == 0) {
        mcs::decide_eval(W.sc, s, laneStart, W.dp, W.dout);
        if (W.dout.status != MCS_OK) R.status = 200 - W.dout.status;
        W.plan_ax = W.dout.ax; W.plan_vy = W.dout.vy;
        if (W.kind == 2) { me.commit = W.dout.commit; me.keep_step = W.dout.keepStep; me.target_lane = W.dout.targetLane; }
      }
      __syncthreads();
      MCW_T(6);
      if (R.status != 0) break;
    }
    // ================= move: Agent.step (agent.py:193-208) =================
    if (t == 0) {
      double xn, yn, vn;
      int k;
      mce::step_along_path_fn(T, c, sx[i], sy[i], me.v, W.plan_ax, W.plan_vy, wp.dt, xn, yn, vn, k);
      const double yaw = T.cyaw[c][k], cs = T.ccos[c][k], sn = T.csin[c][k];
      me.x = xn; me.y = yn; me.v = vn; me.yaw = yaw;
      me.a = vn == 0.0 ? 0.0 : W.plan_ax;
      me.vy = W.plan_vy;
      me.ax_last = W.plan_ax;
      // vehicle_pose_to_rect (utility.py:12-20)
      const double ax_ = me.l / 2.0 * cs, ay_ = me.l / 2.0 * sn, bx = me.w / 2.0 * sn, by = me.w / 2.0 * cs;
      const double hx = xn + ax_, hy = yn + ay_, tx = xn - ax_, ty = yn - ay_;
      me.hull[0] = hx - bx; me.hull[1] = hy + by; me.hull[2] = hx + bx; me.hull[3] = hy - by;
      me.hull[4] = tx + bx; me.hull[5] = ty - by; me.hull[6] = tx - bx; me.hull[7] = ty + by;
      sx[i] = xn; sy[i] = yn; sv[i] = vn; svy[i] = me.vy; stale[i] = 7;
      attrs[i] = xn; attrs[(size_t)n + i] = yn; attrs[2 * (size_t)n + i] = vn; attrs[3 * (size_t)n + i] = me.vy; attrs[4 * (size_t)n + i] = me.a;
      R.moved++;
    }
    __syncthreads();
    MCW_T(7);
    {
      const double* src = reinterpret_cast<const double*>(&W.me);
      double* dst = reinterpret_cast<double*>(veh + i);
      for (int q = t; q < (int)(sizeof(WalkVehicle) / 8); q += blockDim.x) dst[q] = src[q];
    }
    __syncthreads();
    MCW_T(8);
  }
#ifdef MCW_TIMING
  if (t == 0) printf("walk %d..%d stage %lld in %lld plan %lld build %lld prep %lld dstage %lld deval %lld move %lld out %lld\n", wp.first, wp.last,
                     tm_[0], tm_[1], tm_[2], tm_[3], tm_[4], tm_[5], tm_[6], tm_[7], tm_[8]);
#endif
  if (t == 0) *res = R;
}

}  // namespace mcw

#include "scene.hpp"

#include <algorithm>
#include <cmath>
#include <unordered_map>

#include "../../include/mcsim_math.h"

namespace mcs {

namespace {
// lane-change prior, reference: Environment::matchAgentsOnCorridor 0x10bb97-0x10bd8a (first match of an agent)
void prior_from_offset(double y, double vy, double centerY, double out[3]) {
  const double d = ((y - centerY) * 0.333 + 0.6805 * vy) - 0.01;
  auto gauss = [](double t, double w) { return mcs_exp(((-t) * t) / w); };
  const double a = (0.8 * gauss(d - 1.04, 0.055) + 0.8 * gauss(d - 0.35, 0.055)) + gauss(d - 0.7, 0.07) * 0.65;
  const double c = (0.75 * gauss(1.04 + d, 0.055) + 0.75 * gauss(0.35 + d, 0.055)) + gauss(0.7 + d, 0.07) * 0.6;
  const double b = gauss(d, 0.055) * 2.5;
  const double s = (a + b) + c;
  out[0] = a / s; out[1] = b / s; out[2] = c / s;
}
}  // namespace

int build_scene_host(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, SceneHost& s) {
  if (!cs || !as || nc <= 0 || na <= 0) { s.error = "empty scene"; return MCS_ERR_ARG; }
  if (nc > MCS_MAX_LANES || na > MCS_MAX_AGENTS) { s.error = "scene exceeds the compiled limits: at most " + std::to_string(MCS_MAX_AGENTS) + " vehicles per rollout and " +
                                                               std::to_string(MCS_MAX_LANES) + " corridors (got " + std::to_string(na) + " / " + std::to_string(nc) + ")";
                                                     return MCS_ERR_CAPACITY; }
  // ---- lanes: hash order, neighbours by ascending centre y (MapMultiLane ctor)
  std::unordered_map<int, int> lane_order;
  for (int k = 0; k < nc; ++k) lane_order.emplace(cs[k].id, k);
  std::vector<int> order;
  for (auto& kv : lane_order) order.push_back(kv.second);
  s.n_lanes = (int)order.size();
  std::vector<int> by_y(order);
  std::sort(by_y.begin(), by_y.end(), [&](int a, int b) { return cs[a].center[1] < cs[b].center[1]; });
  std::unordered_map<int, int> slot_of_lane_input;
  for (int k = 0; k < s.n_lanes; ++k) slot_of_lane_input[order[k]] = k;
  s.lanes.resize(s.n_lanes);
  int leftmost_slot = -1;
  for (int k = 0; k < s.n_lanes; ++k) {
    const mcs_corridor& c = cs[order[k]];
    LaneHost& L = s.lanes[k];
    L.id = c.id; L.type = c.type;
    L.leftY = c.left[1]; L.rightY = c.right[1]; L.centerY = c.center[1]; L.vLimit = c.v_limit; L.endX = c.center[2]; L.startX = c.center[0];
    L.leftIdx = L.rightIdx = -1; L.leftmost = 0;
    for (int j : by_y) if (cs[j].center[1] > c.center[1]) { L.leftIdx = slot_of_lane_input[j]; break; }
    for (auto it = by_y.rbegin(); it != by_y.rend(); ++it)
      if (c.center[1] > cs[*it].center[1]) { L.rightIdx = slot_of_lane_input[*it]; break; }
    if (c.type == MCS_LANE_EXIT) s.has_exit_lanes = true;
    if (c.type != MCS_LANE_MAIN) s.needs_merge_path = true;
    if (L.leftIdx < 0) leftmost_slot = k;   // last lane without a left neighbour in iteration order
  }
  if (leftmost_slot >= 0) s.lanes[leftmost_slot].leftmost = 1;
  // ---- agents: visit order = iteration order of unordered_map<int,AgentPtr> after emplace in input order
  std::unordered_map<int, int> agent_order;
  for (int k = 0; k < na; ++k)
    if (!agent_order.emplace(as[k].id, k).second) { s.error = "duplicate agent id"; return MCS_ERR_ARG; }
  std::vector<int> visit;
  for (auto& kv : agent_order) visit.push_back(kv.second);
  s.n = na;
  s.n_pad = na <= 32 ? 32 : na <= 64 ? 64 : na <= 128 ? 128 : 256;   // one kernel instantiation per block size
  s.f.assign((size_t)F_COUNT * s.n_pad, 0.0);
  s.i.assign((size_t)I_COUNT * s.n_pad, -1);
  s.slot_of_input.assign(na, -1);
  for (int k = 0; k < na; ++k) {
    const mcs_agent& a = as[visit[k]];
    s.slot_of_input[visit[k]] = k;
    s.F(F_X, k) = a.x; s.F(F_Y, k) = a.y; s.F(F_V, k) = a.v; s.F(F_VY, k) = a.vy; s.F(F_A, k) = a.a;
    s.F(F_L, k) = a.l; s.F(F_W, k) = a.w; s.F(F_VD, k) = a.vd;
    if (a.behavior == MCS_BEH_MERGING || a.behavior == MCS_BEH_EXIT) s.needs_merge_path = true;
    s.F(F_IDM_ACCMAX, k) = a.idm[0]; s.F(F_IDM_MINDIS, k) = a.idm[1]; s.F(F_IDM_ACCMIN, k) = a.idm[2];
    s.F(F_IDM_ACCCOM, k) = a.idm[3]; s.F(F_IDM_THW, k) = a.idm[5];
    // Agent ctor 0xf8941-0xf89db: softened copy for yielding, cars (l < 7) only
    s.F(F_IDMY_ACCMAX, k) = (7.0 > a.l) ? a.idm[0] - 0.8 : a.idm[0];
    s.F(F_IDMY_ACCMIN, k) = (7.0 > a.l) ? 1.0 + a.idm[2] : a.idm[2];
    s.F(F_RSS_RTOTHER, k) = a.rss[0]; s.F(F_RSS_RTEGO, k) = a.rss[1]; s.F(F_RSS_AMIN, k) = a.rss[2]; s.F(F_RSS_AMAX, k) = a.rss[3]; s.F(F_RSS_DSOFT, k) = a.rss[5];
    s.F(F_YP_BIAS, k) = a.yp[0]; s.F(F_YP_A, k) = a.yp[1]; s.F(F_YP_B, k) = a.yp[2]; s.F(F_YP_C, k) = a.yp[3];
    s.F(F_YP_PF, k) = a.yp[4]; s.F(F_YP_DA, k) = a.yp[5]; s.F(F_YP_BA, k) = a.yp[6];
    s.I(I_ID, k) = a.id; s.I(I_INPUT, k) = visit[k]; s.I(I_BEH, k) = a.behavior; s.I(I_BEH0, k) = a.behavior;
    s.I(I_COMMIT, k) = a.commit; s.I(I_KEEPSTEP, k) = a.commit_keep_lane_step; s.I(I_TARGETLANE, k) = a.target_lane_id;
    // first lane match: first lane in iteration order whose lateral interval contains y
    int lane = -1;
    for (int j = 0; j < s.n_lanes; ++j)
      if (s.lanes[j].leftY >= a.y && a.y >= s.lanes[j].rightY) { lane = j; break; }
    s.I(I_LANE, k) = lane;
    if (lane >= 0) {
      if (s.lanes[lane].type == MCS_LANE_MAIN && a.behavior == MCS_BEH_MERGING) s.I(I_BEH, k) = MCS_BEH_IDM_LC;
      if (s.lanes[lane].type == MCS_LANE_EXIT && a.behavior == MCS_BEH_EXIT) s.I(I_BEH, k) = MCS_BEH_IDM;
      double p[3];
      prior_from_offset(a.y, a.vy, s.lanes[lane].centerY, p);
      s.F(F_PRIOR_L, k) = p[0]; s.F(F_PRIOR_K, k) = p[1]; s.F(F_PRIOR_R, k) = p[2];
    }
  }
  // ---- initial per-lane vectors: push in visit order, sort ascending x (ties keep visit order)
  s.lane_start.assign(s.n_lanes + 1, 0);
  s.lane_vec.clear();
  for (int j = 0; j < s.n_lanes; ++j) {
    std::vector<int> v;
    for (int k = 0; k < na; ++k) if (s.I(I_LANE, k) == j) v.push_back(k);
    std::stable_sort(v.begin(), v.end(), [&](int p, int q) { return s.F(F_X, p) < s.F(F_X, q); });
    s.lane_start[j] = (int)s.lane_vec.size();
    for (int k : v) s.lane_vec.push_back((uint8_t)k);
  }
  s.lane_start[s.n_lanes] = (int)s.lane_vec.size();
  s.lane_vec.resize(s.n_pad, 0);
  return MCS_OK;
}

CandHost resolve_candidate(const SceneHost& s, const mcs_candidate& c) {
  CandHost r{-1, c.action, c.gap_id, -1, 0, -1, 0};
  for (int k = 0; k < s.n; ++k) {
    const int id = s.i[(size_t)I_ID * s.n_pad + k];
    if (id == c.ego_id && r.ego_slot < 0) r.ego_slot = k;
    if (c.gap_id >= 0 && id == c.gap_id && r.gap_slot < 0) r.gap_slot = k;
  }
  if (r.ego_slot < 0) return r;                       // "No ego agent with id"
  // Simulation::run 0x10ed15: a non-ego vehicle that is neither idm, idm_lc nor merging gets no action
  for (int k = 0; k < s.n; ++k) {
    const int b = s.i[(size_t)I_BEH * s.n_pad + k];
    if (k != r.ego_slot && b != MCS_BEH_IDM && b != MCS_BEH_IDM_LC && b != MCS_BEH_MERGING) r.bad_behavior = 1;
  }
  const int lane = s.i[(size_t)I_LANE * s.n_pad + r.ego_slot];
  if (lane < 0) return r;                             // "is not matched to any corridor"
  const LaneHost& L = s.lanes[lane];
  switch (c.action) {
    case MCS_ACT_LEFT: if (L.leftIdx >= 0) { r.target_lane_id = s.lanes[L.leftIdx].id; r.ok = 1; } break;
    case MCS_ACT_RIGHT: if (L.rightIdx >= 0) { r.target_lane_id = s.lanes[L.rightIdx].id; r.ok = 1; } break;
    case MCS_ACT_KEEP_LANE: case MCS_ACT_ACC: case MCS_ACT_DCC: r.target_lane_id = L.id; r.ok = 1; break;
    default: break;
  }
  return r;
}

}  // namespace mcs

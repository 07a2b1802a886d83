// Host-side scene preparation for libmcsim_b200.so: everything the reference does ONCE per
// simulationMultiThread call before the rollouts start, restated for a structure-of-arrays device layout.
//   MapMultiLane(vector<Corridor>)            mc_sim_highway.so 0x109f00  → lane order + neighbour lanes
//   Environment(map, vector<Agent>)            0x10c6c0                   → agent visit order
//   Environment::matchAgentsOnCorridor (first) 0x10b810                   → initial lane, behaviour switch, prior
//   Agent::Agent                               0xf8380                    → softened IDM / relaxed RSS parameters
// The two std::unordered_map<int,...> iteration orders the reference walks (agents: plan order and
// random-draw order; corridors: first-match order) are reproduced by replaying the same emplace
// sequence into a libstdc++ std::unordered_map.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/mcsim.h"

namespace mcs {

// per-agent double fields, one array of n_pad doubles each, slot order = reference visit order
enum AgentField {
  F_X = 0, F_Y, F_V, F_VY, F_A, F_L, F_W, F_VD,
  F_IDM_ACCMAX, F_IDM_MINDIS, F_IDM_ACCMIN, F_IDM_ACCCOM, F_IDM_THW,
  F_IDMY_ACCMAX, F_IDMY_ACCMIN,
  F_RSS_RTOTHER, F_RSS_RTEGO, F_RSS_AMIN, F_RSS_AMAX, F_RSS_DSOFT,
  F_YP_BIAS, F_YP_A, F_YP_B, F_YP_C, F_YP_PF, F_YP_DA, F_YP_BA,
  F_PRIOR_L, F_PRIOR_K, F_PRIOR_R,
  F_COUNT
};
// per-agent int fields
enum AgentIntField { I_ID = 0, I_INPUT, I_BEH, I_BEH0, I_LANE, I_COMMIT, I_KEEPSTEP, I_TARGETLANE, I_COUNT };

struct LaneHost {
  int id, type;
  double leftY, rightY, centerY, vLimit, endX, startX;  // center.p2.x, center.p1.x
  int leftIdx, rightIdx;                        // slot of the neighbour lane or -1
  int leftmost;                                 // 1 if this is MapMultiLane's "leftmost" lane id (0x18)
};

struct SceneHost {
  int n = 0, n_pad = 0, n_lanes = 0;
  bool has_exit_lanes = false;
  bool needs_merge_path = false;  // a merging/exit lane or a merging/exit vehicle: yielding + gap behaviours are reachable
  std::vector<double> f;          // F_COUNT * n_pad
  std::vector<int32_t> i;         // I_COUNT * n_pad
  std::vector<LaneHost> lanes;    // hash iteration order
  std::vector<uint8_t> lane_vec;  // initial per-lane vectors (slots), concatenated in lane order
  std::vector<int32_t> lane_start;  // n_lanes + 1
  std::vector<int32_t> slot_of_input;  // input index → slot
  std::string error;
  double& F(int field, int slot) { return f[(size_t)field * n_pad + slot]; }
  int32_t& I(int field, int slot) { return i[(size_t)field * n_pad + slot]; }
};

// returns MCS_OK or an error code; fills `out`
int build_scene_host(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents, SceneHost& out);

// Environment::setEgoAgentIdAndDrivingBehavior 0xf3f20 evaluated on the host for one candidate
struct CandHost { int ego_slot, action, gap_id, target_lane_id, ok, gap_slot /* -1 = last gap / unknown */, bad_behavior; };
CandHost resolve_candidate(const SceneHost& s, const mcs_candidate& c);

}  // namespace mcs

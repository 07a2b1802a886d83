// Single-decision kernel: the four Python-facing wrappers of the reference's binding, evaluated on the device-resident
// scene exactly as on a freshly constructed Environment (ctor + first matchAgentsOnCorridor, done on the host in scene.cpp).
//   laneChangeBehaviorMOBILWrapper        mc_sim_highway.so 0xb9960  → kind 0
//   customizedLaneChangeBehaviorWrapper   0xb9aa0..                  → kind 1
//   customizedMergingBehaviorWrapper      0xb9cc0                    → kind 2
//   closestGapMergingBehaviorWrapper                                 → kind 3
// One block stages the scene in shared memory with the rollout kernel's layout; the thread of the asked vehicle evaluates
// the same device functions the rollouts use.
#pragma once
#include "rollout_kernel.cuh"

namespace mcs {

struct DecideParams {
  int slot, kind, action, gap_slot, require_rss, pad;
  double ax, vy;       // MOBIL: the vehicle's IDM action (Action idmAction)
  uint64_t key;        // MOBIL: random stream; the decision draw is draw n_agents (after the n Agent-ctor draws)
};
struct DecideOut { double ax, vy; int commit, keepStep, targetLane, status; };

// Stage one vehicle slot of the scene into the rollout kernels' shared-memory layout (every thread k < np of the block)
__device__ __forceinline__ void decide_stage(const SceneDev& sc, const Smem& s, int* laneStart, int k) {
  const int np = sc.n_pad, n = sc.n;
  const double* F = sc.f;
  const int32_t* I = sc.i;
  const bool live = k < n;
  const int lane = live ? I[(size_t)I_LANE * np + k] : -1;
  s.x[k] = live ? F[(size_t)F_X * np + k] : 0.0;
  s.v[k] = live ? F[(size_t)F_V * np + k] : 0.0;
  s.l[k] = live ? F[(size_t)F_L * np + k] : 0.0;
  s.vd[k] = live ? F[(size_t)F_VD * np + k] : 1.0;
  s.idm[k] = live ? F[(size_t)F_IDM_ACCMAX * np + k] : 1.0;
  s.idm[np + k] = live ? F[(size_t)F_IDM_MINDIS * np + k] : 0.0;
  s.idm[2 * np + k] = live ? F[(size_t)F_IDM_ACCMIN * np + k] : 0.0;
  s.idm[3 * np + k] = live ? F[(size_t)F_IDM_ACCCOM * np + k] : -1.0;
  s.idm[4 * np + k] = live ? F[(size_t)F_IDM_THW * np + k] : 0.0;
  { const double sq = sqrt((-s.idm[k]) * s.idm[3 * np + k]); s.idm[5 * np + k] = sq + sq; }
  s.lane[k] = (uint8_t)(lane < 0 ? 0xff : lane);
  s.vec[k] = sc.lane_vec[k];
  if (k <= sc.n_lanes) laneStart[k] = sc.lane_start[k];
}

// One decision of vehicle dp.slot on the staged scene (one thread)
__device__ __forceinline__ void decide_eval(const SceneDev& sc, const Smem& s, const int* laneStart, const DecideParams& dp, DecideOut& out) {
  const int np = sc.n_pad, n = sc.n, k = dp.slot;
  const double* F = sc.f;
  const int32_t* I = sc.i;
  const int lane = I[(size_t)I_LANE * np + k];
  DecideOut o;
  o.ax = 0.0; o.vy = 0.0; o.status = MCS_OK;
  o.commit = I[(size_t)I_COMMIT * np + k]; o.keepStep = I[(size_t)I_KEEPSTEP * np + k]; o.targetLane = I[(size_t)I_TARGETLANE * np + k];
  if (lane < 0) {       // "is not matched!" — the wrappers return a zero action; MOBIL dereferences an empty history
    o.status = dp.kind == 0 ? MCS_ERR_SCENE : MCS_OK;
    out = o;
    return;
  }
  const LaneDev& L = sc.lanes[lane];
  const double x = s.x[k], v = s.v[k], el = s.l[k], evd = s.vd[k];
  const double y = F[(size_t)F_Y * np + k];
  const double rTOther = F[(size_t)F_RSS_RTOTHER * np + k], rTEgo = F[(size_t)F_RSS_RTEGO * np + k];
  const double aMin = F[(size_t)F_RSS_AMIN * np + k], aMax = F[(size_t)F_RSS_AMAX * np + k];
  const int lo = laneStart[lane], cnt = laneStart[lane + 1] - lo;
  const int fpos = lane_upper(s, lo, cnt, x);
  const int front = fpos < cnt ? (int)s.vec[lo + fpos] : -1;

  if (dp.kind == 0) {
    double ax = dp.ax, avy = dp.vy;
    MobilState ms{o.commit, o.keepStep, o.targetLane};
    if (ms.commit == MCS_COMMIT_KEEP_LANE && ms.keepStep <= 9) {
      ms.keepStep = ms.keepStep + 1;
    } else {
      const bool canL = L.leftIdx >= 0 && sc.lanes[L.leftIdx].type == MCS_LANE_MAIN;
      const bool canR = L.rightIdx >= 0 && sc.lanes[L.rightIdx].type == MCS_LANE_MAIN;
      if (canL || canR) {
        const int bpos = lane_reverse(s, lo, cnt, x);
        const int back = bpos > 0 ? (int)s.vec[lo + bpos - 1] : -1;
        const MobilGeo geo = mobil_geometry_all(s, sc, laneStart, np, k, front, back);
        const double u = mcs_unit31(mcs_rand31(dp.key, (uint64_t)n));
        mobil_decide(geo, sc, np, k, lane, v, el, aMin, F[(size_t)F_YP_PF * np + k], F[(size_t)F_YP_DA * np + k],
                     F[(size_t)F_YP_BA * np + k], dp.require_rss ? 2 : 3, false, F + (size_t)F_PRIOR_L * np, u, ms, ax, avy);
      }
    }
    o.ax = ax; o.vy = avy; o.commit = ms.commit; o.keepStep = ms.keepStep; o.targetLane = ms.targetLane;
  } else if (dp.kind == 1) {
    double ax = 0.0, avy = centering_vy(y, L.centerY);
    customized_lane_change(s, sc, laneStart, np, k, lane, front, dp.action, x, v, el, evd, rTOther, rTEgo, aMin, aMax, ax, avy);
    o.ax = ax; o.vy = avy;
  } else {
    MergeSelf me;
    me.k = k; me.lane = lane; me.front = front; me.x = x; me.y = y; me.v = v; me.l = el;
    me.w = F[(size_t)F_W * np + k];
    me.rss.rTOther = rTOther; me.rss.rTEgo = rTEgo; me.rss.aMin = aMin; me.rss.aMax = aMax;
    me.rss.dSoft = F[(size_t)F_RSS_DSOFT * np + k];
    me.idm = load_idm(s, np, k);
    double ax = 0.0, avy = 0.0;
    bool ok;
    if (dp.kind == 2) ok = customized_merging(s, sc, laneStart, me, evd, dp.action == MCS_ACT_LEFT, dp.gap_slot, ax, avy);
    else ok = merging_closest_gap(s, sc, laneStart, me, evd, false, dp.action == MCS_ACT_LEFT, ax, avy);
    if (!ok) o.status = MCS_ERR_SCENE;     // no corridor on that side: the reference reads a missing corridor
    o.ax = ax; o.vy = avy;
  }
  out = o;
}

static __global__ void __launch_bounds__(256, 1) decide_kernel(const SceneDev sc, const DecideParams dp, DecideOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = threadIdx.x;
  Smem s;
  carve_smem(s, smem_raw, sc.n_pad, 0);
  int* laneStart = s.cnt;
  decide_stage(sc, s, laneStart, k);
  __syncthreads();
  if (k != dp.slot) return;
  DecideOut o;
  decide_eval(sc, s, laneStart, dp, o);
  *out = o;
}

}  // namespace mcs

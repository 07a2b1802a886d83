// Bundle kernel for sm_100a: Simulation::run (mc_sim_highway.so 0x10e790) for SMALL scenes (<= 64 vehicles — every size
// the reference itself hands to its inner simulator, and BASELINE.json configs[1]) with the Monte-Carlo dimension across
// the lanes of a warp.
//
// rollout_kernel.cuh puts the vehicles of ONE rollout on the lanes of a warp.  In a small merging / exit scene every
// vehicle runs different code — the ego executes its gap, ramp vehicles pick the closest gap, their main-lane neighbours
// evaluate yielding intentions, the rest plain IDM / MOBIL — so a warp walks all of those paths one after the other (ncu
// r01q: 8 of 32 threads active per instruction, barrier stalls 7.3 per issue).  Rollouts of one decision, however, differ
// only in their random draws: vehicle k is in the same behaviour, usually the same branch, in all of them.  Here a block
// is a BUNDLE of RW rollouts in lockstep and a warp holds RW rollouts x 32/RW vehicle slots: lane = (rollout, slot), one
// thread per (rollout, vehicle), the slots sorted by behaviour class on the host.  The device functions are the ones
// rollout_kernel uses (same operation order, bit-identical results); what changes is the plumbing between them:
//   * per-rollout state that neighbours read (x, v, lane vectors) in a private shared-memory region per rollout, region
//     stride an odd number of 8-byte words so that the RW rollouts of a warp hit different banks; static per-vehicle
//     rows (l, vd, IDM, RSS) once per block;
//   * leader / follower always by the reference's binary searches (no "is every lane vector ordered" vote), all seven
//     MOBIL sub-tasks in the vehicle's own thread (no task queue: the same vehicle re-decides at the same step in most
//     rollouts), the draw positions from a byte array of per-vehicle draw counts summed four at a time;
//   * a lane-vector rebuild is a full re-rank by (x, slot) with per-byte SIMD compares over the 64 lane bytes;
//   * 3 block barriers per step (4 when some rollout of the bundle changed a lane membership) instead of 6-10.
#pragma once
#include "rollout_kernel.cuh"

namespace mcs {

constexpr int BNP = 64;                                   // slot capacity = row stride of the scene blob for this kernel
constexpr size_t BUNDLE_STATIC_BYTES = 11 * BNP * 8;      // l vd idm[6] rss[3], shared by the rollouts of a block

struct BundleParams {
  int rw_log2;        // log2(rollouts per warp) = log2(rollouts per block)
  int ns;             // slot threads per rollout: vehicles rounded up to a multiple of 32 >> rw_log2
  int region_bytes;   // shared memory of one rollout
  int pad;
};

// shared memory of one rollout (host and device agree through this one function)
__host__ __device__ inline size_t bundle_region_bytes(int steps, bool pool) {
  const size_t hist = (size_t)((steps + 2) & ~1);
  size_t b = 2 * BNP * 8 + 4 * hist * 8 + (32 + 16) * 4 + 3 * BNP      // x v | histories | lane starts x 2, flags | lane vec ndraw
             + (pool ? (size_t)BUNDLE_YCAP * 12 : 0);
  b = (b + 15) & ~(size_t)15;
  return b + 8;                                     // an odd number of 8-byte words: rollouts of a warp on different banks
}

// bytes of `word` (four slots) that are below slot `k` when the word holds slots [base, base + 4)
__device__ __forceinline__ unsigned below_mask(int base, int k) {
  const int m = k - base;
  return m >= 4 ? 0xffffffffu : (m <= 0 ? 0u : (1u << (8 * m)) - 1u);
}

template <bool kMerge, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 384 ? 2 : 1) bundle_kernel(const SceneDev sc, const CandDev* __restrict__ cands, const LaunchParams lp,
                                                        const BundleParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int np = BNP;
  const int t = threadIdx.x, ln = t & 31, w = t >> 5;
  const int RW = 1 << bp.rw_log2;
  const int rr = ln & (RW - 1);                          // rollout of the bundle
  const int sidx = w * (32 >> bp.rw_log2) + (ln >> bp.rw_log2);   // this thread among the NS slot threads of its rollout
  const int NS = bp.ns;
  const int n = sc.n;
  const int k = sc.thread_slot[sidx];                    // vehicle slot (sorted by behaviour class on the host); >= n: padding
  const int cand = blockIdx.y;
  const int64_t ridx = (int64_t)blockIdx.x * RW + rr;
  const CandDev cd = cands[cand];
  mcs_status* st_out = lp.status + ((size_t)cand * lp.count + ridx);
  const bool active = ridx < lp.count;
  if (!cd.ok) {      // setEgoAgentIdAndDrivingBehavior failed (runMTimes 0xc2935 runs nothing); uniform over the block
    if (active && sidx == 0) { mcs_status z; memset(&z, 0, sizeof z); *st_out = z; }
    return;
  }

  const double* F = sc.f;
  const int32_t* I = sc.i;

  // ---- shared memory: static rows once per block, then one region per rollout
  Smem s;
  {
    double* d = (double*)smem_raw;
    s.l = d; s.vd = d + np; s.idm = d + 2 * np; s.rss = d + 8 * np;
    unsigned char* p = smem_raw + BUNDLE_STATIC_BYTES + (size_t)rr * bp.region_bytes;
    const int hist = (lp.steps + 2) & ~1;
    s.x = (double*)p; p += 8 * np; s.v = (double*)p; p += 8 * np;
    s.chist = (double*)p; p += 8 * hist; s.csort = (double*)p; p += 8 * hist;
    s.vhist = (double*)p; p += 8 * hist; s.thist = (double*)p; p += 8 * hist;
    s.ycap = kMerge ? BUNDLE_YCAP : 0;
    s.yp0 = (double*)p; p += 8 * s.ycap;
    s.cnt = (int*)p; p += 32 * 4; s.flags = (int*)p; p += 16 * 4;
    s.ynext = (uint16_t*)p; p += 2 * s.ycap;
    s.lane = p; p += np; s.vec = p; p += np; s.pos = p; p += np;       // s.pos holds the per-vehicle draw counts here
    s.ykey = p; s.yflag = p + s.ycap;
    s.mob = nullptr; s.acc = nullptr; s.red = nullptr; s.scan = nullptr; s.tasks = nullptr; s.mflag = nullptr;
  }
  uint8_t* ndraw = s.pos;
  int* laneStart = s.cnt;        // [n_lanes + 1]; flips between s.cnt and s.cnt + 16 at every rebuild of this rollout

  for (int q = t; q < np; q += blockDim.x) {
    const bool lv = q < n;
    s.l[q] = lv ? F[(size_t)F_L * np + q] : 0.0; s.vd[q] = lv ? F[(size_t)F_VD * np + q] : 1.0;
    const double accMax = lv ? F[(size_t)F_IDM_ACCMAX * np + q] : 1.0, accCom = lv ? F[(size_t)F_IDM_ACCCOM * np + q] : -1.0;
    s.idm[q] = accMax;
    s.idm[np + q] = lv ? F[(size_t)F_IDM_MINDIS * np + q] : 0.0;
    s.idm[2 * np + q] = lv ? F[(size_t)F_IDM_ACCMIN * np + q] : 0.0;
    s.idm[3 * np + q] = accCom;
    s.idm[4 * np + q] = lv ? F[(size_t)F_IDM_THW * np + q] : 0.0;
    { const double sq = sqrt((-accMax) * accCom); s.idm[5 * np + q] = sq + sq; }   // idmAccFrontBack 0xebd2b, hoisted
    s.rss[q] = lv ? F[(size_t)F_RSS_RTEGO * np + q] : 0.0;
    s.rss[np + q] = lv ? F[(size_t)F_RSS_AMIN * np + q] : -1.0;
    s.rss[2 * np + q] = lv ? F[(size_t)F_RSS_AMAX * np + q] : -1.0;
  }

  // ---- per-vehicle registers
  const bool live = active && k < n;
  const uint64_t key = mcs_rollout_key(lp.seed + (uint64_t)cand, (uint64_t)(lp.first + ridx));
  double x = 0, y = 0, v = 0, vy = 0, a0 = 0;
  int lane = -1, beh = MCS_BEH_OTHER, commit = 0, keepStep = 0, targetLane = -1, inputIdx = 0;
  bool aggressive = false;
  const bool isEgo = live && k == cd.ego_slot;
  if (live) {
    x = F[(size_t)F_X * np + k]; y = F[(size_t)F_Y * np + k]; v = F[(size_t)F_V * np + k]; vy = F[(size_t)F_VY * np + k];
    a0 = F[(size_t)F_A * np + k];
    lane = I[(size_t)I_LANE * np + k]; beh = I[(size_t)I_BEH * np + k]; commit = I[(size_t)I_COMMIT * np + k];
    keepStep = I[(size_t)I_KEEPSTEP * np + k]; targetLane = I[(size_t)I_TARGETLANE * np + k];
    inputIdx = I[(size_t)I_INPUT * np + k];
    aggressive = (mcs_unit31(mcs_rand31(key, (uint64_t)inputIdx))) > 0.8;     // Agent ctor draw 0xf89e6
    if (isEgo) targetLane = cd.target_lane_id;
  }
  if (active) {
    // every cell of the rollout's arrays gets its initial value from the scene, padding cells included (lane 0xff, no draws)
    for (int q = sidx; q < np; q += NS) {
      const bool lv = q < n;
      s.x[q] = lv ? F[(size_t)F_X * np + q] : 0.0; s.v[q] = lv ? F[(size_t)F_V * np + q] : 0.0;
      const int lq = lv ? I[(size_t)I_LANE * np + q] : -1;
      s.lane[q] = (uint8_t)(lq < 0 ? 0xff : lq);
      s.vec[q] = sc.lane_vec[q];
      ndraw[q] = 0;
    }
    for (int q = sidx; q <= sc.n_lanes; q += NS) laneStart[q] = sc.lane_start[q];
    for (int q = sidx; q < 16; q += NS) s.flags[q] = 0;
  }
  __syncthreads();
  if (isEgo) {
    s.flags[2] = beh;                   // the ego's current behaviour: `exit` vehicles are yielding candidates (0x104c90)
    s.flags[4] = cd.gap_slot;           // Agent::targetGapId as a slot; < 0 = last gap
    s.flags[5] = lp.steps;              // < lp.steps once the ego has reached its target lane (StatusOneSim::finish)
    s.chist[0] = a0; s.vhist[0] = v;
  }
  __syncthreads();

  // per-object statistics (Simulation::run 0x10ee71-0x10efd6) accumulated on the fly in history order
  double accAbs = fabs(a0), accV = v;
  uint32_t drawCtr = (uint32_t)n;   // draws 0..n-1 were the ctor flags
  int step = 0, executed = 0;
  YieldChain yt;                    // this vehicle's chain of yielding intentions in the rollout's pool
  yt.init();
  bool running = active;

  while (true) {
    const bool run = running && lp.steps > step;
    // ================= plan, part A: searches, MOBIL geometry, draw counts =================
    double ax = 0.0, avy = 0.0;
    int front = -1;
    bool needMobil = false, noise = false;
    int mode = 0;   // 1 idmBehaviorAndYielding, 2 customizedLaneChange (ego), 3 customizedMerging (ego), 4 closest gap
    const bool hasLane = run && live && lane >= 0;
    double centerY = 0.0, vLimit = 0.0;
    int ycand[5];
    int nY = 0;
    MobilGeo geo;
    geo.accNewL = geo.accNewR = geo.dNewL = geo.dNewR = geo.deltaOld = 0.0; geo.mf = 0;
    bool mobilDraw = false;
    if (hasLane) {
      const LaneDev& L = sc.lanes[lane];
      centerY = L.centerY; vLimit = L.vLimit;
      if (isEgo) {
        const bool lr = cd.action == MCS_ACT_LEFT || cd.action == MCS_ACT_RIGHT;
        const bool reached = s.flags[0] != 0;
        if (beh == MCS_BEH_IDM || beh == MCS_BEH_IDM_LC) mode = (lr && reached) ? 1 : 2;
        else if (beh == MCS_BEH_MERGING || beh == MCS_BEH_EXIT) mode = reached ? 1 : 3;
        else s.flags[6] = 1;
      } else if (beh == MCS_BEH_IDM_LC) {
        mode = 1; noise = true;
        const bool canL = L.leftIdx >= 0 && sc.lanes[L.leftIdx].type == MCS_LANE_MAIN;
        const bool canR = L.rightIdx >= 0 && sc.lanes[L.rightIdx].type == MCS_LANE_MAIN;
        needMobil = !(commit == MCS_COMMIT_KEEP_LANE && keepStep <= 9) && (canL || canR);
      } else if (beh == MCS_BEH_IDM) {
        mode = 1; noise = true;
      } else if (beh == MCS_BEH_MERGING) {
        mode = 4;
      } else {
        s.flags[6] = 1;       // the binary pushes no action for such a vehicle and then reads past its action vector (0x10ed15)
      }
      if (kMerge) {
        if (mode == 1) {
          // yielding candidates: neighbours on a merging lane to the right (0x104469), `exit` vehicles on the main lane to
          // the left (0x1044df) — inside a rollout only the ego can still be an `exit` vehicle
          if (L.rightIdx >= 0 && sc.lanes[L.rightIdx].type == MCS_LANE_MERGING) nY = side_vector(s, laneStart, L.rightIdx, x, ycand);
          if (sc.has_exit_lanes && L.leftIdx >= 0 && sc.lanes[L.leftIdx].type == MCS_LANE_MAIN) {
            const int eg = cd.ego_slot;
            if (eg != k && s.lane[eg] == L.leftIdx && s.flags[2] == MCS_BEH_EXIT) ycand[nY++] = eg;
          }
        }
      } else if (mode == 3 || mode == 4) {
        s.flags[6] = 1;         // cannot happen: the host selects kMerge for such scenes
      }
      // own-lane leader the way the reference finds it: binary search over the (possibly stale-ordered) vector
      const int lo = laneStart[lane], cnt = laneStart[lane + 1] - lo;
      const int fpos = lane_upper(s, lo, cnt, x);
      front = fpos < cnt ? s.vec[lo + fpos] : -1;
      if (needMobil) {
        const int bpos = lane_reverse(s, lo, cnt, x);
        const int back = bpos > 0 ? s.vec[lo + bpos - 1] : -1;
        geo = mobil_geometry_all(s, sc, laneStart, np, k, front, back);
        const int laneId = L.id;
        const bool contL = commit == MCS_COMMIT_LEFT && targetLane != laneId && (geo.mf & MF_RELAX_L);
        const bool contR = commit == MCS_COMMIT_RIGHT && targetLane != laneId && (geo.mf & MF_RELAX_R);
        mobilDraw = !(contL || contR);
      }
      // The gap behaviours of the ego (mode 3) and of the ramp vehicles (mode 4) consume no draws and read only the
      // pre-step state: they run here, next to the MOBIL sub-tasks of the other warps, not after the barrier where the
      // yielding vehicles' chain (IDM, intentions, decision) is the long one.
      if (kMerge && (mode == 3 || mode == 4)) {
        const double el = s.l[k];
        MergeSelf me;
        me.k = k; me.lane = lane; me.front = front; me.x = x; me.y = y; me.v = v; me.l = el;
        me.w = F[(size_t)F_W * np + k];
        me.rss.rTOther = F[(size_t)F_RSS_RTOTHER * np + k]; me.rss.rTEgo = s.rss[k]; me.rss.aMin = s.rss[np + k]; me.rss.aMax = s.rss[2 * np + k];
        me.rss.dSoft = F[(size_t)F_RSS_DSOFT * np + k];
        me.idm = load_idm(s, np, k);
        // mode 3: the ego executes its gap (customizedMerging); mode 4: a non-ego `merging` vehicle picks the closest gap to the left
        const bool okSide = merging_behavior(s, sc, laneStart, me, s.vd[k], mode == 3, false, mode == 3 ? cd.action == MCS_ACT_LEFT : true,
                                             mode == 3 ? s.flags[4] : -1, ax, avy);
        if (!okSide) { s.flags[6] = 1; ax = 0.0; avy = 0.0; }
      }
    }
    // draw counts: [yielding candidates][noise][MOBIL decision draw], in slot (= the reference's visit) order
    if (run) ndraw[k] = (uint8_t)(hasLane ? nY + (noise ? 1 : 0) + (mobilDraw ? 1 : 0) : 0);
    if (!__syncthreads_or(run ? 1 : 0)) break;

    // ================= plan, part B: actions =================
    if (run) {
      unsigned before = 0, total = 0;
      const unsigned* dw = (const unsigned*)ndraw;
#pragma unroll 4
      for (int q = 0; q < np / 4; ++q) {
        const unsigned word = dw[q];
        total += __vsadu4(word, 0u);
        before += __vsadu4(word & below_mask(4 * q, k), 0u);
      }
      const uint64_t drawBase = (uint64_t)drawCtr + before;
      drawCtr += total;
      if (sidx == 0) s.flags[1] = 0;       // "a lane membership changed": set in the match below, read after its barrier
      if (hasLane) {
        const IdmP pe = load_idm(s, np, k);
        const double el = s.l[k], evd = s.vd[k];
        const double rTEgo = s.rss[k], aMin = s.rss[np + k], aMax = s.rss[2 * np + k];
        if (isEgo) {    // Agent::thwRatioHistory is read for the ego only (0x10f14b)
          double thwPush = 1.0;
          if (front >= 0) thwPush = (((s.x[front] - x) - (s.l[front] + el) * 0.5) / v) / pe.thw;
          s.thist[s.flags[3]++] = thwPush;
        }
        if (mode == 1 || mode == 2) {
          avy = centering_vy(y, centerY);

          // One idmAccFrontBack call per vehicle and step: the branches only choose its operands (rollout_kernel.cuh)
          IdmP p = pe;
          int f0 = front, f1 = -1, bk = -1;
          double vdd = evd;
          bool evalIdm = true, lateral = false;
          const int cmd = cd.action;
          if (mode == 1) {
            if (isEgo) vdd = MCS_MINSD(evd, 1.1 * vLimit);
          } else {
            const double vdCap = MCS_MINSD(evd, 1.1 * vLimit);
            vdd = vdCap;
            if (cmd == MCS_ACT_DCC) { p.thw = pe.thw + pe.thw; vdd = 0.8 * vdCap; }
            else if (cmd == MCS_ACT_ACC) { p.thw = 0.9; vdd = 1.1 * vdCap; }
            else if (cmd != MCS_ACT_KEEP_LANE) {
              const LaneDev& L = sc.lanes[lane];
              const int sl = cmd == MCS_ACT_LEFT ? L.leftIdx : L.rightIdx;
              if (sl >= 0 && sc.lanes[sl].type == MCS_LANE_MAIN) {
                int leader, follower;
                side_leader_follower(s, laneStart, sl, x, false, leader, follower);
                f0 = leader; f1 = front; bk = follower; lateral = true;
              } else {
                evalIdm = false;      // no main lane on that side: ax = 0 unless the RSS emergency fires (0xf6976)
              }
            }
          }
          const double pw = mcs_pow4(v / vdd);
          double axFront = 0.0;
          if (evalIdm) axFront = idm_acc(s, p, f0, f1, bk, x, v, el, pw);
          const bool emergency = front >= 0 && rss_emergency(s, front, x, v, el, rTEgo, aMin, aMax);
          if (mode == 2) {
            // for left/right the binary's emergency select reads an uninitialised stack slot (0xf6263); only the aMin outcome is observable
            ax = emergency ? aMin : axFront;
            if (lateral && rss_safe(s, f0, bk, x, v, el, F[(size_t)F_RSS_RTOTHER * np + k], rTEgo, aMin, aMax)) {
              const double m = MCS_MINSD(1.0, 0.17 * v);
              avy = m * (cmd == MCS_ACT_LEFT ? 1.0 : -1.0);
            }
          } else {
            IdmP py = pe;
            py.accMax = F[(size_t)F_IDMY_ACCMAX * np + k]; py.accMin = F[(size_t)F_IDMY_ACCMIN * np + k];
            double axY;
            if (kMerge && nY > 0) {
              axY = yielding_acc(s, F, np, k, py, ycand, nY, key, drawBase, yt, x, v, el, pw);
            } else {
              // no yielding candidate: idmAccFrontBack over an empty list is the clamped free-road term
              const double r = MCS_MINSD(0.0 + (1.0 - pw) * py.accMax, py.accMax);
              axY = MCS_MAXSD(py.accMin, r);
            }
            ax = MCS_MINSD(axY, axFront);
            if (noise) ax = ax + (mcs_unit31(mcs_rand31(key, drawBase + (uint64_t)nY)) - 0.5);
            if (emergency) ax = aMin;
            if (beh == MCS_BEH_IDM_LC && !isEgo) {
              // laneChangeBehaviorMOBIL 0xf42d0, decision part
              if (commit == MCS_COMMIT_KEEP_LANE && keepStep <= 9) {
                keepStep = keepStep + 1;
              } else if (needMobil) {
                const double u = mcs_unit31(mcs_rand31(key, drawBase + (uint64_t)nY + 1));
                MobilState ms{commit, keepStep, targetLane};
                mobil_decide(geo, sc, np, k, lane, v, el, aMin, F[(size_t)F_YP_PF * np + k], F[(size_t)F_YP_DA * np + k],
                             F[(size_t)F_YP_BA * np + k], /*mode=*/(step == 0) ? 0 : 1, aggressive, F + (size_t)F_PRIOR_L * np, u, ms, ax, avy);
                commit = ms.commit; keepStep = ms.keepStep; targetLane = ms.targetLane;
              }
            }
          }
        }
      }
    }
    __syncthreads();   // every read of the pre-step x / v is done

    // ================= integrate (0x10ea36-0x10eb80) and matchAgentsOnCorridor (0x10b810) =================
    bool changed = false;
    if (run && live) {
      double acc;
      double v1 = ax * lp.dt + v;
      if (0.0 > v1) {
        v = 0.0; x = 0.0 * lp.dt + x; vy = avy; y = avy * lp.dt + y; acc = 0.0;
      } else {
        v = v1; x = v1 * lp.dt + x; vy = avy; y = avy * lp.dt + y; acc = (v1 == 0.0) ? 0.0 : ax;
      }
      s.x[k] = x; s.v[k] = v;
      accAbs = accAbs + fabs(acc); accV = accV + v;
      if (isEgo) { s.chist[step + 1] = acc; s.vhist[step + 1] = v; }
      bool stay = false;
      if (lane >= 0) { const LaneDev& L = sc.lanes[lane]; stay = L.leftY >= y && y >= L.rightY; }
      if (!stay) {
        changed = true;
        for (int j = 0; j < sc.n_lanes; ++j) {
          const LaneDev& L = sc.lanes[j];
          if (!(L.leftY >= y && y >= L.rightY)) continue;
          lane = j;
          if (L.type == MCS_LANE_MAIN && beh == MCS_BEH_MERGING) beh = MCS_BEH_IDM_LC;
          if (L.type == MCS_LANE_EXIT && beh == MCS_BEH_EXIT) beh = MCS_BEH_IDM;
          break;
        }
        s.lane[k] = (uint8_t)(lane < 0 ? 0xff : lane);
        s.flags[1] = 1;
      }
      if (isEgo && lane >= 0 && s.flags[0] == 0) s.flags[0] = sc.lanes[lane].id == targetLane ? 1 : 0;
      if (kMerge && isEgo) s.flags[2] = beh;
    }
    // one barrier publishes the new x / v / lane and the ego's arrival, and tells whether any rollout of the bundle has to
    // rebuild its lane vectors
    if (__syncthreads_or(changed ? 1 : 0)) {
      // The reference rebuilds every lane vector (members in visit order) and std::sorts it by x (0x10be30-0x10c1a5): the
      // result is "members ordered by (x, slot)".  Position of a vehicle = members of lanes before its own + members of
      // its own lane with a smaller key, both counted over the rollout's 64 lane bytes four at a time.
      const bool rebuild = run && s.flags[1] != 0;
      int* nextStart = laneStart == s.cnt ? s.cnt + 16 : s.cnt;
      if (rebuild) {
        const unsigned* lw = (const unsigned*)s.lane;
        const int myLane = live ? (int)s.lane[k] : 0xff;
        if (myLane != 0xff) {
          const unsigned pat = 0x01010101u * (unsigned)myLane;
          int posn = 0;
          for (int q = 0; q < np / 4; ++q) {
            const unsigned word = lw[q];
            posn += __popc(__vcmpltu4(word, pat)) >> 3;
            unsigned eq = __vcmpeq4(word, pat) & 0x01010101u;
            while (eq) {
              const int b = (__ffs(eq) - 1) >> 3;
              eq &= eq - 1;
              const int j = 4 * q + b;
              const double xj = s.x[j];
              posn += (xj < x || (xj == x && j < k)) ? 1 : 0;
            }
          }
          s.vec[posn] = (uint8_t)k;
        }
        for (int l = sidx; l <= sc.n_lanes; l += NS) {          // start of lane l = vehicles on lanes before it
          const unsigned pat = 0x01010101u * (unsigned)l;
          int c = 0;
          for (int q = 0; q < np / 4; ++q) c += __popc(__vcmpltu4(lw[q], pat)) >> 3;
          nextStart[l] = c;
        }
      }
      __syncthreads();
      if (rebuild) laneStart = nextStart;
    }

    // ================= termination (0x10edd0-0x10ee10) =================
    if (run) {
      const int gapSlot = (kMerge && isEgo) ? s.flags[4] : -1;
      if (kMerge && isEgo && gapSlot >= 0) {
        // the gap's rear vehicle left the target lane: re-target to the last vehicle of that lane behind it, else the
        // last gap (matchAgentsOnCorridor 0x10c320-0x10c39d)
        const int gl = s.lane[gapSlot];
        if (gl == 0xff) s.flags[6] = 1;
        else if (sc.lanes[gl].id != targetLane) {
          int tl = -1;
          for (int j = 0; j < sc.n_lanes; ++j) if (sc.lanes[j].id == targetLane) tl = j;
          if (tl < 0) s.flags[6] = 1;
          else {
            const int lo = laneStart[tl], cnt = laneStart[tl + 1] - lo;
            const double gx = s.x[gapSlot];
            int found = -1;
            for (int q = cnt - 1; q >= 0; --q) {
              const int o = s.vec[lo + q];
              if (gx > s.x[o]) { found = o; break; }
            }
            s.flags[4] = found;
          }
        }
      }
      executed++;
      if (s.flags[0] != 0) {        // the ego has reached its target lane
        if (isEgo) s.flags[5] = min(s.flags[5], step);
        if (lp.min_steps <= step) running = false;
      }
    }
    step++;
  }

  // ================= outcome statistics (0x10ee20-0x10f64d) =================
  // (the loop was left through a barrier: every rollout of the bundle is done with its x / v arrays, which now carry the
  // per-object utility / comfort to the ego's thread)
  const int nStates = executed + 1;
  const int egoSlot = cd.ego_slot;
  const double egoAMin = F[(size_t)F_RSS_AMIN * np + egoSlot], egoVd = s.vd[egoSlot];
  if (live && !isEgo) {
    const double cnt = (double)nStates, evd = s.vd[k], aMin = s.rss[np + k];
    const double meanV = accV / cnt;
    s.v[k] = (accAbs / cnt) / aMin + 1.0;                    // comfort of the object
    s.x[k] = (meanV >= evd) ? 1.0 : meanV / evd;             // utility of the object
  }
  // ego history -> per-state terms, one state per thread: fall-back = two consecutive states at rss.aMin (0x10f797)
  if (active) {
    bool fb = false;
    for (int q = sidx; q < nStates; q += NS) {
      const double a = s.chist[q];
      if (q + 1 < nStates && egoAMin == a && egoAMin == s.chist[q + 1]) fb = true;
    }
    if (fb) s.flags[9] = 1;
  }
  __syncthreads();
  if (active) {      // utility term u(v/vd) (0x10f0c4), comfort sample (0x10f1c3)
    for (int q = sidx; q < nStates; q += NS) {
      const double a = s.chist[q];
      s.chist[q] = a > 0.0 ? fabs(a) * 0.5 : fabs(a);
      const double ve = s.vhist[q];
      const double r = ve / egoVd;
      double u = r;
      if (!(egoVd >= ve)) u = MCS_MAXSD(0.9, 1.0 - (r - 1.0));
      s.vhist[q] = u;
    }
  }
  __syncthreads();
  if (active) {      // the comfort samples in descending order (0x10f1c3-0x10f49a), by rank
    for (int q = sidx; q < nStates; q += NS) {
      const double c = s.chist[q];
      int rank = 0;
#pragma unroll 1
      for (int j = 0; j < nStates; ++j) { const double cj = s.chist[j]; rank += (cj > c || (cj == c && j < q)) ? 1 : 0; }
      s.csort[rank] = c;
    }
  }
  __syncthreads();
  if (isEgo) {
    const double aMin = s.rss[np + k];
    double uo = 1e300, co = 1e300;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      if (j == k) continue;
      uo = MCS_MINSD(s.x[j], uo); co = MCS_MINSD(s.v[j], co);
    }
    if (n <= 1) { uo = 1.0; co = 1.0; }
    mcs_status st;
    memset(&st, 0, sizeof st);
    st.utilityObjects = uo; st.comfortObjects = co;
    double sumU = 0.0;
#pragma unroll 1
    for (int q = 0; q < nStates; ++q) sumU = sumU + s.vhist[q];
    const double util = sumU / (double)nStates;
    const int thwCnt = s.flags[3];
    double thwAcc = 0.0;                                  // 0x10f14b-0x10f19f
#pragma unroll 1
    for (int q = 0; q < thwCnt; ++q) { const double e = s.thist[q]; if (0.9 > e && e > 0.0) thwAcc = (e + thwAcc) - 0.9; }
    const double thwTerm = thwCnt > 0 ? thwAcc / (double)thwCnt : 0.0;
    st.utility = util + thwTerm;
    const int kk = (int)((double)nStates * 0.2);
    double sumC = 0.0;
#pragma unroll 1
    for (int q = 0; q < kk; ++q) sumC = sumC + (s.csort[q] / aMin + 1.0);
    st.comfort = sumC / (double)kk;
    st.finishStep = max(s.flags[5], (int)(1.0 / lp.dt)); st.executedSteps = executed; st.draws = (uint32_t)drawCtr;
    st.finish = s.flags[5] < lp.steps ? 1 : 0;
    const int beh0 = I[(size_t)I_BEH0 * np + k];
    if (beh0 == MCS_BEH_MERGING || beh0 == MCS_BEH_EXIT) st.fallBack = (2.0 > v) ? 1 : 0;
    else st.fallBack = s.flags[9] ? 1 : 0;
    st.ok = 1;
    const int anyUnsupported = s.flags[6], anyFull = s.flags[7];
    st.overflow = (uint8_t)(anyUnsupported ? 2 : (anyFull ? 1 : 0));
    if ((anyUnsupported || anyFull) && lp.err_flags) atomicOr(lp.err_flags, anyUnsupported ? 2u : 1u);
    *st_out = st;
    if (lp.vehicle_steps) atomicAdd(lp.vehicle_steps, (unsigned long long)executed * (unsigned long long)n);
  }
}

}  // namespace mcs

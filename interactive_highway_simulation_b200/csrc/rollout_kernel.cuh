// Rollout kernel for sm_100a: one thread per vehicle, NPS = 32 / 64 / 128 / 256 threads per Monte-Carlo rollout
// (a template constant).  A 256-vehicle rollout owns its thread block; smaller rollouts share a 256- or 512-thread block,
// each on its own named barrier and shared-memory slice, and start every step together so their warps share instruction
// fetches (the kernel body is ~5 K instructions, far beyond the 32 KB L1.5 instruction cache).
//
// Replaces the reference's Simulation::run (mc_sim_highway.so 0x10e790) and everything it calls:
//   idmAccFrontBack 0xebc30, idmBehaviorAndYieldingToMergingAgents 0x104200, laneChangeBehaviorMOBIL 0xf42d0,
//   laneChangeProbability 0xe19f0, rssSafe 0xb8950, customizedLaneChangeBehavior 0xf5e30,
//   Environment::{matchAgentsOnCorridor 0x10b810, frontAgent 0xe2670, followingAgent 0xe2920, neighborAgents 0xe8240};
//   with kMerge also customizedMergingBehavior 0xf6ba0, mergingBehaviorClosestGap 0xfe660, rssSafeRelax 0xb9210,
//   timeToReachGap 0xe1630 and the yielding intentions (merge_behaviors.cuh).
//
// Design (B200): the path is FP64-latency / instruction-issue bound, not HBM bound — a rollout reads the scene once
// (L2-resident, shared by every block) and writes one 64-byte record.  Dynamic vehicle state (x y v vy, commit state)
// lives in registers; what neighbours gather (x, v) and everything static per vehicle (l, vd, IDM and RSS parameters)
// lives in shared memory — registers, not shared memory, limit the resident rollouts per SM.  The reference's per-lane
// std::vector<AgentPtr> (sorted only when lane membership changes, searched by binary search every step) is a byte
// array of slots per lane in shared memory with exactly those semantics, including stale order; while every vector is
// strictly ordered the searches collapse to neighbour lookups.  MOBIL (most of the arithmetic, needed by ≈ 1 vehicle in
// 11 per step) is compacted: vehicles that need it queue in shared memory and their evaluation runs as seven independent
// sub-tasks on seven warps side by side.  Random draws come from the counter stream of include/mcsim_rng.h at the index
// the reference's sequential loop would have consumed them (exclusive scan of per-vehicle draw counts).
//
// Arithmetic is written in the operation order of the reference machine code and compiled with
// -fmad=false: results are bit-identical to the reference binary given the same exp/pow (mcsim_math.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcsim.h"
#include "../../include/mcsim_math.h"
#include "../../include/mcsim_rng.h"
#include "scene.hpp"

namespace mcs {

struct LaneDev {
  double leftY, rightY, centerY, vLimit, endX, startX;
  int id, type, leftIdx, rightIdx, leftmost, nb /* LANE_NB_*: what the neighbour lanes are (lane_neighbour_bits) */;
};
enum { LANE_NB_LEFT_MAIN = 1, LANE_NB_RIGHT_MAIN = 2, LANE_NB_RIGHT_MERGING = 4 };

struct SceneDev {
  int n, n_pad, n_lanes, has_exit_lanes;
  const double* f;          // F_COUNT * n_pad
  const int32_t* i;         // I_COUNT * n_pad
  const uint8_t* lane_vec;  // n_pad
  const uint8_t* thread_slot;   // n_pad: vehicle slot handled by each thread of a rollout (capi.cu thread_slots)
  int lane_start[MCS_MAX_LANES + 1];
  LaneDev lanes[MCS_MAX_LANES];
};

struct SceneDev;
// per lane, once per scene: is the left / right neighbour a main lane, is the right one a merging lane — the plan asks every
// vehicle every step, and the answer through lanes[lanes[l].leftIdx].type is two dependent constant-bank loads
template <class Lanes>
__host__ __device__ inline void lane_neighbour_bits(Lanes& lanes, int n_lanes) {
  for (int j = 0; j < n_lanes; ++j) {
    int nb = 0;
    if (lanes[j].leftIdx >= 0 && lanes[lanes[j].leftIdx].type == MCS_LANE_MAIN) nb |= LANE_NB_LEFT_MAIN;
    if (lanes[j].rightIdx >= 0 && lanes[lanes[j].rightIdx].type == MCS_LANE_MAIN) nb |= LANE_NB_RIGHT_MAIN;
    if (lanes[j].rightIdx >= 0 && lanes[lanes[j].rightIdx].type == MCS_LANE_MERGING) nb |= LANE_NB_RIGHT_MERGING;
    lanes[j].nb = nb;
  }
}

struct CandDev { int ego_slot, action, gap_id, target_lane_id, ok, gap_slot /* slot of gap_id, -1 = last gap */, pad[2]; };

struct LaunchParams {
  double dt;
  int steps, min_steps;
  uint64_t seed;      // stream of candidate c = mcs_rollout_key(seed + c, rollout)
  int64_t first;      // first rollout index
  int64_t count;      // rollouts per candidate in this launch
  int n_cand;
  mcs_status* status; // [n_cand][count]
  double* traj;       // optional [n_cand][count][n][steps+1][4] in INPUT agent order, or nullptr
  int32_t* traj_n;    // optional [n_cand][count]
  unsigned long long* vehicle_steps;  // optional global counter
  unsigned int* err_flags;            // per-launch word: bit 0 = a yielding table overflowed, bit 1 = scene unsupported (any rollout)
};

#define MCS_MINSD(a, b) (((a) < (b)) ? (a) : (b))   /* x86 vminsd: second source when not less / unordered */
#define MCS_MAXSD(a, b) (((a) > (b)) ? (a) : (b))

// num / den, bit-identical to IEEE division.  A zero numerator (ubiquitous here: clamped IDM interaction terms,
// probabilities of gated-off lane changes) sends the hardware division sequence into its out-of-line slow path — 9 % of all
// executed instructions in the r01h profile.  A plain `if (num == 0) return num` does not help: the compiler evaluates the
// division speculatively.  Instead the division itself never sees a zero numerator, and the exact quotient of a zero
// numerator is selected afterwards: +-0 with the sign of num for den > 0, the opposite sign for den < 0, NaN for 0 / NaN.
__device__ __forceinline__ double mcs_div(double num, double den) {
  const bool zero = num == 0.0;
  double safe = zero ? 1.0 : num;
  // opaque to the optimiser: otherwise it may move the select across the division (it did at four call sites — the
  // stand-in 1.0 ended up in the denominator and 6 % of all executed instructions were still the slow path, r01n)
  asm volatile("" : "+d"(safe));
  const double q = safe / den;
  if (!zero) return q;
  if (den > 0.0) return num;
  if (den < 0.0) return -num;
  return q * 0.0;       // den is 0 or NaN: 1/den is inf or NaN, times zero is the NaN that 0/den would be
}

// r / 2147483647.0 for a 31-bit draw r (the reference's (double)rand() / RAND_MAX) without the division sequence: with
// y = RN(1 / c) the product r * y corrected once by its exact remainder, q = fma(fma(-q0, c, r), y, q0), is the correctly
// rounded quotient — for EVERY r in [0, 2^31), checked exhaustively on the host (tests/test_host_logic.py
// test_unit_draw_without_division, tests/unit31_check.c).  3 instructions instead of ~14, once per vehicle and step.
__device__ __forceinline__ double mcs_unit31(int r) {
  const double c = 2147483647.0, y = 1.0 / 2147483647.0;
  const double a = (double)r, q0 = a * y;
  return __fma_rn(__fma_rn(-q0, c, a), y, q0);
}

struct IdmP { double accMax, minDis, accMin, accCom, thw, sq2 /* 2*sqrt(-accMax*accCom), static per vehicle */; };

// shared-memory view of one rollout
struct Smem {
  double* x;       // [n_pad] current longitudinal position
  double* v;       // [n_pad]
  double* l;       // [n_pad] static
  double* vd;      // [n_pad] static
  double* idm;     // [6][n_pad] static: accMax minDis accMin accCom thw 2*sqrt(-accMax*accCom)
  double* mob;     // [6][n_pad] MOBIL geometry results per vehicle: accNewL accNewR dNewL dNewR deltaOld (+1 spare)
  double* rss;     // [3][n_pad] static: rTEgo aMin aMax (read once per step by the RSS emergency test)
  double* acc;     // [2][n_pad] per-object statistics: sum |a_t|, sum v_t
  double* chist;   // [steps+2] ego acceleration history, turned into comfort samples at the end
  double* csort;   // [steps+2] the comfort samples sorted descending
  double* vhist;   // [steps+2] ego speed history, turned into per-state utilities at the end
  double* thist;   // [steps+2] ego thwRatioHistory
  double* red;     // [16] reduction scratch
  double* ego;     // [4] helper-warp hand-over of the ego (kHelper 1): its y; its planned ax, vy and "a lane on that side exists"
  // helper warp with ramp vehicles (kHelper 2): what its lanes read of and answer for the vehicles they serve
  double* hy;      // [n_pad] lateral position of every vehicle (the owner's register copy, published at every integration)
  double* hax;     // [n_pad] planned acceleration of a served vehicle
  double* havy;    // [n_pad] planned lateral speed
  uint8_t* hbeh;   // [n_pad] current behaviour of every vehicle
  uint8_t* hok;    // [n_pad] 0xff = not served by the helper warp; else "a lane on that side exists" of the last plan
  int* scan;       // [32] warp totals of the two per-step scans (8 + 8), lane counters of the re-ranking (16..)
  int* cnt;        // [2][16] lane starts, double-buffered across rebuilds
  int* flags;      // [8] per-rollout scalars: 0 ego reached its target lane, 1 movers queued, 2 ego behaviour, 3 thw samples,
                   //     4 ego's gap slot, 5 finish step, 6 scene unsupported, 7 yield table full
                   //     (bundle_kernel: [16], 1 = a lane membership changed, 8 = yield pool entries in use, 9 = fall-back seen)
  int* tasks;      // [n_pad] MOBIL task queue
  uint8_t* lane;   // [n_pad] lane slot of each vehicle (0xff = none)
  uint8_t* vec;    // [n_pad] concatenated per-lane vectors
  uint8_t* mflag;  // [2][n_pad] MOBIL flags per vehicle (left part, right part)
  uint8_t* pos;    // [n_pad] position of each vehicle inside vec
  // yielding intentions in bundle_kernel: one pool of entries per rollout, chained per vehicle (merge_behaviors.cuh)
  double* yp0;      // [ycap] probability at the first encounter
  uint16_t* ynext;  // [ycap] next entry of the same yielding vehicle, YIELD_END = end of the chain
  uint8_t* ykey;    // [ycap] slot of the vehicle yielded to
  uint8_t* yflag;   // [ycap] current intention
  int ycap;
};
constexpr int YIELD_END = 0xffff;
// entries of a rollout's pool in bundle_kernel (<= 64 vehicles; a vehicle next to a ramp meets the ramp's [BB, B, F, FF]
// and, while it overtakes or is overtaken, a few more; vehicles elsewhere meet nobody)
constexpr int BUNDLE_YCAP = 256;

// shared-memory layout of one rollout (host twin: rollout_smem_bytes in capi.cu)
__device__ __forceinline__ void carve_smem(Smem& s, unsigned char* p, int np, int steps, bool helper = false) {
  s.x = (double*)p; p += sizeof(double) * np;
  s.v = (double*)p; p += sizeof(double) * np;
  s.l = (double*)p; p += sizeof(double) * np;
  s.vd = (double*)p; p += sizeof(double) * np;
  s.idm = (double*)p; p += sizeof(double) * 6 * np;
  s.mob = (double*)p; p += sizeof(double) * 6 * np;
  s.rss = (double*)p; p += sizeof(double) * 3 * np;
  s.acc = (double*)p; p += sizeof(double) * 2 * np;
  s.hy = s.hax = s.havy = nullptr; s.hbeh = s.hok = nullptr;
  if (helper) {
    s.hy = (double*)p; p += sizeof(double) * np;
    s.hax = (double*)p; p += sizeof(double) * np;
    s.havy = (double*)p; p += sizeof(double) * np;
  }
  s.red = (double*)p; p += sizeof(double) * 16;
  s.ego = (double*)p; p += sizeof(double) * 4;
  s.scan = (int*)p; p += sizeof(int) * 32;
  s.cnt = (int*)p; p += sizeof(int) * 32;
  s.flags = (int*)p; p += sizeof(int) * 8;
  s.tasks = (int*)p; p += sizeof(int) * np;
  s.lane = p; p += np;
  s.vec = p; p += np;
  s.mflag = p; p += 2 * np;
  s.pos = p; p += np;      // 5 * np bytes so far, np a multiple of 32: still 8-byte aligned
  if (helper) { s.hbeh = p; p += np; s.hok = p; p += np; }
  s.ycap = 0; s.yp0 = nullptr; s.ynext = nullptr; s.ykey = nullptr; s.yflag = nullptr;
  // the arrays whose size depends on the horizon come last: everything above sits at a compile-time offset (np is the
  // block-size template constant), so the per-step accesses need no address arithmetic
  s.chist = (double*)p; p += sizeof(double) * ((steps + 2) & ~1);
  s.csort = (double*)p; p += sizeof(double) * ((steps + 2) & ~1);
  s.vhist = (double*)p; p += sizeof(double) * ((steps + 2) & ~1);
  s.thist = (double*)p; p += sizeof(double) * ((steps + 2) & ~1);
}

__device__ __forceinline__ IdmP load_idm(const Smem& s, int np, int k) {
  IdmP p;
  p.accMax = s.idm[k]; p.minDis = s.idm[np + k]; p.accMin = s.idm[2 * np + k]; p.accCom = s.idm[3 * np + k];
  p.thw = s.idm[4 * np + k]; p.sq2 = s.idm[5 * np + k];
  return p;
}

// idmAccFrontBack 0xebc30: fronts (slots, -1 = null), back slot or -1.  pw = pow(egoV/vd, 4) supplied by the caller.
#ifndef MCS_HEAVY_INLINE
#define MCS_HEAVY_INLINE __forceinline__
#endif
__device__ MCS_HEAVY_INLINE double idm_acc(const Smem& s, const IdmP& p, int f0, int f1, int back, double egoX,
                                          double egoV, double egoL, double pw) {
  const double freeAcc = (1.0 - pw) * p.accMax;
  const double na = -p.accMax;
  const double sq2 = p.sq2;
  const double dDes = (egoV * p.thw + p.minDis) * 0.7;
  bool any = false;
  double d = 0.0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int f = q == 0 ? f0 : f1;
    if (f < 0) continue;
    const double gap = (s.x[f] - egoX) - (egoL + s.l[f]) * 0.5;
    double term;
    if (gap > 0.0) {
      const double dv = (egoV - s.v[f]) * egoV;
      double t = dv / sq2 + dDes;
      t = MCS_MAXSD(t, 0.0);
      t = mcs_div(t, gap);
      term = (t * t) * na;
    } else {
      term = -10000000000.0;
    }
    if (!any) { d = term; any = true; }
    else d = MCS_MINSD(term, d);
  }
  double sum = any ? d + 0.0 : 0.0;
  if (back >= 0) {
    const double bv = s.v[back];
    const double gap = (s.x[back] - egoX) - (egoL + s.l[back]) * 0.5;
    if (0.0 > gap) {
      const double dDesB = (bv * p.thw + p.minDis) * 0.7;
      const double dv = (bv - egoV) * egoV;
      double t = dv / sq2 + dDesB;
      t = MCS_MAXSD(t, 0.0);
      t = mcs_div(t, gap);
      sum = sum + (t * t) * p.accMax;
    } else {
      sum = sum + 10000000000.0;
    }
  }
  const double r = MCS_MINSD(sum + freeAcc, p.accMax);
  return MCS_MAXSD(p.accMin, r);
}

// std::upper_bound over lane vector [lo, lo+n): first position whose x > key (frontAgent 0xe27e0-0xe2820)
__device__ __forceinline__ int lane_upper(const Smem& s, int lo, int n, double key) {
  int first = 0, len = n;
  while (len > 0) {
    const int half = len >> 1, mid = first + half;
    if (s.x[s.vec[lo + mid]] > key) len = half;
    else { first = mid + 1; len = len - half - 1; }
  }
  return first;
}
// std::lower_bound: first position whose x >= key (neighborAgents 0xe8680-0xe86b1)
__device__ __forceinline__ int lane_lower(const Smem& s, int lo, int n, double key) {
  int first = 0, len = n;
  while (len > 0) {
    const int half = len >> 1, mid = first + half;
    if (s.x[s.vec[lo + mid]] >= key) len = half;
    else { first = mid + 1; len = len - half - 1; }
  }
  return first;
}
// reverse-iterator upper_bound with key > elem (followingAgent 0xe2aa8-0xe2ad4): answer is element first-1, 0 = none
__device__ __forceinline__ int lane_reverse(const Smem& s, int lo, int n, double key) {
  int first = n, len = n;
  while (len > 0) {
    const int half = len >> 1, mid = first - half;
    if (key > s.x[s.vec[lo + mid - 1]]) len = half;
    else { first = mid - 1; len = len - half - 1; }
  }
  return first;
}

// rssSafe 0xb8950 with optionals given as slots (-1 = nullopt); rTOther/rTEgo/aMin/aMax of the evaluating vehicle
__device__ MCS_HEAVY_INLINE bool rss_safe(const Smem& s, int leader, int follower, double ex, double ev, double el,
                                         double rTOther, double rTEgo, double aMin, double aMax) {
  if (leader < 0 && follower < 0) return true;
  if (leader >= 0) {
    const double fGap = (s.x[leader] - ex) - (s.l[leader] + el) * 0.5;
    const double fV = s.v[leader];
    double d = (ev * ev) / (aMin * -2.0) + ev * rTEgo;
    d = d + (fV * fV) / (aMax + aMax);
    d = MCS_MAXSD(0.0, d);
    if (d > fGap) return false;
    if (follower < 0) return true;
  }
  const double bGap = (s.l[follower] + el) * 0.5 + (s.x[follower] - ex);
  const double bV = s.v[follower];
  double d = (bV * bV) / (aMin * -2.0) + bV * rTOther;
  d = d + (ev * ev) / (aMax + aMax);
  d = MCS_MAXSD(0.0, d);
  return !(d > -bGap);
}

// rssSafe 0xb8950 twice for the same leader / follower — with the vehicle's own RSS parameters (bit 0 of the result) and
// with the relaxed ones (aMin -10, aMax -1; bit 1), as laneChangeBehaviorMOBIL asks for both.  One pass: the gaps and
// speeds are fetched once and the four stopping distances are independent chains that overlap in the pipeline, where two
// calls of rss_safe walk up to eight divisions one behind the other (a MOBIL sub-task's time is the length of its chain).
// Each distance is the reference's own expression, operation for operation.
__device__ MCS_HEAVY_INLINE int rss_safe_both(const Smem& s, int leader, int follower, double ex, double ev, double el,
                                              double rTOther, double rTEgo, double aMin, double aMax) {
  if (leader < 0 && follower < 0) return 3;
  const double ev2 = ev * ev;
  bool failS = false, failR = false;
  if (leader >= 0) {
    const double fGap = (s.x[leader] - ex) - (s.l[leader] + el) * 0.5;
    const double fV = s.v[leader];
    const double fV2 = fV * fV, evT = ev * rTEgo;
    double dS = ev2 / (aMin * -2.0) + evT;
    double dR = ev2 / (-10.0 * -2.0) + evT;
    dS = dS + fV2 / (aMax + aMax);
    dR = dR + fV2 / (-1.0 + -1.0);
    dS = MCS_MAXSD(0.0, dS);
    dR = MCS_MAXSD(0.0, dR);
    failS = dS > fGap; failR = dR > fGap;
  }
  if (follower >= 0) {
    const double bGap = (s.l[follower] + el) * 0.5 + (s.x[follower] - ex);
    const double bV = s.v[follower];
    const double bV2 = bV * bV, bvT = bV * rTOther;
    double dS = bV2 / (aMin * -2.0) + bvT;
    double dR = bV2 / (-10.0 * -2.0) + bvT;
    dS = dS + ev2 / (aMax + aMax);
    dR = dR + ev2 / (-1.0 + -1.0);
    dS = MCS_MAXSD(0.0, dS);
    dR = MCS_MAXSD(0.0, dR);
    failS = failS || dS > -bGap; failR = failR || dR > -bGap;
  }
  return (failS ? 0 : 1) | (failR ? 0 : 2);
}

// laneChangeProbability 0xe19f0
__device__ __forceinline__ void lane_change_probability(double p[3], double q0, double q1, double q2, bool leftOk, bool rightOk) {
  const double wP = MCS_MAXSD(p[2], p[0]) * 1.5 + 0.5;
  const double wQ = MCS_MAXSD(q2, q0) * 1.5 + 0.5;
  double L = 0.0, R = 0.0;
  if (leftOk) L = p[0] * wP + q0 * wQ;
  const double K = wP * p[1] + wQ * q1;
  if (rightOk) R = p[2] * wP + q2 * wQ;
  const double S = (L + K) + R;
  p[0] = mcs_div(L, S); p[1] = K / S; p[2] = mcs_div(R, S);
}

// lateral centring 0x10441b-0x10447f
__device__ __forceinline__ double centering_vy(double y, double centerY) {
  const double dy = y - centerY;
  const double ady = fabs(dy);
  double m = 0.1;
  if (!(0.1 > ady)) m = MCS_MINSD(1.3, ady);
  return ((-dy) / (ady + 1e-10)) * m;
}

// RSS emergency test 0x104988-0x1049f9
__device__ __forceinline__ bool rss_emergency(const Smem& s, int front, double ex, double ev, double el, double rTEgo,
                                              double aMin, double aMax) {
  const double gap = (s.x[front] - ex) - (s.l[front] + el) * 0.5;
  const double fv = s.v[front];
  double d = (ev * ev) / (aMin * -2.0) + ev * rTEgo;
  d = d + (fv * fv) / (aMax + aMax);
  d = MCS_MAXSD(0.0, d);
  return gap < d;
}

// MOBIL flag bits (Smem::mflag)
enum { MOBIL_PARTS = 7 };
enum { MF_STRICT_L = 1, MF_RELAX_L = 2, MF_STRICT_R = 4, MF_RELAX_R = 8, MF_CAN_L = 16, MF_CAN_R = 32 };

// side neighbours of vehicle at x in lane `ln`: leader = first with x_elem >= x, follower = last with x_elem < x
__device__ __forceinline__ void side_leader_follower(const Smem& s, const int* laneStart, int ln, double x, bool ordered,
                                                     int& leader, int& follower) {
  const int lo = laneStart[ln], n = laneStart[ln + 1] - lo;
  const int f = lane_lower(s, lo, n, x);
  leader = f < n ? s.vec[lo + f] : -1;
  // `ordered`: no lane vector is stale-ordered this step, so the reference's second (reverse) search ends right below f
  const int r = ordered ? f : lane_reverse(s, lo, n, x);
  follower = r > 0 ? s.vec[lo + r - 1] : -1;
}

// The part of laneChangeBehaviorMOBIL (0xf44c9-0xf4dde, 0xf5259-0xf584e) that does not depend on the vehicle's own
// IDM action: accelerations in the neighbour lanes, follower deltas, RSS gates.  A vehicle's evaluation is seven
// independent sub-tasks of similar cost (part 0: old follower delta; 1-3: left lane own IDM / RSS gates / follower
// delta; 4-6: the same for the right lane) executed by whichever threads picked them from the queue; results go to
// shared memory.
// Everything mobil_decide needs from the seven sub-tasks of one vehicle
struct MobilGeo { double accNewL, accNewR, dNewL, dNewR, deltaOld; int mf; };

// the side lane of vehicle k a MOBIL evaluation looks at (side 0 = left, 1 = right): is there a main lane, and who would be
// k's leader / follower on it — ONE search per vehicle and side, shared by the side's three sub-tasks
__device__ __forceinline__ bool mobil_side_neighbours(const Smem& s, const SceneDev& sc, const int* laneStart, int k, int side,
                                                      bool ordered, int& leader, int& follower) {
  const LaneDev& L = sc.lanes[s.lane[k]];
  const int sl = side == 0 ? L.leftIdx : L.rightIdx;
  leader = -1; follower = -1;
  const bool can = (L.nb & (side == 0 ? LANE_NB_LEFT_MAIN : LANE_NB_RIGHT_MAIN)) != 0;
  if (can) side_leader_follower(s, laneStart, sl, s.x[k], ordered, leader, follower);
  return can;
}

// one sub-task, given the side's neighbours (part 0: can = true, no neighbours).  kStore: the result goes to the vehicle's
// cell of the shared arrays (the task queue of rollout_kernel); otherwise it comes back in `value` (parts 0, 1, 3, 4, 6) or
// `mfOut` (parts 2, 5: the side's flag bits)
template <bool kStore>
__device__ __forceinline__ void mobil_sub_task(const Smem& s, const SceneDev& sc, int np, int k, int front, int back, int part,
                                               bool can, int leader, int follower, double& value, int& mfOut) {
  const LaneDev& L = sc.lanes[s.lane[k]];
  // Which IDM evaluation is this sub-task?  Parts 0, 3, 6 are all "what does vehicle j lose when k drives between it
  // and F" (j = old follower / would-be follower on the left / right); parts 1, 3 are k's own acceleration behind the
  // side lane's leader plus the two RSS gates.  Branches only select the operands; each formula is instantiated once.
  const int side = part > 3 ? 1 : 0, sub = part == 0 ? 2 : (part - 1) - 3 * side;   // sub 0: own IDM, 1: RSS gates, 2: follower delta
  if (sub == 0) {
    double accNew = 0.0;
    if (can) {
      const int sl = side == 0 ? L.leftIdx : L.rightIdx;
      const double ev = s.v[k];
      const double vdCap = MCS_MINSD(s.vd[k], 1.1 * sc.lanes[sl].vLimit);
      accNew = idm_acc(s, load_idm(s, np, k), leader, -1, follower, s.x[k], ev, s.l[k], mcs_pow4(ev / vdCap));
    }
    if (kStore) s.mob[side * np + k] = accNew; else value = accNew;
    return;
  }
  if (sub == 1) {
    int mf = 0;
    if (can) {
      const double ex = s.x[k], ev = s.v[k], el = s.l[k];
      const double rTOther = sc.f[(size_t)F_RSS_RTOTHER * np + k], rTEgo = sc.f[(size_t)F_RSS_RTEGO * np + k];
      const double aMin = sc.f[(size_t)F_RSS_AMIN * np + k], aMax = sc.f[(size_t)F_RSS_AMAX * np + k];
      mf = side == 0 ? MF_CAN_L : MF_CAN_R;
      // strict (Agent+0x118) and relaxed (Agent+0x148: aMin -10, aMax -1) RSS; MF_RELAX_* = MF_STRICT_* << 1
      mf |= rss_safe_both(s, leader, follower, ex, ev, el, rTOther, rTEgo, aMin, aMax) * (side == 0 ? MF_STRICT_L : MF_STRICT_R);
    }
    if (kStore) s.mflag[side * np + k] = (uint8_t)mf; else mfOut = mf;
    return;
  }
  const int j = part == 0 ? back : follower;
  const int F = part == 0 ? front : leader;
  double delta = 0.0;
  if (can && j >= 0) {
    // idmAccFrontBack for j with fronts {F, k} and with {F}: the two calls share every term
    const IdmP pj = load_idm(s, np, j);
    const double xj = s.x[j], vj = s.v[j], lj = s.l[j];
    const double freeAcc = (1.0 - mcs_pow4(vj / s.vd[j])) * pj.accMax;
    const double na = -pj.accMax;
    const double dDes = (vj * pj.thw + pj.minDis) * 0.7;
    double termF = 0.0, termK = 0.0;
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      const int f = q == 0 ? F : k;
      double t = -10000000000.0;
      if (f >= 0) {
        const double gap = (s.x[f] - xj) - (lj + s.l[f]) * 0.5;
        if (gap > 0.0) {
          const double dv = (vj - s.v[f]) * vj;
          t = dv / pj.sq2 + dDes;
          t = MCS_MAXSD(t, 0.0);
          t = mcs_div(t, gap);
          t = (t * t) * na;
        }
      }
      if (q == 0) termF = t; else termK = t;
    }
    const double dWith = F >= 0 ? MCS_MINSD(termK, termF) : termK;
    const double sumWithout = F >= 0 ? termF + 0.0 : 0.0;
    double rw = MCS_MINSD((dWith + 0.0) + freeAcc, pj.accMax);
    rw = MCS_MAXSD(pj.accMin, rw);
    double ro = MCS_MINSD(sumWithout + freeAcc, pj.accMax);
    ro = MCS_MAXSD(pj.accMin, ro);
    delta = part == 0 ? ro - rw : rw - ro;       // old follower: without - with (0xf4583); new follower: with - without
  }
  if (kStore) s.mob[(part == 0 ? 4 : 2 + side) * np + k] = delta; else value = delta;
}

// A task of rollout_kernel's queue: sub-tasks [first, last) of one side lane (`side` 0 / 1: the vehicle's own acceleration
// there, the two RSS gates, the new follower's loss — sub-tasks 1-3 / 4-6 of laneChangeBehaviorMOBIL behind ONE search for
// the side's leader / follower), or with side < 0 the old follower's loss (sub-task 0; first = 2, last = 3).
__device__ __forceinline__ void mobil_task(const Smem& s, const SceneDev& sc, const int* laneStart, int np, int k, int front,
                                           int back, int side, int first, int last, bool ordered) {
  int leader = -1, follower = -1;
  const bool can = side < 0 ? true : mobil_side_neighbours(s, sc, laneStart, k, side, ordered, leader, follower);
  const int base = side < 0 ? -2 : 1 + 3 * side;     // part = base + sub
#pragma unroll 1
  for (int sub = first; sub < last; ++sub) {
    double value;
    int mf;
    mobil_sub_task<true>(s, sc, np, k, front, back, base + sub, can, leader, follower, value, mf);
  }
}
__device__ __forceinline__ MobilGeo load_mobil_geo(const Smem& s, int np, int k) {
  MobilGeo g;
  g.accNewL = s.mob[k]; g.accNewR = s.mob[np + k]; g.dNewL = s.mob[2 * np + k]; g.dNewR = s.mob[3 * np + k];
  g.deltaOld = s.mob[4 * np + k];
  g.mf = s.mflag[k] | s.mflag[np + k];
  return g;
}
// all seven sub-tasks in one thread (bundle_kernel: the lanes of a warp are the same vehicle in different rollouts)
__device__ __forceinline__ MobilGeo mobil_geometry_all(const Smem& s, const SceneDev& sc, const int* laneStart, int np, int k,
                                                       int front, int back) {
  MobilGeo g;
  g.accNewL = g.accNewR = g.dNewL = g.dNewR = g.deltaOld = 0.0; g.mf = 0;
  int leader = -1, follower = -1;
  bool can = true;
#pragma unroll 1
  for (int part = 0; part < MOBIL_PARTS; ++part) {
    double value = 0.0;
    int mf = 0;
    if (part == 1 || part == 4) can = mobil_side_neighbours(s, sc, laneStart, k, part == 4 ? 1 : 0, false, leader, follower);
    mobil_sub_task<false>(s, sc, np, k, front, back, part, can, leader, follower, value, mf);
    if (part == 0) g.deltaOld = value;
    else if (part == 1) g.accNewL = value;
    else if (part == 3) g.dNewL = value;
    else if (part == 4) g.accNewR = value;
    else if (part == 6) g.dNewR = value;
    else g.mf |= mf;
  }
  return g;
}

// Agent fields laneChangeBehaviorMOBIL writes back (commitToDecision 0x90, commitKeepLaneStep 0xb0, targetLaneId 0x68)
struct MobilState { int commit, keepStep, targetLane; };

// laneChangeBehaviorMOBIL 0xf42d0, decision part (0xf4f12-0xf5195), after mobil_geometry has filled s.mob / s.mflag for
// vehicle k and the keep-lane counter test (0xf434c) has passed.  `u` = the uniform draw of 0xf510f (consumed only when
// the vehicle does not continue a committed change).  mode: 0 = inside a rollout at the vehicle's first step (lateral-offset
// prior blended in, 0xf58a0), 1 = inside a rollout later, 2 = Python-facing with requireRssSafety (0xf5858), 3 = without (0xf5053).
__device__ __forceinline__ void mobil_decide(const MobilGeo& geo, const SceneDev& sc, int np, int k, int lane, double v, double el,
                                             double aMin, double pf, double dA, double bA, int mode, bool aggressive,
                                             const double* __restrict__ priors /* F_PRIOR_L row, stride np */, double u, MobilState& ms,
                                             double& ax, double& avy) {
  const int mf = geo.mf;
  const bool canL = mf & MF_CAN_L, canR = mf & MF_CAN_R;
  const bool strictL = mf & MF_STRICT_L, relaxL = mf & MF_RELAX_L, strictR = mf & MF_STRICT_R, relaxR = mf & MF_RELAX_R;
  const double accNewL = geo.accNewL, accNewR = geo.accNewR, dNewL = geo.dNewL, dNewR = geo.dNewR;
  const double deltaOld = geo.deltaOld;
  const double axIdm = ax, vyIdm = avy;
  double axLeft = axIdm, vyLeft = vyIdm, axRight = 0.0, vyRight = 0.0;
  double incL = -10000000000.0, incR = -10000000000.0;
  if (canL) {
    axLeft = (axIdm == aMin) ? axIdm : accNewL;
    vyLeft = MCS_MINSD(1.0, 0.17 * v);
    incL = (((dNewL + deltaOld) * pf + (accNewL - axIdm)) - dA) - bA;
  }
  if (canR) {
    axRight = accNewR;
    vyRight = MCS_MAXSD(-1.0, -0.17 * v);
    incR = (((dNewR + deltaOld) * pf + (accNewR - axIdm)) - dA) + bA;
  }
  const LaneDev& L = sc.lanes[lane];
  if (ms.commit == MCS_COMMIT_LEFT && ms.targetLane != L.id && relaxL) { ax = axLeft; avy = vyLeft; return; }
  if (ms.commit == MCS_COMMIT_RIGHT && ms.targetLane != L.id && relaxR) { ax = axRight; avy = vyRight; return; }
  ms.keepStep = 0;
  const double eL = mcs_exp(incL), eR = mcs_exp(incR);
  const double denom = (1.0 + eL) + eR;
  double p[3];
  p[0] = mcs_div(eL, denom); p[2] = mcs_div(eR, denom); p[1] = (1.0 - p[0]) - p[2];
  // which gates and which prior laneChangeProbability sees (0xf58a0-0xf595a, 0xf5858, 0xf5053); one call below
  double q0 = 0.0, q1 = 0.0, q2 = 0.0;
  bool okL, okR;
  if (mode <= 1) {
    if (ms.commit == MCS_COMMIT_NOT_DECIDED && mode == 0) {
      // lateral-offset prior of the first lane match (computed on the host, scene.cpp).  A vehicle first matched later
      // never uses its prior: this branch needs a history of one state.
      q0 = priors[k]; q1 = priors[np + k]; q2 = priors[2 * np + k];
      if (aggressive) { okL = relaxL; okR = relaxR; }
      else {
        okR = strictR ? true : ((q2 >= 0.5) && relaxR);
        okL = strictL ? true : ((q0 >= 0.5) && relaxL);
      }
    } else if (aggressive) { okL = relaxL; okR = relaxR; }
    else { okL = strictL; okR = strictR; }
  } else if (mode == 2) { okL = strictL; okR = strictR; }
  else { okL = relaxL; okR = relaxR; }
  lane_change_probability(p, q0, q1, q2, okL, okR);
  double pL = p[0];
  if (el > 6.0) {
    pL = pL * 0.5;
    if (L.leftIdx >= 0 && sc.lanes[L.leftIdx].leftmost) pL = 0.0;
    const double k2 = p[1] + p[1], r2 = p[2] + p[2];
    const double S = (k2 + pL) + r2;
    p[1] = k2 / S; p[2] = mcs_div(r2, S);
  }
  if (pL > u && canL) { ms.commit = MCS_COMMIT_LEFT; ms.targetLane = sc.lanes[L.leftIdx].id; ax = axLeft; avy = vyLeft; }
  else if (pL + p[2] > u && canR) { ms.commit = MCS_COMMIT_RIGHT; ms.targetLane = sc.lanes[L.rightIdx].id; ax = axRight; avy = vyRight; }
  else { ms.commit = MCS_COMMIT_KEEP_LANE; ms.targetLane = -1; }
}

// customizedLaneChangeBehavior 0xf5e30 for vehicle k (avy enters as the lane-centring speed 0xf61af-0xf61fb)
__device__ __forceinline__ void customized_lane_change(const Smem& s, const SceneDev& sc, const int* laneStart, int np, int k,
                                                       int lane, int front, int cmd, double x, double v, double el, double evd,
                                                       double rTOther, double rTEgo, double aMin, double aMax, double& ax,
                                                       double& avy) {
  const IdmP pe = load_idm(s, np, k);
  const LaneDev& L = sc.lanes[lane];
  const double vdCap = MCS_MINSD(evd, 1.1 * L.vLimit);
  if (cmd == MCS_ACT_KEEP_LANE || cmd == MCS_ACT_ACC || cmd == MCS_ACT_DCC) {
    IdmP p = pe;
    double vdd = vdCap;
    if (cmd == MCS_ACT_DCC) { p.thw = pe.thw + pe.thw; vdd = 0.8 * vdCap; }
    if (cmd == MCS_ACT_ACC) { p.thw = 0.9; vdd = 1.1 * vdCap; }
    ax = idm_acc(s, p, front, -1, -1, x, v, el, mcs_pow4(v / vdd));
    if (front >= 0 && rss_emergency(s, front, x, v, el, rTEgo, aMin, aMax)) ax = aMin;
    return;
  }
  // for left/right the binary's emergency select reads an uninitialised stack slot (0xf6263); only the aMin outcome is observable
  const bool emergency = front >= 0 && rss_emergency(s, front, x, v, el, rTEgo, aMin, aMax);
  const int sl = cmd == MCS_ACT_LEFT ? L.leftIdx : L.rightIdx;
  if (sl >= 0 && sc.lanes[sl].type == MCS_LANE_MAIN) {
    int leader, follower;
    side_leader_follower(s, laneStart, sl, x, false, leader, follower);
    ax = idm_acc(s, pe, leader, front, follower, x, v, el, mcs_pow4(v / vdCap));
    if (emergency) ax = aMin;
    if (rss_safe(s, leader, follower, x, v, el, rTOther, rTEgo, aMin, aMax)) {
      const double m = MCS_MINSD(1.0, 0.17 * v);
      avy = m * (cmd == MCS_ACT_LEFT ? 1.0 : -1.0);
    }
  } else {
    ax = emergency ? aMin : 0.0;
  }
}

// Barriers of one rollout.  A 256-vehicle rollout owns its thread block; smaller ones share a 256-thread block
// (256 / NPS rollouts side by side, each on its own named barrier and its own slice of shared memory), so that the
// warps of several rollouts fetch the same instructions at about the same time: the small-block instantiations are
// instruction-fetch bound (ncu r01g: 35-48 % `no instruction` stalls with 8 independent 64-thread blocks per SM).
template <int NPS, bool kWholeBlock = (NPS == 256)>
__device__ __forceinline__ void sub_sync(int sub) {
  if (kWholeBlock) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "n"(NPS) : "memory");
}
template <int NPS, bool kWholeBlock = (NPS == 256)>
__device__ __forceinline__ int sub_sync_or(int sub, int pred) {
  if (kWholeBlock) return __syncthreads_or(pred);
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(pred), "r"(sub + 1), "n"(NPS)
      : "memory");
  return r;
}

// Barrier of a rollout's working threads when its last warp is the ego's helper (kHelper: NPS - 32 threads, ids behind the
// G full-rollout barriers); the plain rollout barrier otherwise.
template <int NPS, bool kWholeBlock, bool kHelper, int G>
__device__ __forceinline__ void work_sync(int sub) {
  if (!kHelper) sub_sync<NPS, kWholeBlock>(sub);
  else asm volatile("bar.sync %0, %1;" ::"r"(G + 1 + sub), "n"(NPS - 32) : "memory");
}

// exclusive scan over the NPS threads of one rollout of small per-thread counts (< 2^BITS): one warp vote per bit.
// `flag` is OR-ed over the rollout and comes back in bits 16.. of `total` (nonzero = set somewhere).
template <int NPS, bool kWholeBlock, int BITS, bool kHelper = false, int G = 1>
__device__ __forceinline__ int block_exclusive_scan_small(int v, bool flag, int* scratch, int& total, int k, int sub) {
  constexpr int NW = NPS / 32 - (kHelper ? 1 : 0);
  const int lane = k & 31, w = k >> 5;
  const unsigned below = (1u << lane) - 1u;
  int excl = 0, wtot = 0;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const unsigned m = __ballot_sync(0xffffffffu, (v >> b) & 1);
    excl += __popc(m & below) << b;
    wtot += __popc(m) << b;
  }
  if (__any_sync(0xffffffffu, flag)) wtot |= 1 << 16;
  if (lane == 0) scratch[w] = wtot;
  work_sync<NPS, kWholeBlock, kHelper, G>(sub);
  // the warp totals come in with two 16-byte loads (the scratch rows are 16-byte aligned and zero beyond the rollout's
  // warps); the flags add up in bits 16.. (at most eight of them), the counts below them
  int c[8];
  {
    const int4 a = *reinterpret_cast<const int4*>(scratch);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w;
    if (NW > 4) {
      const int4 b = *reinterpret_cast<const int4*>(scratch + 4);
      c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
      c[4] = c[5] = c[6] = c[7] = 0;
    }
  }
  int base = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < NW; ++q) {
    base += q < w ? c[q] : 0;
    tot += c[q];
  }
  total = tot;
  return (base & 0xffff) + excl;
}

}  // namespace mcs
#include "merge_behaviors.cuh"
namespace mcs {

// kMerge: the scene has a merging/exit lane or vehicle, or a candidate names a gap — yielding intentions and the gap
// behaviours are compiled in.  The plain lane-change instantiation carries none of their registers.
// kHelper: the rollout's last warp carries no vehicles (the host's thread -> slot table leaves it empty) and runs the ego's
// gap execution — customizedMergingBehavior, the longest dependent chain of a merging / exit step, which reads the
// pre-step state only and draws nothing — next to the other warps' whole plan phase instead of inside it.
// kHelper 2 (scenes with ramp vehicles): lanes 1 .. 31 of the helper warp also run mergingBehaviorClosestGap of the first 31
// ramp vehicles — the same kind of chain — and the vehicles pick their actions up like the ego does.
template <bool kTraj, bool kMerge, int NPS, int BT, int kHelper = 0>
// resident 256-thread-equivalents per SM the register allocation aims at: 3 (80 registers) for the lane-change path —
// the kernel is latency-bound, the third block hides more than its few spills cost (measured 6.26e9 -> 7.06e9
// vehicle-steps/s) — and 2 (128 registers) for the much larger merging instantiation
#ifndef MCS_ALIGN_STEPS
#define MCS_ALIGN_STEPS 1   // rollouts that share a block re-align every this many steps
#endif
#ifndef MCS_MIN_BLOCKS
#define MCS_MIN_BLOCKS 3
#endif
#ifndef MCS_MIN_BLOCKS_MERGE
#define MCS_MIN_BLOCKS_MERGE 2
#endif
// (a block of more than 512 threads is alone on its SM and gets what the register file allows: 896 threads -> 72)
// A merging rollout that fills a 256-thread block by itself (129-256 vehicles: the exit scenes of configs[3]) comes in
// launches of many waves, where the third resident block pays as it does for the lane-change path (80 registers, 480 B
// of spills): 4 gaps x 4096 rollouts x 200 vehicles x 50 steps 28.7 -> 27.1 ms, x 512 rollouts 3.67 -> 3.37 ms.
#ifndef MCS_MIN_BLOCKS_MERGE_FULL
#define MCS_MIN_BLOCKS_MERGE_FULL 3
#endif
#define MCS_RESIDENT_BLOCKS_OF(BT, NPS, kMerge) (((kMerge) && (NPS) == 256 && (BT) == 256) ? MCS_MIN_BLOCKS_MERGE_FULL : ((kMerge) && (BT) == 448) ? 2 : MCS_RESIDENT_BLOCKS(BT, kMerge))
#define MCS_RESIDENT_BLOCKS(BT, kMerge) (((kMerge) ? MCS_MIN_BLOCKS_MERGE : MCS_MIN_BLOCKS) * 256 / (BT) > 0 ? ((kMerge) ? MCS_MIN_BLOCKS_MERGE : MCS_MIN_BLOCKS) * 256 / (BT) : 1)
__global__ void __launch_bounds__(BT, MCS_RESIDENT_BLOCKS_OF(BT, NPS, kMerge)) rollout_kernel(const SceneDev sc, const CandDev* __restrict__ cands, const LaunchParams lp, const int region_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // NPS = sc.n_pad = blockDim.x as a compile-time constant: every shared-memory array address and row stride folds into
  // the instruction's immediate offset instead of living in registers
  constexpr int np = NPS;
  constexpr int G = BT / NPS;                          // rollouts per thread block
  const int sub = threadIdx.x / NPS;                   // which of them this warp belongs to
  const int tid = threadIdx.x - sub * NPS;             // thread within the rollout: warp votes, task loops
  static_assert(!kHelper || (kMerge && NPS >= 64 && 2 * G + 1 <= 15), "helper warp: merging instantiation, a spare warp, barrier ids");
  const bool helper = kHelper && tid >= NPS - 32;      // the ego's helper warp
  constexpr bool kRamp = kHelper == 2;                 // ... which also serves the ramp vehicles
  constexpr int NPW = kHelper ? NPS - 32 : NPS;        // working threads of a rollout
  // Vehicle slot of this thread.  Slots keep the reference's visit order and every shared array is indexed by slot; the
  // threads follow the slots in increasing order, so scans over threads are scans over slots, but when a rollout has
  // spare warps the host leaves gaps so that vehicles that start on different lanes — i.e. run different behaviour code
  // (merging / yielding / plain) — sit in different warps and their code runs side by side instead of in series.
  const int n = sc.n, k = sc.thread_slot[tid];
  const int cand = blockIdx.y;
  const int64_t ridx = (int64_t)blockIdx.x * G + sub;  // rollout within this launch
  const CandDev cd = cands[cand];
  mcs_status* st_out = lp.status + ((size_t)cand * lp.count + ridx);
  if (ridx >= lp.count || !cd.ok) {
    // tail of the grid, or setEgoAgentIdAndDrivingBehavior failed (the reference runs nothing, runMTimes 0xc2935): no
    // rollout here — but the sub-group keeps answering the block's per-step alignment barrier until the others are done
    if (ridx < lp.count && k == 0) { mcs_status z; memset(&z, 0, sizeof z); *st_out = z; }
    if (G > 1) while (__syncthreads_or(0)) {}
    return;
  }

  Smem s;
  carve_smem(s, smem_raw + (G > 1 ? sub * region_bytes : 0), np, lp.steps, kRamp);
  int* laneStart = s.cnt;  // [n_lanes + 1]; flips between s.cnt and s.cnt + 16 at every rebuild

  const bool live = k < n;
  const uint64_t key = mcs_rollout_key(lp.seed + (uint64_t)cand, (uint64_t)(lp.first + ridx));

  // ---- per-vehicle registers
  const double* F = sc.f;
  const int32_t* I = sc.i;
  double x = 0, y = 0, v = 0, vy = 0;
  int lane = -1, beh = MCS_BEH_OTHER, commit = 0, keepStep = 0, targetLane = -1, inputIdx = 0;
  bool aggressive = false;
  const bool isEgo = live && k == cd.ego_slot;
  if (live) {
    x = F[(size_t)F_X * np + k]; y = F[(size_t)F_Y * np + k]; v = F[(size_t)F_V * np + k]; vy = F[(size_t)F_VY * np + k];
    lane = I[(size_t)I_LANE * np + k]; beh = I[(size_t)I_BEH * np + k]; commit = I[(size_t)I_COMMIT * np + k];
    keepStep = I[(size_t)I_KEEPSTEP * np + k]; targetLane = I[(size_t)I_TARGETLANE * np + k];
    inputIdx = I[(size_t)I_INPUT * np + k];
    // Agent ctor draw (0xf89e6): draw index = position in the INPUT list
    aggressive = (mcs_unit31(mcs_rand31(key, (uint64_t)inputIdx))) > 0.8;
    if (isEgo) targetLane = cd.target_lane_id;
  }
  // stage what neighbours gather
  // static per-vehicle values live in shared memory only (registers are what limits the resident blocks per SM)
  s.x[k] = x; s.v[k] = v;
  s.l[k] = live ? F[(size_t)F_L * np + k] : 0.0; s.vd[k] = live ? F[(size_t)F_VD * np + k] : 1.0;
  s.rss[k] = live ? F[(size_t)F_RSS_RTEGO * np + k] : 0.0;
  s.rss[np + k] = live ? F[(size_t)F_RSS_AMIN * np + k] : -1.0;
  s.rss[2 * np + k] = live ? F[(size_t)F_RSS_AMAX * np + k] : -1.0;
  s.idm[k] = live ? F[(size_t)F_IDM_ACCMAX * np + k] : 1.0;
  s.idm[np + k] = live ? F[(size_t)F_IDM_MINDIS * np + k] : 0.0;
  s.idm[2 * np + k] = live ? F[(size_t)F_IDM_ACCMIN * np + k] : 0.0;
  s.idm[3 * np + k] = live ? F[(size_t)F_IDM_ACCCOM * np + k] : -1.0;
  s.idm[4 * np + k] = live ? F[(size_t)F_IDM_THW * np + k] : 0.0;
  { const double sq = sqrt((-s.idm[k]) * s.idm[3 * np + k]); s.idm[5 * np + k] = sq + sq; }   // idmAccFrontBack 0xebd2b, hoisted
  s.lane[k] = (uint8_t)(lane < 0 ? 0xff : lane);
  s.vec[k] = sc.lane_vec[k];
  int myPos = 0;                // its position in s.vec
  if (k <= sc.n_lanes) laneStart[k] = sc.lane_start[k];
  if (k < 8) s.flags[k] = 0;
  if (tid < 32) s.scan[tid] = 0;       // warp totals of the scans: zero beyond the rollout's working warps
  if (kRamp) { s.hy[k] = y; s.hbeh[k] = (uint8_t)(live ? beh : 0xff); s.hok[k] = 0xff; }
  sub_sync<NPS, G == 1>(sub);
  if (k < laneStart[sc.n_lanes]) s.pos[s.vec[k]] = (uint8_t)k;
  // The helper warp's lanes 1 .. 31 serve the first 31 ramp vehicles (behaviour `merging`, not the ego) in slot order:
  // mergingBehaviorClosestGap, like the ego's gap execution, reads the pre-step state only and draws nothing.  A lane
  // keeps its vehicle for the whole rollout (it idles once the vehicle has reached the main lane).
  int hSlot = -1;
  if (kRamp && helper) {
    const int hl = tid - (NPS - 32);
    int seen = 0;
    for (int base = 0; base < n; base += 32) {
      const int j = base + hl;
      const unsigned m = __ballot_sync(0xffffffffu, j < n && j != cd.ego_slot && s.hbeh[j] == MCS_BEH_MERGING);
      const int want = hl - 1 - seen;                    // which set bit of this chunk is mine
      if (hl >= 1 && hSlot < 0 && want >= 0 && want < __popc(m)) {
        unsigned mm = m;
        for (int q = 0; q < want; ++q) mm &= mm - 1;
        hSlot = base + __ffs(mm) - 1;
      }
      seen += __popc(m);
    }
    if (hSlot >= 0) s.hok[hSlot] = 1;
  }
  if (isEgo) {
    if (kHelper == 1) s.ego[0] = y;
    s.flags[2] = beh;                   // the ego's current behaviour: `exit` vehicles are yielding candidates (0x104c90)
    s.flags[4] = cd.gap_slot;           // Agent::targetGapId as a slot; < 0 = last gap
    s.flags[5] = lp.steps;              // < lp.steps once the ego has reached its target lane (StatusOneSim::finish)
  }
  sub_sync<NPS, G == 1>(sub);
  myPos = s.pos[k];
  const bool served = kRamp && live && s.hok[k] != 0xff;     // this vehicle's closest-gap plan comes from the helper warp

  // statistics (Simulation::run 0x10ee71-0x10f49a) accumulated on the fly in history order
  // statistics (Simulation::run 0x10ee71-0x10f49a): per-object sums accumulated on the fly in history order; the ego's own
  // history (acceleration, speed, thw ratio) is kept in shared memory and reduced in the same order after the loop
  // states recorded so far = step + 1 (every executed step appends one)
  {
    const double a0 = live ? F[(size_t)F_A * np + k] : 0.0;
    s.acc[k] = fabs(a0); s.acc[np + k] = v;
    if (isEgo) { s.chist[0] = a0; s.vhist[0] = v; }
  }
  if (kTraj && live) {
    double* t = lp.traj + ((((size_t)cand * lp.count + ridx) * n + inputIdx) * (size_t)(lp.steps + 1)) * 4;
    t[0] = x; t[1] = y; t[2] = v; t[3] = F[(size_t)F_A * np + k];
  }
  uint32_t drawCtr = (uint32_t)n;   // draws 0..n-1 were the ctor flags
  int step = 0;
  // scalars only the ego thread touches (arrival, gap slot, finish step) and the two sticky error marks live in s.flags:
  // registers are what limits the resident rollouts per SM
  YieldLocal yt;
  yt.init();

  bool running = true;
  int window = 0;   // steps left until the block's next alignment point
  int hGap = cd.gap_slot;   // helper warp: Agent::targetGapId as a slot (the helper re-targets it, below)
  while (true) {
    if (G > 1) {
      // the rollouts of a block start every MCS_ALIGN_STEPS-th step together, so their warps run through the same code at
      // about the same time and share its instruction fetches; a finished rollout idles here until the last one is done
      if (MCS_ALIGN_STEPS == 1 || window == 0) {
        if (!__syncthreads_or((running && lp.steps > step) ? 1 : 0)) break;
        if (!(running && lp.steps > step)) continue;
        window = MCS_ALIGN_STEPS;
      }
      if (MCS_ALIGN_STEPS > 1) {
        if (!(running && lp.steps > step)) { window = 0; continue; }
        --window;
      }
    } else if (!(lp.steps > step)) {
      break;
    }
    if (kHelper && helper) {
      // ---- the ego's helper warp: customizedMergingBehavior on the pre-step state, side by side with the plan of everybody else
      if (!kRamp && tid == NPS - 32) {
        const int eg = cd.ego_slot;
        const int elane = s.lane[eg], ebeh = s.flags[2];
        if (elane != 0xff && (ebeh == MCS_BEH_MERGING || ebeh == MCS_BEH_EXIT) && s.flags[0] == 0) {
          const int lo = laneStart[elane], cnt = laneStart[elane + 1] - lo;
          MergeSelf me;
          me.k = eg; me.lane = elane; me.x = s.x[eg]; me.y = s.ego[0]; me.v = s.v[eg]; me.l = s.l[eg];
          const int fpos = lane_upper(s, lo, cnt, me.x);          // frontAgent 0xe2670, as the reference searches it
          me.front = fpos < cnt ? (int)s.vec[lo + fpos] : -1;
          me.w = F[(size_t)F_W * np + eg];
          me.rss.rTOther = F[(size_t)F_RSS_RTOTHER * np + eg]; me.rss.rTEgo = s.rss[eg]; me.rss.aMin = s.rss[np + eg];
          me.rss.aMax = s.rss[2 * np + eg]; me.rss.dSoft = F[(size_t)F_RSS_DSOFT * np + eg];
          me.idm = load_idm(s, np, eg);
          double hax = 0.0, havy = 0.0;
          const bool okSide = merging_behavior(s, sc, laneStart, me, s.vd[eg], true, false, cd.action == MCS_ACT_LEFT, hGap, hax, havy);
          s.ego[1] = hax; s.ego[2] = havy; s.ego[3] = okSide ? 1.0 : 0.0;
        }
      }
      // ---- the helper warp: the gap behaviours on the pre-step state, side by side with the plan of everybody else — lane 0
      // the ego's customizedMergingBehavior, lanes 1 .. 31 the ramp vehicles' mergingBehaviorClosestGap
      if (kRamp) {
        const int eg = cd.ego_slot;
        int slot = -1;
        bool fixedGap = false;
        if (tid == NPS - 32) {
          const int ebeh = s.flags[2];
          if (s.lane[eg] != 0xff && (ebeh == MCS_BEH_MERGING || ebeh == MCS_BEH_EXIT) && s.flags[0] == 0) { slot = eg; fixedGap = true; }
        } else if (hSlot >= 0 && s.hbeh[hSlot] == MCS_BEH_MERGING && s.lane[hSlot] != 0xff) {
          slot = hSlot;
        }
        if (slot >= 0) {
          const int vlane = s.lane[slot];
          const int lo = laneStart[vlane], cnt = laneStart[vlane + 1] - lo;
          MergeSelf me;
          me.k = slot; me.lane = vlane; me.x = s.x[slot]; me.y = s.hy[slot]; me.v = s.v[slot]; me.l = s.l[slot];
          const int fpos = lane_upper(s, lo, cnt, me.x);          // frontAgent 0xe2670, as the reference searches it
          me.front = fpos < cnt ? (int)s.vec[lo + fpos] : -1;
          me.w = F[(size_t)F_W * np + slot];
          me.rss.rTOther = F[(size_t)F_RSS_RTOTHER * np + slot]; me.rss.rTEgo = s.rss[slot]; me.rss.aMin = s.rss[np + slot];
          me.rss.aMax = s.rss[2 * np + slot]; me.rss.dSoft = F[(size_t)F_RSS_DSOFT * np + slot];
          me.idm = load_idm(s, np, slot);
          double hax = 0.0, havy = 0.0;
          // two call sites: each is the behaviour specialised for its kind of gap (the ego's lane walks the code it walked
          // before the ramp vehicles moved in)
          const bool okSide = fixedGap ? merging_behavior(s, sc, laneStart, me, s.vd[slot], true, false, cd.action == MCS_ACT_LEFT, hGap, hax, havy)
                                       : merging_behavior(s, sc, laneStart, me, s.vd[slot], false, false, true, -1, hax, havy);
          s.hax[slot] = hax; s.havy[slot] = havy; s.hok[slot] = okSide ? 1 : 0;
        }
      }
      sub_sync<NPS, G == 1>(sub);                                   // the plan is done: the ego picks its action up
      const int anyChangedH = sub_sync_or<NPS, G == 1>(sub, 0);     // moved and matched
      const bool egoReachedH = s.flags[0] != 0;
      if (anyChangedH) {
        const int anyInvH = sub_sync_or<NPS, G == 1>(sub, 0);
        if (anyInvH) sub_sync<NPS, G == 1>(sub);
        laneStart = laneStart == s.cnt ? s.cnt + 16 : s.cnt;
        sub_sync<NPS, G == 1>(sub);
      }
      if (tid == NPS - 32 && hGap >= 0) {
        // the gap's rear vehicle left the target lane: re-target to the last vehicle of that lane behind it, else the
        // last gap (matchAgentsOnCorridor 0x10c320-0x10c39d)
        const int gl = s.lane[hGap];
        if (gl == 0xff) s.flags[6] = 1;
        else if (sc.lanes[gl].id != cd.target_lane_id) {
          int tl = -1;
          for (int j = 0; j < sc.n_lanes; ++j) if (sc.lanes[j].id == cd.target_lane_id) tl = j;
          if (tl < 0) s.flags[6] = 1;
          else {
            const int lo = laneStart[tl], cnt = laneStart[tl + 1] - lo;
            const double gx = s.x[hGap];
            int found = -1;
            for (int q = cnt - 1; q >= 0; --q) {
              const int o = s.vec[lo + q];
              if (gx > s.x[o]) { found = o; break; }
            }
            hGap = found;
          }
        }
      }
      step++;
      if (egoReachedH && lp.min_steps <= step - 1) { if (G > 1) { running = false; continue; } else break; }
      continue;
    }
    // ================= plan, part A: own-lane search, IDM, draw counts =================
    double ax = 0.0, avy = 0.0;
    int front = -1, back = -1;
    bool needMobil = false, noise = false, inversion = false, ordered = false;
    int mode = 0;   // 1 idmBehaviorAndYielding, 2 customizedLaneChange (ego), 3 customizedMerging (ego), 4 closest gap
    int nDraw = 0;
    const bool hasLane = live && lane >= 0;
    double centerY = 0.0, vLimit = 0.0;
    int ycand[5];
    int nY = 0;
    if (tid == 0) s.flags[1] = 0;     // the movers' queue of this step's lane match (its last reader was before the step's end)
    if (hasLane) {
      const LaneDev& L = sc.lanes[lane];
      centerY = L.centerY; vLimit = L.vLimit;
      // own-lane leader / follower: while every lane vector is strictly ordered by x (checked pairwise, OR-ed through the
      // scan below) upper_bound / the reverse search of frontAgent / followingAgent land on the neighbours in the vector
      if (myPos > laneStart[lane]) inversion = !(s.x[s.vec[myPos - 1]] < x);
      if (isEgo) {
        const bool lr = cd.action == MCS_ACT_LEFT || cd.action == MCS_ACT_RIGHT;
        const bool reached = s.flags[0] != 0;
        if (beh == MCS_BEH_IDM || beh == MCS_BEH_IDM_LC) mode = (lr && reached) ? 1 : 2;
        else if (beh == MCS_BEH_MERGING || beh == MCS_BEH_EXIT) mode = reached ? 1 : 3;
        else s.flags[6] = 1;
      } else if (beh == MCS_BEH_IDM_LC) {
        mode = 1; noise = true;
        needMobil = !(commit == MCS_COMMIT_KEEP_LANE && keepStep <= 9) && (L.nb & (LANE_NB_LEFT_MAIN | LANE_NB_RIGHT_MAIN)) != 0;
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
          if (L.nb & LANE_NB_RIGHT_MERGING) nY = side_vector(s, laneStart, L.rightIdx, x, ycand);
          if (sc.has_exit_lanes && (L.nb & LANE_NB_LEFT_MAIN)) {
            const int eg = cd.ego_slot;
            if (eg != k && s.lane[eg] == L.leftIdx && s.flags[2] == MCS_BEH_EXIT) ycand[nY++] = eg;
          }
        }
      } else if (mode == 3 || mode == 4) {
        s.flags[6] = 1;         // cannot happen: the host selects kMerge for such scenes
      }
    }
    // compaction: queue MOBIL tasks
    int nTasks;
    {
      int total;
      const int pos = block_exclusive_scan_small<NPS, G == 1, 1, kHelper != 0, G>(needMobil ? 1 : 0, inversion, s.scan, total, tid, sub);
      nTasks = total & 0xffff;
      ordered = (total >> 16) == 0;
      if (hasLane) {
        const int lo = laneStart[lane], cnt = laneStart[lane + 1] - lo;
        if (total >> 16) {          // some lane vector is stale-ordered: search it the way the reference does
          const int fpos = lane_upper(s, lo, cnt, x);
          front = fpos < cnt ? s.vec[lo + fpos] : -1;
          if (needMobil) { const int bpos = lane_reverse(s, lo, cnt, x); back = bpos > 0 ? s.vec[lo + bpos - 1] : -1; }
        } else {
          front = myPos + 1 < lo + cnt ? s.vec[myPos + 1] : -1;
          if (needMobil) back = myPos > lo ? s.vec[myPos - 1] : -1;
        }
      }
      if (needMobil) s.tasks[pos] = k | (front & 0xff) << 8 | (back & 0xff) << 16 | ((front < 0) ? 1 << 24 : 0) | ((back < 0) ? 1 << 25 : 0);
    }
    work_sync<NPS, G == 1, kHelper != 0, G>(sub);
    // The seven sub-tasks of a queued vehicle can run as 7, 5 or 3 tasks: every sub-task alone (each side sub-task behind
    // its own search; shortest chain, most warps), per side [own acceleration + RSS gates] and [new follower's loss], or
    // per side all three behind one search (longest chain, fewest warps).  Groups of like tasks start on warp
    // boundaries.  The kernel is latency-bound per step: a merging / exit rollout with seven or eight warps finishes the
    // seven short tasks soonest (measured, scripts/config_times.py / kernel_time.py: configs[3] 16.3 ms with 7 tasks, 17.7
    // with 5, 18.4 with 3), the 256-vehicle lane-change rollout is as fast with 5 as with 7 at 6 % fewer instructions
    // (configs[4] 1.057 / 1.061 / 1.032e10 vehicle-steps/s, 64.7 -> 60.9 warp instructions per vehicle-step), a rollout with
    // one to four warps wants the few long ones (24 vehicles x 5 x 4096 rollouts: 3.87 / 3.07 / 2.90 ms; configs[2] x 512
    // rollouts on 128 threads: 1.096 / 1.054 / 1.067 ms).
    constexpr int split = NPS >= 256 ? (kMerge ? 7 : 5) : (NPS >= 128 ? 5 : 3);
    const int w1 = (nTasks + 31) >> 5, nSide = 2 * nTasks, w2 = (nSide + 31) >> 5, nT32 = w1 << 5, nSide32 = w2 << 5;
    // part = t / nT32 without the integer-division sequence: nT32 = 32 m with m <= 8, t / 32 < 64, and
    // (u * ceil(2^16 / m)) >> 16 == u / m for u < 64 (error < 64 / 2^16 < 1 / m)
    const unsigned magic = (nT32 == 32 ? 65536u : nT32 == 64 ? 32768u : nT32 == 96 ? 21846u : nT32 == 128 ? 16384u
                            : nT32 == 160 ? 13108u : nT32 == 192 ? 10923u : nT32 == 224 ? 9363u : 8192u);
    const int baseOld = split == 5 ? 2 * nSide32 : nSide32;       // where the old-follower tasks start (splits 5 and 3)
    const int nSlots = split == 7 ? 7 * nT32 : baseOld + nTasks;
    for (int t = tid; t < nSlots; t += NPW) {
      int side = -1, ti = 0, first = 2, last = 3;
      if constexpr (split == 7) {
        const int part = (int)(((unsigned)t >> 5) * magic >> 16);
        ti = t - part * nT32;
        if (ti >= nTasks) continue;
        const int rec = s.tasks[ti];
        const int vk = rec & 0xff;
        const int vf = (rec >> 24 & 1) ? -1 : (rec >> 8 & 0xff);
        const int vb = (rec >> 25 & 1) ? -1 : (rec >> 16 & 0xff);
        int leader = -1, follower = -1, mf;
        double value;
        const bool can = part == 0 ? true : mobil_side_neighbours(s, sc, laneStart, vk, part > 3 ? 1 : 0, ordered, leader, follower);
        mobil_sub_task<true>(s, sc, np, vk, vf, vb, part, can, leader, follower, value, mf);
        continue;
      } else if (t >= baseOld) {
        side = -1; ti = t - baseOld; first = 2; last = 3;
      } else {
        const int second = t >= nSide32 ? 1 : 0;      // split 5: the group of the new followers' losses
        const int q = t - second * nSide32;
        if (q >= nSide) continue;
        side = q >= nTasks ? 1 : 0;
        ti = q - side * nTasks;
        first = second ? 2 : 0;
        last = (split == 5 && !second) ? 2 : 3;
      }
      const int rec = s.tasks[ti];
      const int vk = rec & 0xff;
      const int vf = (rec >> 24 & 1) ? -1 : (rec >> 8 & 0xff);
      const int vb = (rec >> 25 & 1) ? -1 : (rec >> 16 & 0xff);
      mobil_task(s, sc, laneStart, np, vk, vf, vb, side, first, last, ordered);
    }
    work_sync<NPS, G == 1, kHelper != 0, G>(sub);
    // draw counts: [noise][MOBIL decision draw]
    bool mobilDraw = false;
    if (needMobil) {
      const int mf = s.mflag[k] | s.mflag[np + k];
      const int laneId = sc.lanes[lane].id;
      const bool contL = commit == MCS_COMMIT_LEFT && targetLane != laneId && (mf & MF_RELAX_L);
      const bool contR = commit == MCS_COMMIT_RIGHT && targetLane != laneId && (mf & MF_RELAX_R);
      mobilDraw = !(contL || contR);
    }
    nDraw = nY + (noise ? 1 : 0) + (mobilDraw ? 1 : 0);
    int totalDraws;
    // nDraw <= 5 yielding candidates + noise + MOBIL draw
    const uint64_t drawBase = (uint64_t)drawCtr + (uint64_t)block_exclusive_scan_small<NPS, G == 1, kMerge ? 3 : 2, kHelper != 0, G>(nDraw, false, s.scan + 8, totalDraws, tid, sub);
    drawCtr += (uint32_t)totalDraws;

    // ================= plan, part B: actions =================
    if (hasLane) {
      const IdmP pe = load_idm(s, np, k);
      const double el = s.l[k], evd = s.vd[k];
      const double rTEgo = s.rss[k], aMin = s.rss[np + k], aMax = s.rss[2 * np + k];
      if (isEgo) {    // Agent::thwRatioHistory is read for the ego only (0x10f14b)
        double thwPush = 1.0;
        if (front >= 0) thwPush = (((s.x[front] - x) - (s.l[front] + el) * 0.5) / v) / pe.thw;
        s.thist[s.flags[3]++] = thwPush;
      }
      avy = centering_vy(y, centerY);
      if (kHelper && (mode == 3 || (kRamp && mode == 4 && served))) {
        // the helper warp computes this step's gap behaviour; the action is picked up behind the barrier below
      } else if (kMerge && (mode == 3 || mode == 4)) {
        MergeSelf me;
        me.k = k; me.lane = lane; me.front = front; me.x = x; me.y = y; me.v = v; me.l = el;
        me.w = F[(size_t)F_W * np + k];
        me.rss.rTOther = F[(size_t)F_RSS_RTOTHER * np + k]; me.rss.rTEgo = rTEgo; me.rss.aMin = aMin; me.rss.aMax = aMax;
        me.rss.dSoft = F[(size_t)F_RSS_DSOFT * np + k];
        me.idm = pe;
        // mode 3: the ego executes its gap (customizedMerging); mode 4: a non-ego `merging` vehicle picks the closest gap to the left
        const bool fixedGap = !kHelper && mode == 3;
        const bool okSide = merging_behavior(s, sc, laneStart, me, evd, fixedGap, false, fixedGap ? cd.action == MCS_ACT_LEFT : true,
                                             fixedGap ? s.flags[4] : -1, ax, avy);
        if (!okSide) { s.flags[6] = 1; ax = 0.0; avy = 0.0; }
      } else if (mode == 1 || mode == 2) {
        // One idmAccFrontBack call per vehicle and step: the branches only choose its operands.
        //   mode 1  idmBehaviorAndYieldingToMergingAgents 0x104200: own leader, vd (capped for the ego 0x104301)
        //   mode 2  customizedLaneChangeBehavior 0xf5e30: keep / acc (thw 0.9, 1.1 vd) / dcc (2 thw, 0.8 vd) to the own
        //           leader; left / right to {target-lane leader, own leader} with the target-lane follower behind
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
              side_leader_follower(s, laneStart, sl, x, ordered, leader, follower);
              f0 = leader; f1 = front; bk = follower; lateral = true;
            } else {
              evalIdm = false;      // no main lane on that side: ax = 0 unless the RSS emergency fires (0xf6976)
            }
          }
        }
        const double pw = mcs_pow4(v / vdd);
        // the noise draw (0x104d98) does not depend on the acceleration: it starts here, next to the divisions below
        const double nz = mcs_unit31(mcs_rand31(key, drawBase + (uint64_t)nY)) - 0.5;
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
        if (noise) ax = ax + nz;
        if (emergency) ax = aMin;
        if (beh == MCS_BEH_IDM_LC && !isEgo) {
          // laneChangeBehaviorMOBIL 0xf42d0, decision part
          if (commit == MCS_COMMIT_KEEP_LANE && keepStep <= 9) {
            keepStep = keepStep + 1;
          } else if (needMobil) {
            const double u = mcs_unit31(mcs_rand31(key, drawBase + (uint64_t)nY + 1));
            MobilState ms{commit, keepStep, targetLane};
            mobil_decide(load_mobil_geo(s, np, k), sc, np, k, lane, v, el, aMin, F[(size_t)F_YP_PF * np + k], F[(size_t)F_YP_DA * np + k],
                         F[(size_t)F_YP_BA * np + k], /*mode=*/(step == 0) ? 0 : 1, aggressive, F + (size_t)F_PRIOR_L * np, u, ms, ax, avy);
            commit = ms.commit; keepStep = ms.keepStep; targetLane = ms.targetLane;
          }
        }
        }
      }
    }
    sub_sync<NPS, G == 1>(sub);   // every read of the pre-step x / v is done
    if (kRamp && (mode == 3 || (mode == 4 && served))) {
      if (s.hok[k] != 0) { ax = s.hax[k]; avy = s.havy[k]; }
      else { s.flags[6] = 1; ax = 0.0; avy = 0.0; }
    } else if (kHelper == 1 && mode == 3) {
      if (s.ego[3] != 0.0) { ax = s.ego[1]; avy = s.ego[2]; }
      else { s.flags[6] = 1; ax = 0.0; avy = 0.0; }
    }

    // ================= integrate (0x10ea36-0x10eb80) =================
    double acc = 0.0;
    if (live) {
      double v1 = ax * lp.dt + v;
      if (0.0 > v1) {
        v = 0.0; x = 0.0 * lp.dt + x; vy = avy; y = avy * lp.dt + y; acc = 0.0; v1 = 0.0;
      } else {
        v = v1; x = v1 * lp.dt + x; vy = avy; y = avy * lp.dt + y; acc = (v1 == 0.0) ? 0.0 : ax;
      }
      s.x[k] = x; s.v[k] = v;
      s.acc[k] = s.acc[k] + fabs(acc); s.acc[np + k] = s.acc[np + k] + v;
      if (isEgo) { s.chist[step + 1] = acc; s.vhist[step + 1] = v; if (kHelper == 1) s.ego[0] = y; }
      if (kRamp) s.hy[k] = y;
      if (kTraj) {
        double* t = lp.traj + ((((size_t)cand * lp.count + ridx) * n + inputIdx) * (size_t)(lp.steps + 1) + step + 1) * 4;
        t[0] = x; t[1] = y; t[2] = v; t[3] = acc;
      }
    }

    // ================= matchAgentsOnCorridor (0x10b810) =================
    bool changed = false;
    const int laneBefore = lane;      // the lane whose vector holds this vehicle (vectors are rebuilt whenever a lane changes)
    if (live) {
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
        if (kRamp) s.hbeh[k] = (uint8_t)beh;
      }
      if (isEgo && lane >= 0 && s.flags[0] == 0) s.flags[0] = sc.lanes[lane].id == targetLane ? 1 : 0;
    }
    if (kMerge && isEgo) s.flags[2] = beh;
    if (live && lane != laneBefore) {
      // a mover queues itself for the splice below: slot, old position, old lane, new lane
      const int slot = atomicAdd(&s.flags[1], 1);
      s.tasks[slot] = k | (laneBefore >= 0 ? myPos : 0) << 8 | (laneBefore < 0 ? 0xff : laneBefore) << 16 | (lane < 0 ? 0xff : lane) << 24;
    }
    // one barrier publishes the new x / v / lane, the movers' queue and the ego's arrival, and tells whether any membership changed
    const int anyChanged = sub_sync_or<NPS, G == 1>(sub, changed ? 1 : 0);
    const bool egoReached = s.flags[0] != 0;
    if (anyChanged) {
      // The reference rebuilds every lane vector (members in visit order) and std::sorts it by x (0x10be30-0x10c1a5):
      // the result is "members ordered by (x, slot)".  Vehicles that kept their lane are already in that order unless
      // two of them passed through each other since the last rebuild, so the usual case only has to splice the few
      // movers (queued above) in.  In the coordinates of the old concatenated vector a mover is two events: it leaves
      // position q (everything behind q moves up one) and it is inserted in front of position `ins` of its new lane's
      // old vector (everything from `ins` on moves back one).  The movers' searches run on the first threads of the
      // rollout side by side (one mover per thread, not one divergent search per warp that owns a mover); every other
      // vehicle adds up the events in front of it: integer compares only.  An order inversion anywhere falls back to
      // ranking every vehicle against every other one.
      const int newLane = live ? (int)s.lane[k] : 0xff;
      const int oldL = laneBefore < 0 ? 0xff : laneBefore;
      const bool inVec = live && oldL != 0xff;
      const bool mover = live && newLane != oldL;
      bool inversion = false;
      if (inVec && myPos > laneStart[oldL]) {
        const int pj = s.vec[myPos - 1];
        const double xp = s.x[pj];
        inversion = xp > x || (xp == x && pj > k);
      }
      const int nm = s.flags[1];
      int* ev = (int*)s.mob;        // [np] event words of the movers (the MOBIL results were consumed in part B)
      int myIns = 0xffff, myRec = 0;
      if (tid < nm) {
        myRec = s.tasks[tid];
        const int mk = myRec & 0xff, mpos = myRec >> 8 & 0xff, mold = myRec >> 16 & 0xff, mnew = myRec >> 24 & 0xff;
        if (mnew != 0xff) {
          // members of the destination lane with a smaller (x, slot) key: binary search in its (ordered) old vector
          const int lo = laneStart[mnew], cntL = laneStart[mnew + 1] - lo;
          const double xm = s.x[mk];
          int first = 0, len = cntL;
          while (len > 0) {
            const int half = len >> 1, mid = first + half;
            const int j = s.vec[lo + mid];
            const double xj = s.x[j];
            if (xj < xm || (xj == xm && j < mk)) { first = mid + 1; len = len - half - 1; }
            else len = half;
          }
          myIns = lo + first;
        }
        ev[tid] = (int)((unsigned)(mold != 0xff ? mpos + 1 : 0xffff) | (unsigned)myIns << 16);
      }
      if (k < MCS_MAX_LANES + 1) s.scan[16 + k] = 0;
      const int anyInv = sub_sync_or<NPS, G == 1>(sub, inversion ? 1 : 0);
      int* nextStart = laneStart == s.cnt ? s.cnt + 16 : s.cnt;
      if (anyInv) {
        int rank = 0;
        if (newLane != 0xff) {
          for (int j = 0; j < n; ++j) {
            const double xj = s.x[j];
            rank += (s.lane[j] == newLane && (xj < x || (xj == x && j < k))) ? 1 : 0;
          }
          atomicAdd(&s.scan[16 + newLane], 1);
        }
        sub_sync<NPS, G == 1>(sub);
        // every thread derives the new lane starts itself (≤ 8 lanes) and places its vehicle; the shared copy goes to the
        // other buffer, so one barrier publishes the vectors and the starts together
        int run = 0, myStart = 0;
        for (int j = 0; j < sc.n_lanes; ++j) {
          const int c = s.scan[16 + j];
          if (j == newLane) myStart = run;
          if (k == 0) nextStart[j] = run;
          run += c;
        }
        if (k == 0) nextStart[sc.n_lanes] = run;
        if (newLane != 0xff) { myPos = myStart + rank; s.vec[myPos] = (uint8_t)k; }
      } else {
        if (inVec && !mover) {
          int d = 0;
#pragma unroll 1
          for (int q = 0; q < nm; ++q) {
            const unsigned e = (unsigned)ev[q];      // [leave + 1 | insert << 16], 0xffff = none
            d += ((int)(e >> 16) <= myPos ? 1 : 0) - ((int)(e & 0xffffu) <= myPos ? 1 : 0);
          }
          myPos += d;
          s.vec[myPos] = (uint8_t)k;
        }
        if (tid < nm && myIns != 0xffff) {
          // the mover this thread looks after: its own removal and the other movers' events in front of its insertion point;
          // arrivals at the same point go by (new lane, x, slot) — a lane's end is the next lane's start
          const int mk = myRec & 0xff, mnew = myRec >> 24 & 0xff;
          const double xm = s.x[mk];
          int pos = myIns;
#pragma unroll 1
          for (int q = 0; q < nm; ++q) {
            const unsigned e = (unsigned)ev[q];      // [leave + 1 | insert << 16], 0xffff = none
            pos -= (int)(e & 0xffffu) <= myIns ? 1 : 0;
            const int insq = (int)(e >> 16);
            if (q == tid || insq > myIns) continue;
            bool ahead = insq < myIns;
            if (!ahead) {
              const int rq = s.tasks[q];
              const int qk = rq & 0xff, qnew = rq >> 24 & 0xff;
              if (qnew != mnew) ahead = qnew < mnew;
              else { const double xq = s.x[qk]; ahead = xq < xm || (xq == xm && qk < mk); }
            }
            pos += ahead ? 1 : 0;
          }
          s.vec[pos] = (uint8_t)mk;
          s.pos[mk] = (uint8_t)pos;
        }
        // new lane starts: old start + arrivals into - departures from the lanes in front (a few threads of the second warp,
        // next to the movers' threads of the first)
        constexpr int kStartBase = NPS >= 64 ? 32 : 0;
        if (tid >= kStartBase && tid <= kStartBase + sc.n_lanes) {
          const int j = tid - kStartBase;
          int st = laneStart[j];
#pragma unroll 1
          for (int q = 0; q < nm; ++q) {
            const int rq = s.tasks[q];
            st += ((rq >> 24 & 0xff) < j ? 1 : 0) - ((rq >> 16 & 0xff) < j ? 1 : 0);
          }
          nextStart[j] = st;
        }
      }
      laneStart = nextStart;
      sub_sync<NPS, G == 1>(sub);
      if (!anyInv && mover && newLane != 0xff) myPos = s.pos[k];
    }
    // ================= termination (0x10edd0-0x10ee10) =================
    const int gapSlot = (kMerge && !kHelper && isEgo) ? s.flags[4] : -1;
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
    step++;
    if (egoReached) {
      if (isEgo) s.flags[5] = min(s.flags[5], step - 1);
      if (lp.min_steps <= step - 1) { if (G > 1) { running = false; continue; } else break; }
    }
  }

  // ================= outcome statistics (0x10ee20-0x10f64d) =================
  const int executed = step, nStates = step + 1;
  double utilO = 1e300, comfO = 1e300;   // min identity; replaced by 1.0 when the rollout has no objects
  const double aMin = s.rss[np + k];
  if (live && !isEgo) {
    const double cnt = (double)nStates, evd = s.vd[k];
    const double meanV = s.acc[np + k] / cnt;
    comfO = (s.acc[k] / cnt) / aMin + 1.0;
    utilO = (meanV >= evd) ? 1.0 : meanV / evd;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double u2 = __shfl_xor_sync(0xffffffffu, utilO, o), c2 = __shfl_xor_sync(0xffffffffu, comfO, o);
    utilO = MCS_MINSD(u2, utilO); comfO = MCS_MINSD(c2, comfO);
  }
  if ((tid & 31) == 0) { s.red[(tid >> 5) * 2] = utilO; s.red[(tid >> 5) * 2 + 1] = comfO; }
  sub_sync<NPS, G == 1>(sub);
  // ego history -> per-state terms, one state per thread: fall-back = two consecutive states at rss.aMin (0x10f797),
  // utility term u(v/vd) (0x10f0c4), comfort sample (0x10f1c3)
  const int egoSlot = cd.ego_slot;
  const double egoAMin = F[(size_t)F_RSS_AMIN * np + egoSlot], egoVd = s.vd[egoSlot];
  bool fb = false;
  for (int t = tid; t < nStates; t += NPS) {
    const double a = s.chist[t];
    if (t + 1 < nStates && egoAMin == a && egoAMin == s.chist[t + 1]) fb = true;
  }
  const int anyFallBack = sub_sync_or<NPS, G == 1>(sub, fb ? 1 : 0);
  for (int t = tid; t < nStates; t += NPS) {
    const double a = s.chist[t];
    s.chist[t] = a > 0.0 ? fabs(a) * 0.5 : fabs(a);
    const double ve = s.vhist[t];
    const double r = ve / egoVd;
    double u = r;
    if (!(egoVd >= ve)) u = MCS_MAXSD(0.9, 1.0 - (r - 1.0));
    s.vhist[t] = u;
  }
  sub_sync<NPS, G == 1>(sub);
  // ego comfort: mean over the k largest samples of 1 + c/aMin, summed in descending order (0x10f1c3-0x10f49a)
  {
    const int ns = nStates;
    double* sorted = s.csort;
    for (int t = tid; t < ns; t += NPS) {
      const double c = s.chist[t];
      int rank = 0;
#pragma unroll 1
      for (int j = 0; j < ns; ++j) { const double cj = s.chist[j]; rank += (cj > c || (cj == c && j < t)) ? 1 : 0; }
      sorted[rank] = c;
    }
  }
  sub_sync<NPS, G == 1>(sub);
  if (isEgo) {
    constexpr int nw = NPS >> 5;
    double uo = s.red[0], co = s.red[1];
    for (int q = 1; q < nw; ++q) { uo = MCS_MINSD(s.red[2 * q], uo); co = MCS_MINSD(s.red[2 * q + 1], co); }
    if (n <= 1) { uo = 1.0; co = 1.0; }
    mcs_status st;
    memset(&st, 0, sizeof st);
    st.utilityObjects = uo; st.comfortObjects = co;
    double sumU = 0.0;
#pragma unroll 1
    for (int t = 0; t < nStates; ++t) sumU = sumU + s.vhist[t];
    const double util = sumU / (double)nStates;
    const int thwCnt = s.flags[3];
    double thwAcc = 0.0;                                  // 0x10f14b-0x10f19f
#pragma unroll 1
    for (int t = 0; t < thwCnt; ++t) { const double e = s.thist[t]; if (0.9 > e && e > 0.0) thwAcc = (e + thwAcc) - 0.9; }
    const double thwTerm = thwCnt > 0 ? thwAcc / (double)thwCnt : 0.0;
    st.utility = util + thwTerm;
    const int kk = (int)((double)nStates * 0.2);
    double sumC = 0.0;
#pragma unroll 1
    for (int t = 0; t < kk; ++t) sumC = sumC + (s.csort[t] / aMin + 1.0);
    st.comfort = sumC / (double)kk;
    st.finishStep = max(s.flags[5], (int)(1.0 / lp.dt)); st.executedSteps = executed; st.draws = (uint32_t)drawCtr;
    st.finish = s.flags[5] < lp.steps ? 1 : 0;
    const int beh0 = I[(size_t)I_BEH0 * np + k];
    if (beh0 == MCS_BEH_MERGING || beh0 == MCS_BEH_EXIT) st.fallBack = (2.0 > v) ? 1 : 0;
    else st.fallBack = anyFallBack ? 1 : 0;
    st.ok = 1;
    st.overflow = 0;
    *st_out = st;
    if (kTraj && lp.traj_n) lp.traj_n[(size_t)cand * lp.count + ridx] = nStates;
    if (lp.vehicle_steps) atomicAdd(lp.vehicle_steps, (unsigned long long)executed * (unsigned long long)n);
  }
  sub_sync<NPS, G == 1>(sub);
  const int anyUnsupported = s.flags[6], anyFull = s.flags[7];
  if ((anyUnsupported || anyFull) && isEgo) {
    // the caller fails loudly (MCS_ERR_SCENE / MCS_ERR_CAPACITY): the record carries the mark, and so does the launch-wide
    // word every entry point checks after its synchronisation — whichever rollout of whichever candidate it happened in
    st_out->overflow = (uint8_t)(anyUnsupported ? 2 : 1);
    if (lp.err_flags) atomicOr(lp.err_flags, anyUnsupported ? 2u : 1u);
  }
}

}  // namespace mcs

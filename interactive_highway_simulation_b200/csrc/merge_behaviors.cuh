// Device functions of the merging / exit behaviours (included by rollout_kernel.cuh after its helpers).
//
//   customizedMergingBehavior  mc_sim_highway.so 0xf6ba0   ego executes a fixed gap (fixedGapMergingBehavior)
//   mergingBehaviorClosestGap  0xfe660                     gap selection by time-to-reach (closestGapMergingBehavior,
//                                                           and every non-ego `merging` vehicle inside a rollout)
//   rssSafeRelax               0xb9210                     two-second look-ahead safety check of a gap
//   idmAccFrontBack (IdmMatch) 0xeb8d0                     the IDM variant rssSafeRelax advances the ego with
//   timeToReachGap             0xe1630
//   yielding intentions        0x1045b0-0x1048f4           (idmBehaviorAndYieldingToMergingAgents)
// Same rules as the rest of the kernel: binary64, the operation order of the machine code, no contraction.
#pragma once

namespace mcs {

// idmAccFrontBack 0xebc30 with an arbitrary list of fronts (slots, -1 = null)
__device__ __forceinline__ double idm_acc_list(const Smem& s, const IdmP& p, const int* fronts, int nf, int back,
                                               double egoX, double egoV, double egoL, double pw) {
  const double freeAcc = (1.0 - pw) * p.accMax;
  const double na = -p.accMax;
  const double sq = sqrt(na * p.accCom);
  const double sq2 = sq + sq;
  const double dDes = (egoV * p.thw + p.minDis) * 0.7;
  bool any = false;
  double d = 0.0;
#pragma unroll 1
  for (int q = 0; q < nf; ++q) {
    const int f = fronts[q];
    if (f < 0) continue;
    const double gap = (s.x[f] - egoX) - (egoL + s.l[f]) * 0.5;
    double term;
    if (gap > 0.0) {
      const double dv = (egoV - s.v[f]) * egoV;
      double t = dv / sq2 + dDes;
      t = MCS_MAXSD(t, 0.0);
      t = mcs_div(t, gap);          // t is clamped at 0: keep the zero off the division's slow path
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

// idmAccFrontBack(IdmParam, vector<IdmMatch>{one}, IdmMatch back, v, vd) 0xeb8d0: accMin where the Agent overload has -accMax
__device__ __forceinline__ double idm_acc_match(const IdmP& p, int fid, double fdis, double fv, int bid, double bdis,
                                                double bv, double v, double vd) {
  const double freeAcc = (1.0 - mcs_pow4(v / vd)) * p.accMax;
  double sum = 0.0;
  if (fid >= 0) {
    double term;
    if (fdis > 0.0) {
      const double dv = (v - fv) * v;
      const double dDes = (v * p.thw + p.minDis) * 0.7;
      const double sq = sqrt(p.accMin * p.accCom);
      double t = dv / (sq + sq) + dDes;
      t = MCS_MAXSD(t, 0.0);
      t = mcs_div(t, fdis);
      term = (t * t) * p.accMin;
    } else {
      term = -10000000000.0;
    }
    sum = term + 0.0;
  }
  if (bid >= 0) {
    if (0.0 > bdis) {
      const double dDes = (bv * p.thw + p.minDis) * 0.7;
      const double dv = (bv - v) * v;
      const double sq = sqrt(p.accMin * p.accCom);
      double t = dv / (sq + sq) + dDes;
      t = MCS_MAXSD(t, 0.0);
      t = mcs_div(t, bdis);
      sum = sum + (t * t) * p.accMax;
    } else {
      sum = sum + 10000000000.0;
    }
  }
  const double r = MCS_MINSD(sum + freeAcc, p.accMax);
  return MCS_MAXSD(p.accMin, r);
}

struct RssP { double rTOther, rTEgo, aMin, aMax, dSoft; };

// rssSafeRelax 0xb9210: entity 1 = leader of the gap's front vehicle, 3 = the gap's front vehicle, 5 = the gap's rear vehicle
// Inlined like every helper here, each at its single call site: a call that takes `s` (or any struct) by reference forces
// the Smem pointer table into local memory, and then EVERY shared-memory access of the kernel becomes a generic load
// through a pointer re-read from the stack (SASS of the merging instantiation: 3 LDS / 145 LD / 83 LDL before,
// 450 LDS / 0 LD after; 4 x 512 x 40-vehicle decision 2.43 -> 2.03 ms).
__device__ __forceinline__ bool rss_safe_relax(bool h1, double gap1, double v1, bool h3, double gap3, double v3, bool h5,
                                               double gapb, double vb, double egoV, double vd, const RssP& r, const IdmP& idm) {
  double v = egoV;
  const int id1 = h1 ? 1 : -1, id3 = h3 ? 1 : -1, idb = h5 ? 1 : -1;
  if (!h1) { gap1 = 0.0; v1 = 0.0; }
  if (!h3) { gap3 = 0.0; v3 = 0.0; }
  if (!h5) { gapb = 0.0; vb = 0.0; }
#pragma unroll 1
  for (int it = 2;; it = 1) {
    const double aMaxHalf = r.aMax * 0.5;
    double a3 = aMaxHalf;
    if (id1 == 1) {
      const double num = (-0.5 * v3) * v3;
      double den = (v1 * v1) / (r.aMax * -2.0);
      den = den + (gap1 - gap3);
      den = den - v3 * r.rTOther;
      den = MCS_MAXSD(den, 0.01);
      a3 = num / den;
      a3 = MCS_MAXSD(r.aMax, a3);
      a3 = MCS_MINSD(aMaxHalf, a3);
    }
    double aEgo = aMaxHalf;
    if (id3 != -1) {
      const double q = mcs_div(v3 * v3, a3 + a3);
      const double vr = v * r.rTEgo;
      double d = (v * v) / (-2.0 * r.aMin);
      d = d + vr;
      d = d + q;
      d = MCS_MAXSD(0.0, d);
      const double num = (v * -0.5) * v;
      double den = (v3 * v3) / (r.aMax * -2.0);
      den = den + gap3;
      den = den - vr;
      den = MCS_MAXSD(den, 0.01);
      double ae = num / den;
      ae = MCS_MAXSD(r.aMax, ae);
      aEgo = MCS_MINSD(aMaxHalf, ae);
      if (d > gap3) return false;
    }
    if (idb >= 0) {
      const double q = (v * v) / (aEgo + aEgo);
      double d = (vb * vb) / (-2.0 * r.aMin);
      d = d + vb * r.rTOther;
      d = d + q;
      d = MCS_MAXSD(0.0, d);
      if (d > -gapb) return false;
    }
    const double acc = idm_acc_match(idm, id3, gap3, v3, idb, gapb, vb, v, vd);
    double sdist = acc * 0.5 + v;
    const double vn = v + acc;
    v = MCS_MAXSD(0.0, vn);
    sdist = MCS_MAXSD(sdist, 0.0);
    if (id1 != -1) gap1 = (v1 + gap1) - sdist;
    if (id3 != -1) gap3 = (v3 + gap3) - sdist;
    if (idb >= 0) {
      gapb = ((vb + gapb) - r.dSoft * 0.5) - sdist;
      const double t = vb - r.dSoft;
      vb = MCS_MAXSD(t, 0.0);
    }
    if (it == 1) return true;
  }
}

// timeToReachGap(xBack, vBack, xFront, vFront, xEgo, vEgo, vTarget) 0xe1630
__device__ __forceinline__ double time_to_reach_gap(double xB, double vB, double xF, double vF, double xE, double vE, double vT) {
  const double a = (vT - vE) / 5.0;
  if (xB > xE) {
    const double acc = (a > 2.0) ? 2.0 : MCS_MAXSD(0.5, a);
    const double dv = vB - vE;
    const double dx = xB - xE;
    const double inv = 1.0 / acc;
    const double disc = dx * (acc + acc) + dv * dv;
    const double sq = sqrt(disc);
    const double r1 = (dv + sq) * inv;
    const double r2 = (dv - sq) * inv;
    return MCS_MAXSD(r1, r2);
  }
  if (xE > xF) {
    const double m = vE - vF;
    const double dv = vF - vE;
    const double disc = dv * dv - (xF - xE) * 4.0;
    const double sq = sqrt(disc);
    const double r1 = (m + sq) * 0.5;
    const double r2 = (m - sq) * 0.5;
    return MCS_MAXSD(r1, r2);
  }
  const double X = (fabs(vB - vE) < fabs(vF - vE)) ? vB : vF;
  double acc = a;
  if (a > 2.0) acc = 2.0;
  else if (0.5 > a) acc = 0.5;
  const double h = X * 0.5;
  if (vE > h) return 0.0;
  const double t = h - vE;
  return (t + t) / acc;
}

// one vehicle as the merging code sees it
struct MergeSelf {
  int k, lane, front;
  double x, y, v, l, w;
  RssP rss;
  IdmP idm;
};

// Neighbor::left()/right() 0xeae70: [BB, B, F, FF] of lane `ln` without the nulls; returns the count
__device__ __forceinline__ int side_vector(const Smem& s, const int* laneStart, int ln, double x, int out[4]) {
  const int lo = laneStart[ln], n = laneStart[ln + 1] - lo;
  const int f = lane_lower(s, lo, n, x);
  const int r = lane_reverse(s, lo, n, x);
  int c = 0;
  if (r > 1) out[c++] = s.vec[lo + r - 2];
  if (r > 0) out[c++] = s.vec[lo + r - 1];
  if (f < n) out[c++] = s.vec[lo + f];
  if (f + 1 < n) out[c++] = s.vec[lo + f + 1];
  return c;
}

// Environment::frontAgent 0xe2670 of another vehicle g (its own lane vector)
__device__ __forceinline__ int front_of(const Smem& s, const int* laneStart, int g) {
  const int ln = s.lane[g];
  if (ln == 0xff) return -1;
  const int lo = laneStart[ln], n = laneStart[ln + 1] - lo;
  const int f = lane_upper(s, lo, n, s.x[g]);
  return f < n ? (int)s.vec[lo + f] : -1;
}

// common tail 0xf7155-0xf72e8 / 0xff02e-0xff10e: RSS emergency to the own leader, room left on the own lane, lateral motion
__device__ __forceinline__ void merge_finish(const Smem& s, const SceneDev& sc, const MergeSelf& a, int side, bool left,
                                             bool safe, double axGap, double& ax, double& vy) {
  const LaneDev& c = sc.lanes[a.lane];
  const double brake = (a.v * a.v) / (a.rss.aMin * -2.0) + a.v * a.rss.rTEgo;
  bool restricted = false;
  if (a.front >= 0) {
    const double gap = (s.x[a.front] - a.x) - (s.l[a.front] + a.l) * 0.5;
    const double fv = s.v[a.front];
    double d = (fv * fv) / (a.rss.aMax + a.rss.aMax) + brake;
    d = MCS_MAXSD(0.0, d);
    if (d > gap) restricted = true;
  }
  if (!restricted) {
    const bool room = (c.endX - a.x) - 15.0 >= brake;
    if (!room && !safe) restricted = true;
  }
  ax = restricted ? a.rss.aMin : axGap;
  const bool lateral = restricted ? false : safe;
  const bool far = left ? (-0.5 > (a.w * 0.5 + a.y) - c.leftY) : ((a.y - a.w * 0.5) - c.rightY > 0.5);
  const double sign = left ? 1.0 : -1.0;
  const double m = a.v * 0.17;
  const bool toward = far || (lateral && a.x >= sc.lanes[side].startX);
  if (toward) vy = (m > 1.0) ? sign : m * sign;
  else vy = ((m > 1.0) ? -1.0 : -m) * sign;
}

__device__ __forceinline__ bool gap_is_safe(const Smem& s, const MergeSelf& a, int gff, int gf, int g, double vd) {
  const bool h1 = gff >= 0, h3 = gf >= 0, h5 = g >= 0;
  const double g1 = h1 ? (s.x[gff] - a.x) - (s.l[gff] + a.l) * 0.5 : 0.0;
  const double g3 = h3 ? (s.x[gf] - a.x) - (s.l[gf] + a.l) * 0.5 : 0.0;
  const double g5 = h5 ? (s.l[g] + a.l) * 0.5 + (s.x[g] - a.x) : 0.0;
  return rss_safe_relax(h1, g1, h1 ? s.v[gff] : 0.0, h3, g3, h3 ? s.v[gf] : 0.0, h5, g5, h5 ? s.v[g] : 0.0, a.v, vd, a.rss, a.idm);
}

// customizedMergingBehavior 0xf6ba0 (fixedGap = true; gapSlot: slot of the gap's rear vehicle, < 0 = "last gap") and
// mergingBehaviorClosestGap 0xfe660 (fixedGap = false) share everything after the choice of the gap — its front vehicle
// gf, that one's leader gff and its rear vehicle g: rssSafeRelax on the three, one idmAccFrontBack, and the common tail.
// One body, so that the merging instantiation (instruction-cache bound) carries each formula once.  Returns false when
// the reference would dereference a missing corridor (no lane on that side).
__device__ __forceinline__ bool merging_behavior(const Smem& s, const SceneDev& sc, const int* laneStart, const MergeSelf& a,
                                                 double evd, bool fixedGap, bool isEgo, bool left, int gapSlot, double& ax,
                                                 double& vy) {
  const LaneDev& c = sc.lanes[a.lane];
  const int side = left ? c.leftIdx : c.rightIdx;
  if (side < 0) return false;
  // customizedMerging caps vd for everybody (0xf6f90), closest-gap only for the ego (0xfe757)
  const double vd = (fixedGap || isEgo) ? MCS_MINSD(evd, 1.1 * c.vLimit) : evd;
  const double pw = mcs_pow4(a.v / vd);
  int g = -1, gf = -1, gff = -1;
  int fronts[3];
  int nf;
  bool execute = true;
  if (fixedGap && gapSlot >= 0) {
    g = gapSlot;
    gf = front_of(s, laneStart, g);
    if (gf >= 0) gff = front_of(s, laneStart, gf);
  } else {
    int v4[4];
    const int n = side_vector(s, laneStart, side, a.x, v4);
    if (fixedGap) {                         // "go to last gap": behind the rearmost neighbour (0xf7488)
      if (n >= 1) gf = v4[0];
      if (n >= 2) gff = v4[1];
    } else if (n == 0) {
      execute = false;                      // nobody on the target lane: plain IDM to the own leader (0xff2bf)
    } else {
      // std::map<int,double> of time-to-reach per gap; gap j lies behind v4[j], gap n in front of the last neighbour;
      // strict > keeps the smallest key on ties (0xfed88-0xfedaf)
      double best = 0.0;
      int kbest = 0;
#pragma unroll 1
      for (int j = 0; j <= n; ++j) {
        const int B = j > 0 ? v4[j - 1] : -1, F = j < n ? v4[j] : -1;
        const double xB = B >= 0 ? (a.l + s.l[B]) * 0.5 + s.x[B] : -10000000000.0;
        const double vB = B >= 0 ? s.v[B] : 0.0;
        const double xF = F >= 0 ? s.x[F] - (a.l + s.l[F]) * 0.5 : 10000000000.0;
        const double vF = F >= 0 ? s.v[F] : 100.0;
        const double t = time_to_reach_gap(xB, vB, xF, vF, a.x, a.v, vd);
        if (j == 0 || best > t) { best = t; kbest = j; }
      }
      gf = kbest < n ? v4[kbest] : -1;
      gff = kbest + 1 < n ? v4[kbest + 1] : -1;
      g = kbest > 0 ? v4[kbest - 1] : -1;
    }
  }
  if (fixedGap) { fronts[0] = gf; fronts[1] = gff; fronts[2] = a.front; nf = 3; }      // 0xf707a-0xf70d9
  else if (execute) { fronts[0] = gf; nf = 1; }                                        // 0xff370 / 0xff610 / 0xfee31
  else { fronts[0] = a.front; nf = 1; }
  const double axI = idm_acc_list(s, a.idm, fronts, nf, execute ? g : -1, a.x, a.v, a.l, pw);
  if (!execute) {
    const bool far = left ? (-0.5 > (a.w * 0.5 + a.y) - c.leftY) : ((a.y - a.w * 0.5) - c.rightY > 0.5);
    ax = axI;
    vy = 0.0;
    if (a.x >= sc.lanes[side].startX || far) {
      const double sign = left ? 1.0 : -1.0;
      const double m = a.v * 0.17;
      vy = (m > 1.0) ? sign : m * sign;
    }
    return true;
  }
  const bool safe = gap_is_safe(s, a, gff, gf, g, vd);
  merge_finish(s, sc, a, side, left, safe, axI, ax, vy);
  return true;
}

__device__ __forceinline__ bool customized_merging(const Smem& s, const SceneDev& sc, const int* laneStart, const MergeSelf& a,
                                                   double evd, bool left, int gapSlot, double& ax, double& vy) {
  return merging_behavior(s, sc, laneStart, a, evd, true, false, left, gapSlot, ax, vy);
}
__device__ __forceinline__ bool merging_closest_gap(const Smem& s, const SceneDev& sc, const int* laneStart, const MergeSelf& a,
                                                    double evd, bool isEgo, bool left, double& ax, double& vy) {
  return merging_behavior(s, sc, laneStart, a, evd, false, isEgo, left, -1, ax, vy);
}

// ---- yielding intentions (Agent+0x238: unordered_map<int, YieldingIntention{id, p0, threshold 0.625, yielding}>) ----
// Two stores with one interface.  update(): `p` = yielding probability now, `y0` = the initial intention drawn for a first
// encounter (0x10469c-0x1046ec); returns the intention after the hysteresis update (0x104724: a non-yielding vehicle starts
// once (1-p0) 0.625 >= 1-p, a yielding one stops once 0.625 p0 >= p).  A full store sets flags[7] (MCS_ERR_CAPACITY) and
// drops the entry.
__device__ __forceinline__ bool yield_hysteresis(bool yl, double p0, double p) {
  if (!yl) { if ((1.0 - p0) * 0.625 >= 1.0 - p) yl = true; }
  else { if (0.625 * p0 >= p) yl = false; }
  return yl;
}

// rollout_kernel: a table of 32 entries per vehicle thread (local memory; L1-resident).  The shared-memory pool below was
// measured in its place and is slower there: 3 KB more per rollout no longer fit fourteen 64-thread rollouts per block
// (configs[1] 1.60 -> 1.82 ms), and a pool sized per rollout overflows on long horizons where 32 per vehicle do not.
#define MCS_YCAP 32
struct YieldLocal {
  uint8_t key[MCS_YCAP];
  double p0[MCS_YCAP];
  uint32_t yielding;   // bit per entry
  int n;
  __device__ __forceinline__ void init() { n = 0; yielding = 0u; }
  __device__ __forceinline__ bool update(const Smem& s, int j, double p, bool y0) {
    int e = -1;
    for (int t = 0; t < n; ++t) if (key[t] == (uint8_t)j) { e = t; break; }
    if (e < 0) {
      if (n >= MCS_YCAP) { s.flags[7] = 1; return false; }
      e = n++;
      key[e] = (uint8_t)j; p0[e] = p;
      if (y0) yielding |= 1u << e; else yielding &= ~(1u << e);
    }
    const bool yl = yield_hysteresis((yielding >> e) & 1u, p0[e], p);
    if (yl) yielding |= 1u << e; else yielding &= ~(1u << e);
    return yl;
  }
};

// bundle_kernel: ONE pool of entries per rollout in shared memory (Smem::yp0 / ykey / ynext / yflag, handed out by an
// atomic counter); every vehicle threads its own entries into a chain whose head lives in a register.  Chains are short (a
// vehicle meets a handful of distinct candidates per rollout) and nothing sits in local memory.
struct YieldChain {
  int head;
  __device__ __forceinline__ void init() { head = YIELD_END; }
  __device__ __forceinline__ bool update(const Smem& s, int j, double p, bool y0) {
    int e = head;
    while (e != YIELD_END && s.ykey[e] != (uint8_t)j) e = s.ynext[e];
    if (e == YIELD_END) {
      e = atomicAdd(&s.flags[8], 1);
      if (e >= s.ycap) { s.flags[7] = 1; return false; }
      s.ykey[e] = (uint8_t)j; s.yp0[e] = p; s.ynext[e] = (uint16_t)head; s.yflag[e] = y0 ? 1 : 0;
      head = e;
    }
    const bool yl = yield_hysteresis(s.yflag[e] != 0, s.yp0[e], p);
    s.yflag[e] = yl ? 1 : 0;
    return yl;
  }
};

// The part of idmBehaviorAndYieldingToMergingAgents that looks at the yielding candidates (0x1045b0-0x1048f4): per
// candidate the intention probability — logistic of (time headway, relative speed, own initial acceleration) — one draw
// each (draws drawBase .. drawBase + nY - 1), the hysteresis update, then idmAccFrontBack with the softened parameters
// `py` over the candidates the vehicle currently yields to.
template <class YieldStore>
__device__ __forceinline__ double yielding_acc(const Smem& s, const double* __restrict__ F, int np, int k, const IdmP& py,
                                               const int* ycand, int nY, uint64_t key, uint64_t drawBase, YieldStore& yt, double x,
                                               double v, double el, double pw) {
  const double ypBias = F[(size_t)F_YP_BIAS * np + k], ypA = F[(size_t)F_YP_A * np + k], ypB = F[(size_t)F_YP_B * np + k],
               ypC = F[(size_t)F_YP_C * np + k];
  int ylist[5];
  int ny = 0;
#pragma unroll 1
  for (int q = 0; q < nY; ++q) {
    const int j = ycand[q];
    double p = 0.0;
    const double half = (el + s.l[j]) * 0.5;
    const double dx = s.x[j] - x;
    if (!(0.0 > dx + half)) {
      const double thw = (dx - half) / MCS_MAXSD(1e-10, v);
      const double rel = (v - s.v[j]) / MCS_MAXSD(v, 0.001);
      double r = thw * ypA + ypBias;
      r = r + rel * ypB;
      r = r + F[(size_t)F_A * np + k] * ypC;     // Agent::statesHistory[0].a
      double arg;
      if (100.0 > r) arg = (-100.0 > r) ? 100.0 : -r;
      else arg = -100.0;
      p = 1.0 / (mcs_exp(arg) + 1.0);
    }
    const int rnd = mcs_rand31(key, drawBase + (uint64_t)q);
    const bool y0 = (100.0 * p) >= (double)(rnd % 100);
    if (yt.update(s, j, p, y0)) ylist[ny++] = j;
  }
  return idm_acc_list(s, py, ylist, ny, -1, x, v, el, pw);
}

}  // namespace mcs

// TEST INFRASTRUCTURE ONLY — see mcsim_oracle.h.  Scalar restatement of the reference's inner
// simulator from its disassembly (mc_sim_highway.so; addresses cited as so:0x...).
//
// Floating-point discipline: the reference binary contains no FMA instruction and evaluates
// everything in scalar binary64, so this file is compiled with -O2 -ffp-contract=off and every
// expression below is written in the exact operation order of the machine code.  x86 vminsd /
// vmaxsd return their SECOND source when the comparison is false or unordered; MINSD/MAXSD
// below keep that operand order (first argument = first AVX source).
#include "mcsim_oracle.h"
#include "../include/mcsim_math.h"
#include "../include/mcsim_rng.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

int g_math_mode = 0;
inline double EXP(double x) { return g_math_mode ? std::exp(x) : mcs_exp(x); }
inline double POW4(double x) { return g_math_mode ? std::pow(x, 4.0) : mcs_pow4(x); }
inline double MINSD(double a, double b) { return a < b ? a : b; }
inline double MAXSD(double a, double b) { return a > b ? a : b; }

struct Idm { double accMax, minDis, accMin, accCom, aE, thw; };
struct Rss { double rTOther, rTEgo, aMin, aMax, aSoft, dSoft; };
struct State { double x, y, v, a; };
struct Intention { double p0; double thr; bool yielding; };

struct Corr {
  int id, type;
  double leftX1, leftY, leftX2, rightX1, rightY, rightX2, centerX1, centerY, centerX2, centerY2;
  double vLimit;
  bool hasLeft = false, hasRight = false;
  int leftId = 0, rightId = 0;
};

struct Agent {
  int id, beh, beh0;
  double x, y, v, vy, a, l, w, vd;
  double prior[3] = {0, 0, 0};  // left, keep, right (Agent+0x48)
  bool isEgo = false, reached = false;
  int targetGapId = -2, targetLaneId = -1;
  int egoCmd = MCS_ACT_KEEP_LANE;
  int commit = MCS_COMMIT_NOT_DECIDED, keepStep = 0;
  bool aggressive = false;
  Idm idm, idmY;
  Rss rss, rssR;
  double yp[7];
  std::vector<State> states;
  std::vector<double> thw;
  std::vector<int> lanes;
  std::map<int, Intention> intents;
  int lane() const { return lanes.back(); }
};

struct Rng {
  uint64_t key = 0, ctr = 0;
  int next() { return mcs_rand31(key, ctr++); }
};

struct Action { double ax, vy; };

struct Env {
  std::vector<Corr> corr;        // input order
  std::vector<int> corrOrder;    // iteration order of unordered_map<int,Corridor> → index into corr
  std::unordered_map<int, int> corrIdx;
  int leftmostLaneId = -1;       // MapMultiLane+0x18
  bool hasExitLanes = false;     // byType.find("exit")
  std::vector<Agent> agents;     // input order
  std::vector<int> visit;        // iteration order of unordered_map<int,AgentPtr> → index into agents
  std::unordered_map<int, int> agentIdx;
  std::map<int, std::vector<int>> laneVec;  // agentsOnCorridor: corridor id → agent indices
  int ego = -1;                  // Environment+0xc0
  int egoId = -1, targetGapId = -2, stepCount = 0;
  Rng rng;
  bool error = false;

  Corr* corrById(int id) {
    auto it = corrIdx.find(id);
    if (it == corrIdx.end()) { error = true; return nullptr; }
    return &corr[it->second];
  }
};

// ---------------------------------------------------------------------------------------------
// MapMultiLane::MapMultiLane(vector<Corridor>)  so:0x109f00-0x10a880
void build_map(Env& e, const mcs_corridor* cs, int n) {
  std::unordered_map<int, int> order;  // same insertion sequence as corridorsById.emplace(id, c)
  for (int i = 0; i < n; ++i) {
    Corr c;
    c.id = cs[i].id; c.type = cs[i].type;
    c.leftX1 = cs[i].left[0]; c.leftY = cs[i].left[1]; c.leftX2 = cs[i].left[2];
    c.rightX1 = cs[i].right[0]; c.rightY = cs[i].right[1]; c.rightX2 = cs[i].right[2];
    c.centerX1 = cs[i].center[0]; c.centerY = cs[i].center[1]; c.centerX2 = cs[i].center[2]; c.centerY2 = cs[i].center[3];
    c.vLimit = cs[i].v_limit;
    if (c.type == MCS_LANE_EXIT) e.hasExitLanes = true;
    e.corr.push_back(c);
    if (order.emplace(c.id, i).second) e.corrIdx.emplace(c.id, i);
  }
  for (auto& kv : order) e.corrOrder.push_back(kv.second);
  // sorted copy, ascending center.p1.y (comparator so:0x1006f0), then neighbours by scanning (so:0x10a780-0x10a839)
  std::vector<int> sorted(n);
  for (int i = 0; i < n; ++i) sorted[i] = i;
  std::sort(sorted.begin(), sorted.end(), [&](int a, int b) { return e.corr[a].centerY < e.corr[b].centerY; });
  for (int ci : e.corrOrder) {
    Corr& c = e.corr[ci];
    for (int k = 0; k < n; ++k)
      if (e.corr[sorted[k]].centerY > c.centerY) { c.hasLeft = true; c.leftId = e.corr[sorted[k]].id; break; }
    for (int k = n - 1; k >= 0; --k)
      if (c.centerY > e.corr[sorted[k]].centerY) { c.hasRight = true; c.rightId = e.corr[sorted[k]].id; break; }
  }
  for (int ci : e.corrOrder)
    if (!e.corr[ci].hasLeft) e.leftmostLaneId = e.corr[ci].id;  // so:0x10a840-0x10a856
}

// Agent::Agent(id,x,y,vx,vy,vd,a,l,w,IdmParam,RSSParam,YieldingParam,behavior)  so:0xf8380-0xf8a58
Agent make_agent(const mcs_agent& in, Rng& rng) {
  Agent a;
  a.id = in.id; a.beh = a.beh0 = in.behavior;
  a.x = in.x; a.y = in.y; a.v = in.v; a.vy = in.vy; a.vd = in.vd; a.a = in.a; a.l = in.l; a.w = in.w;
  a.idm = {in.idm[0], in.idm[1], in.idm[2], in.idm[3], in.idm[4], in.idm[5]};
  a.rss = {in.rss[0], in.rss[1], in.rss[2], in.rss[3], in.rss[4], in.rss[5]};
  std::memcpy(a.yp, in.yp, sizeof a.yp);
  a.states.push_back({a.x, a.y, a.v, a.a});
  a.idmY = a.idm;
  a.rssR = a.rss; a.rssR.aMin = -10.0; a.rssR.aMax = -1.0;     // so:0xf88e7-0xf8962
  if (7.0 > a.l) {                                              // so:0xf8941-0xf89db
    a.idmY.accMin = 1.0 + a.idm.accMin;
    a.idmY.accMax = a.idm.accMax - 0.8;
  }
  a.aggressive = ((double)rng.next() / 2147483647.0) > 0.8;     // so:0xf89e6-0xf8a0a
  // fields the Python side may poke after construction (behaviors.py:77-79)
  a.commit = in.commit; a.keepStep = in.commit_keep_lane_step; a.targetLaneId = in.target_lane_id;
  return a;
}

// lane-change prior from the lateral offset, computed at an agent's first lane match  so:0x10bb97-0x10bd8a
void lane_change_prior(Agent& a, const Corr& c) {
  double d = ((a.y - c.centerY) * 0.333 + 0.6805 * a.vy) - 0.01;
  auto g = [](double t, double w) { return EXP(((-t) * t) / w); };
  double A = (0.8 * g(d - 1.04, 0.055) + 0.8 * g(d - 0.35, 0.055)) + g(d - 0.7, 0.07) * 0.65;
  // note: evaluation order in the binary is e1,e2,e3 then A, then e4,e5,e6 then C, then e7
  double C = (0.75 * g(1.04 + d, 0.055) + 0.75 * g(0.35 + d, 0.055)) + g(0.7 + d, 0.07) * 0.6;
  double B = g(d, 0.055) * 2.5;
  double S = (A + B) + C;
  a.prior[0] = A / S; a.prior[1] = B / S; a.prior[2] = C / S;
}

// Environment::matchAgentsOnCorridor  so:0x10b810-0x10c6b0
void match_agents(Env& e) {
  bool changed = false;
  for (int idx : e.visit) {
    Agent& a = e.agents[idx];
    bool first = a.lanes.empty();
    bool search = true;
    if (!first) {
      Corr* c = e.corrById(a.lane());
      if (!c) return;
      if (c->leftY >= a.y && a.y >= c->rightY) search = false;   // so:0x10b8ce-0x10b8e4
    }
    if (search) {
      for (int ci : e.corrOrder) {                                // so:0x10b8ea-0x10ba3c / 0x10bab8-0x10bdac
        const Corr& c = e.corr[ci];
        if (!(c.leftY >= a.y && a.y >= c.rightY)) continue;
        a.lanes.push_back(c.id);
        if (c.type == MCS_LANE_MAIN && a.beh == MCS_BEH_MERGING) a.beh = MCS_BEH_IDM_LC;   // so:0x10c3a8 / 0x10c4b8
        if (c.type == MCS_LANE_EXIT && a.beh == MCS_BEH_EXIT) a.beh = MCS_BEH_IDM;          // so:0x10c268 / 0x10c430
        if (first) lane_change_prior(a, c);
        break;
      }
      changed = true;                                             // so:0x10b9c8 (also when nothing matched)
    }
    if (a.isEgo) {                                                // so:0x10b9ce-0x10baad
      e.ego = idx;
      if (!a.reached && !a.lanes.empty()) a.reached = (a.lane() == a.targetLaneId);
    }
  }
  if (changed) {                                                  // so:0x10be30-0x10c1a5
    e.laneVec.clear();
    for (int ci : e.corrOrder) e.laneVec[e.corr[ci].id];
    for (int idx : e.visit) {
      Agent& a = e.agents[idx];
      if (a.lanes.empty()) continue;
      auto it = e.laneVec.find(a.lane());
      if (it == e.laneVec.end()) { e.error = true; return; }
      it->second.push_back(idx);
    }
    for (int ci : e.corrOrder) {
      auto& v = e.laneVec[e.corr[ci].id];
      std::sort(v.begin(), v.end(), [&](int p, int q) { return e.agents[p].x < e.agents[q].x; });  // libstdc++ introsort, as the binary
    }
  }
  if (e.ego >= 0) {                                               // so:0x10b9eb-0x10be28, 0x10c320-0x10c39d
    Agent& ego = e.agents[e.ego];
    if (ego.targetGapId >= 0) {
      auto git = e.agentIdx.find(ego.targetGapId);
      if (git == e.agentIdx.end()) { e.error = true; return; }
      Agent& g = e.agents[git->second];
      if (g.lanes.empty()) { e.error = true; return; }
      if (g.lane() != ego.targetLaneId) {
        auto it = e.laneVec.find(ego.targetLaneId);
        if (it == e.laneVec.end()) { e.error = true; return; }
        int found = ego.targetGapId;
        for (int k = (int)it->second.size() - 1; k >= 0; --k) {
          const Agent& o = e.agents[it->second[k]];
          if (g.x > o.x) { found = o.id; ego.targetGapId = found; break; }
        }
        if (g.id == found) ego.targetGapId = -1;
      }
    }
  }
}

// Environment::Environment(map, agents)  so:0x10c6c0 (agent insertion 0x10c8b0-0x10c9be, first match 0x10c9ef)
void build_env(Env& e, const mcs_corridor* cs, int nc, const mcs_agent* as, int na, uint64_t key) {
  e.rng.key = key; e.rng.ctr = 0;
  build_map(e, cs, nc);
  std::unordered_map<int, int> order;  // replays agents.emplace(id, ptr): same bucket growth, same iteration order
  e.agents.reserve(na);
  for (int i = 0; i < na; ++i) {
    e.agents.push_back(make_agent(as[i], e.rng));
    if (order.emplace(as[i].id, i).second) e.agentIdx.emplace(as[i].id, i);
  }
  for (auto& kv : order) e.visit.push_back(kv.second);
  match_agents(e);
}

// Environment::setEgoAgentIdAndDrivingBehavior  so:0xf3f20-0xf42b3
bool set_ego(Env& e, int egoId, int action, int gapId) {
  auto it = e.agentIdx.find(egoId);
  if (it == e.agentIdx.end()) return false;
  Agent& a = e.agents[it->second];
  e.ego = it->second;
  a.isEgo = true; a.egoCmd = action; a.reached = false; a.targetGapId = gapId;
  e.egoId = egoId; e.targetGapId = gapId;
  if (a.lanes.empty()) return false;
  Corr* c = e.corrById(a.lane());
  if (!c) return false;
  switch (action) {
    case MCS_ACT_LEFT: if (c->hasLeft) a.targetLaneId = c->leftId; return c->hasLeft;
    case MCS_ACT_RIGHT: if (c->hasRight) a.targetLaneId = c->rightId; return c->hasRight;
    case MCS_ACT_KEEP_LANE: case MCS_ACT_ACC: case MCS_ACT_DCC: a.targetLaneId = a.lane(); return true;
    default: return false;
  }
}

// Environment::frontAgent  so:0xe2670 (std::upper_bound on the possibly stale-sorted lane vector)
int front_agent(Env& e, int idx) {
  const Agent& a = e.agents[idx];
  if (a.lanes.empty()) return -1;
  auto it = e.laneVec.find(a.lane());
  if (it == e.laneVec.end()) { e.error = true; return -1; }
  const auto& v = it->second;
  long first = 0, len = (long)v.size();
  while (len > 0) {
    long half = len >> 1, mid = first + half;
    if (e.agents[v[mid]].x > a.x) len = half;
    else { first = mid + 1; len = len - half - 1; }
  }
  return first == (long)v.size() ? -1 : v[first];
}

// reverse search used by followingAgent so:0xe2a81-0xe2ae3 and neighborAgents so:0xe871e-0xe8770:
// position `first` such that v[first-1] is the answer (0 = none)
long reverse_bound(const Env& e, const std::vector<int>& v, double x) {
  long first = (long)v.size(), len = (long)v.size();
  while (len > 0) {
    long half = len >> 1, mid = first - half;
    if (x > e.agents[v[mid - 1]].x) len = half;
    else { first = mid - 1; len = len - half - 1; }
  }
  return first;
}

// Environment::followingAgent  so:0xe2920
int following_agent(Env& e, int idx) {
  const Agent& a = e.agents[idx];
  if (a.lanes.empty()) return -1;
  auto it = e.laneVec.find(a.lane());
  if (it == e.laneVec.end()) { e.error = true; return -1; }
  long f = reverse_bound(e, it->second, a.x);
  return f == 0 ? -1 : it->second[f - 1];
}

// one side of Environment::neighborAgents so:0xe8608-0xe87bb: out = [BB, B, F, FF] (Neighbor::possibleLeft/Right order)
void side_neighbors(Env& e, int idx, int laneId, int out[4]) {
  out[0] = out[1] = out[2] = out[3] = -1;
  auto it = e.laneVec.find(laneId);
  if (it == e.laneVec.end()) { e.error = true; return; }
  const auto& v = it->second;
  const double x = e.agents[idx].x;
  long first = 0, len = (long)v.size();
  while (len > 0) {                       // std::lower_bound: first element with x_elem >= x
    long half = len >> 1, mid = first + half;
    if (e.agents[v[mid]].x >= x) len = half;
    else { first = mid + 1; len = len - half - 1; }
  }
  if (first < (long)v.size()) {
    out[2] = v[first];
    if (first + 1 < (long)v.size()) out[3] = v[first + 1];
  }
  long r = reverse_bound(e, v, x);
  if (r > 0) {
    out[1] = v[r - 1];
    if (r > 1) out[0] = v[r - 2];
  }
}

// idmAccFrontBack(IdmParam, vector<AgentPtr> fronts, AgentPtr back, egoX, egoV, egoL, vd)  so:0xebc30-0xebfc8
double idm_acc(Env& e, const Idm& p, const int* fronts, int nf, int back, double egoX, double egoV, double egoL, double vd) {
  double freeAcc = (1.0 - POW4(egoV / vd)) * p.accMax;
  bool any = false;
  double d = 0.0;
  for (int k = 0; k < nf; ++k) {
    if (fronts[k] < 0) continue;
    const Agent& f = e.agents[fronts[k]];
    double gap = (f.x - egoX) - (egoL + f.l) * 0.5;
    double term;
    if (gap > 0.0) {
      double dDes = (egoV * p.thw + p.minDis) * 0.7;
      double na = -p.accMax;
      double dv = (egoV - f.v) * egoV;
      double s = std::sqrt(na * p.accCom);
      double t = MAXSD(dv / (s + s) + dDes, 0.0);
      t = t / gap;
      term = (t * t) * na;
    } else {
      term = -10000000000.0;
    }
    if (!any) { d = term; any = true; }
    else d = MINSD(term, d);
  }
  double sum = any ? d + 0.0 : 0.0;
  if (back >= 0) {
    const Agent& b = e.agents[back];
    double gap = (b.x - egoX) - (egoL + b.l) * 0.5;
    if (0.0 > gap) {
      double dDes = (b.v * p.thw + p.minDis) * 0.7;
      double dv = (b.v - egoV) * egoV;
      double s = std::sqrt((-p.accMax) * p.accCom);
      double t = MAXSD(dv / (s + s) + dDes, 0.0);
      t = t / gap;
      sum = sum + (t * t) * p.accMax;
    } else {
      sum = sum + 10000000000.0;
    }
  }
  double r = MINSD(sum + freeAcc, p.accMax);
  return MAXSD(p.accMin, r);
}

// lateral centring  so:0x10441b-0x10447f (same code at 0xf61af-0xf61fb)
double centering_vy(const Agent& a, const Corr& c) {
  double dy = a.y - c.centerY;
  double ady = std::fabs(dy);
  double m = 0.1;
  if (!(0.1 > ady)) m = MINSD(1.3, ady);
  return ((-dy) / (ady + 1e-10)) * m;
}

// RSS emergency override  so:0x104988-0x104a06 (same code at 0xf6210-0xf629d)
bool rss_emergency(const Agent& a, const Agent& f) {
  double gap = (f.x - a.x) - (f.l + a.l) * 0.5;
  double d = (a.v * a.v) / (a.rss.aMin * -2.0) + a.v * a.rss.rTEgo;
  d = d + (f.v * f.v) / (a.rss.aMax + a.rss.aMax);
  d = MAXSD(0.0, d);
  return gap < d;
}
double rss_override(const Agent& a, const Agent& f, double ax) { return rss_emergency(a, f) ? a.rss.aMin : ax; }

void push_thw_ratio(Agent& a, const Env& e, int front) {  // so:0x1042af-0x1042f9, 0x104bc0
  if (front >= 0) {
    const Agent& f = e.agents[front];
    a.thw.push_back((((f.x - a.x) - (f.l + a.l) * 0.5) / a.v) / a.idm.thw);
  } else {
    a.thw.push_back(1.0);
  }
}

// idmBehaviorAndYieldingToMergingAgents(env, agent, addNoise)  so:0x104200-0x104ec8
Action idm_behavior(Env& e, int idx, bool addNoise) {
  Agent& a = e.agents[idx];
  if (a.lanes.empty()) return {0.0, 0.0};
  Corr* cp = e.corrById(a.lane());
  if (!cp) return {0.0, 0.0};
  const Corr c = *cp;
  int front = front_agent(e, idx);
  push_thw_ratio(a, e, front);
  double vd = a.vd;
  if (a.isEgo) vd = MINSD(vd, 1.1 * c.vLimit);                      // so:0x104301-0x104325
  double axFront = idm_acc(e, a.idm, &front, 1, -1, a.x, a.v, a.l, vd);
  double vy = centering_vy(a, c);

  std::vector<int> cands;
  if (c.hasRight) {                                                  // so:0x104469-0x1044d9, 0x104c08-0x104c81
    Corr* rc = e.corrById(c.rightId);
    if (!rc) return {0.0, 0.0};
    if (rc->type == MCS_LANE_MERGING) {
      int nb[4];
      side_neighbors(e, idx, c.rightId, nb);
      for (int k = 0; k < 4; ++k) if (nb[k] >= 0) cands.push_back(nb[k]);   // Neighbor::right(): BB,B,F,FF non-null
    }
  }
  if (e.hasExitLanes && c.hasLeft) {                                 // so:0x1044df-0x104579, 0x104c90-0x104e64
    Corr* lc = e.corrById(c.leftId);
    if (!lc) return {0.0, 0.0};
    if (lc->type == MCS_LANE_MAIN)
      for (int j : e.visit) {
        const Agent& b = e.agents[j];
        if (!b.lanes.empty() && b.lane() == c.leftId && b.beh == MCS_BEH_EXIT) cands.push_back(j);
      }
  }
  std::vector<int> yielding;
  for (int j : cands) {                                              // so:0x1045b0-0x1048f4
    const Agent& o = e.agents[j];
    double p = 0.0;
    double half = (a.l + o.l) * 0.5;
    double dx = o.x - a.x;
    if (!(0.0 > dx + half)) {
      double thw = (dx - half) / MAXSD(1e-10, a.v);
      double rel = (a.v - o.v) / MAXSD(a.v, 0.001);
      double r = thw * a.yp[1] + a.yp[0];
      r = r + rel * a.yp[2];
      r = r + a.states[0].a * a.yp[3];
      double arg;
      if (100.0 > r) arg = (-100.0 > r) ? 100.0 : -r;              // so:0x104b30-0x104b46, 0x104ec0
      else arg = -100.0;
      p = 1.0 / (EXP(arg) + 1.0);
    }
    int rnd = e.rng.next();                                          // so:0x10469c
    bool y0 = (100.0 * p) >= (double)(rnd % 100);
    auto ins = a.intents.emplace(o.id, Intention{p, 0.625, y0});     // no-op when the id is already known
    Intention& in = ins.first->second;
    if (!in.yielding) {                                              // so:0x104862-0x104880
      if ((1.0 - in.p0) * in.thr >= 1.0 - p) in.yielding = true;
    } else {                                                         // so:0x104a98-0x104aaa
      if (in.thr * in.p0 >= p) in.yielding = false;
    }
    if (in.yielding) yielding.push_back(j);
  }
  double axY = idm_acc(e, a.idmY, yielding.data(), (int)yielding.size(), -1, a.x, a.v, a.l, vd);  // so:0x1048fa-0x10493b
  double ax = MINSD(axY, axFront);                                   // so:0x10496a
  if (addNoise) ax = ax + ((double)e.rng.next() / 2147483647.0 - 0.5);   // so:0x104d98-0x104dc5
  if (front >= 0) ax = rss_override(a, e.agents[front], ax);
  return {ax, vy};
}

// rssSafe(optional frontGap, frontV, backGap, backV, egoV, RSSParam)  so:0xb8950-0xb8a5c
bool rss_safe(bool hasF, double fGap, double fV, bool hasB, double bGap, double bV, double v, const Rss& r) {
  if (!hasB && !hasF) return true;
  if (hasF) {
    double d = (v * v) / (r.aMin * -2.0) + v * r.rTEgo;
    d = d + (fV * fV) / (r.aMax + r.aMax);
    d = MAXSD(0.0, d);
    if (d > fGap) return false;
    if (!hasB) return true;
  }
  double d = (bV * bV) / (r.aMin * -2.0) + bV * r.rTOther;
  d = d + (v * v) / (r.aMax + r.aMax);
  d = MAXSD(0.0, d);
  return !(d > -bGap);
}

// laneChangeProbability(p, prior, leftOk, rightOk)  so:0xe19f0-0xe1a86
void lane_change_probability(double p[3], const double q[3], bool leftOk, bool rightOk) {
  double wP = MAXSD(p[2], p[0]) * 1.5 + 0.5;
  double wQ = MAXSD(q[2], q[0]) * 1.5 + 0.5;
  double L = 0.0, R = 0.0;
  if (leftOk) L = p[0] * wP + q[0] * wQ;
  double K = wP * p[1] + wQ * q[1];
  if (rightOk) R = p[2] * wP + q[2] * wQ;
  double S = (L + K) + R;
  p[0] = L / S; p[1] = K / S; p[2] = R / S;
}

struct MobilOut { double ax, vy; int commit, keepStep, targetLane; };

// gaps/speeds of the would-be leader and follower in a neighbour lane as optionals  so:0xf4a19-0xf4ab0
struct SideGaps { bool hasF, hasB; double fGap, fV, bGap, bV; };
SideGaps side_gaps(const Env& e, const Agent& a, int leader, int follower) {
  SideGaps g{false, false, 0, 0, 0, 0};
  if (leader >= 0) {
    const Agent& f = e.agents[leader];
    g.hasF = true; g.fGap = (f.x - a.x) - (f.l + a.l) * 0.5; g.fV = f.v;
  }
  if (follower >= 0) {
    const Agent& b = e.agents[follower];
    g.hasB = true; g.bGap = (b.l + a.l) * 0.5 + (b.x - a.x); g.bV = b.v;
  }
  return g;
}

// laneChangeBehaviorMOBIL(env, agent, idmAction, commit&, keepStep&, targetLane&, inRollout, requireRss)  so:0xf42d0-0xf5bcb
MobilOut mobil(Env& e, int idx, Action act, bool inRollout, bool requireRss) {
  Agent& a = e.agents[idx];
  auto out = [&](double ax, double vy) { return MobilOut{ax, vy, a.commit, a.keepStep, a.targetLaneId}; };
  if (a.commit == MCS_COMMIT_KEEP_LANE && a.keepStep <= 9) {          // so:0xf434c-0xf4360, 0xf4de8
    a.keepStep = a.keepStep + 1;
    return out(act.ax, act.vy);
  }
  Corr* cp = e.corrById(a.lane());
  if (!cp) return out(act.ax, act.vy);
  const Corr c = *cp;
  bool canLeft = false, canRight = false;                             // so:0xf43ac-0xf44c2, 0xf4e40-0xf4eec
  if (c.hasLeft) { Corr* lc = e.corrById(c.leftId); if (!lc) return out(act.ax, act.vy); canLeft = lc->type == MCS_LANE_MAIN; }
  if (c.hasRight) { Corr* rc = e.corrById(c.rightId); if (!rc) return out(act.ax, act.vy); canRight = rc->type == MCS_LANE_MAIN; }
  if (!canLeft && !canRight) return out(act.ax, act.vy);              // so:0xf43db
  const double pf = a.yp[4], dA = a.yp[5], bA = a.yp[6];
  int back = following_agent(e, idx);
  int front = front_agent(e, idx);
  double deltaOld = 0.0;                                              // so:0xf4583-0xf478b, 0xf5240
  if (back >= 0) {
    const Agent& b = e.agents[back];
    int f2[2] = {front, idx};
    double withEgo = idm_acc(e, b.idm, f2, 2, -1, b.x, b.v, b.l, b.vd);
    double without = idm_acc(e, b.idm, &front, 1, -1, b.x, b.v, b.l, b.vd);
    deltaOld = without - withEgo;
  }
  double incL = -10000000000.0, incR = -10000000000.0;
  bool strictL = false, relaxL = false, strictR = false, relaxR = false;
  double axLeft = act.ax, vyLeft = act.vy, axRight = 0.0, vyRight = 0.0;
  int leftLaneId = 0, rightLaneId = 0;
  if (canLeft) {                                                      // so:0xf5259-0xf584e
    const Corr& lc = *e.corrById(c.leftId);
    leftLaneId = lc.id;
    int nb[4];
    side_neighbors(e, idx, c.leftId, nb);
    int leader = nb[2], follower = nb[1];
    double vdCap = MINSD(a.vd, 1.1 * lc.vLimit);
    double accNew = idm_acc(e, a.idm, &leader, 1, follower, a.x, a.v, a.l, vdCap);
    SideGaps g = side_gaps(e, a, leader, follower);
    strictL = rss_safe(g.hasF, g.fGap, g.fV, g.hasB, g.bGap, g.bV, a.v, a.rss);
    relaxL = rss_safe(g.hasF, g.fGap, g.fV, g.hasB, g.bGap, g.bV, a.v, a.rssR);
    axLeft = (act.ax == a.rss.aMin) ? act.ax : accNew;               // so:0xf55b5-0xf55c9, 0xf5970
    vyLeft = MINSD(1.0, 0.17 * a.v);
    double egoGain = accNew - act.ax;
    double dNew = 0.0;
    if (follower >= 0) {
      const Agent& nf = e.agents[follower];
      double without = idm_acc(e, nf.idm, &leader, 1, -1, nf.x, nf.v, nf.l, nf.vd);
      int f2[2] = {leader, idx};
      double withEgo = idm_acc(e, nf.idm, f2, 2, -1, nf.x, nf.v, nf.l, nf.vd);
      dNew = withEgo - without;
    }
    incL = (((dNew + deltaOld) * pf + egoGain) - dA) - bA;
  }
  if (canRight) {                                                     // so:0xf47e4-0xf4dde
    const Corr& rc = *e.corrById(c.rightId);
    rightLaneId = rc.id;
    int nb[4];
    side_neighbors(e, idx, c.rightId, nb);
    int leader = nb[2], follower = nb[1];
    double vdCap = MINSD(a.vd, 1.1 * rc.vLimit);
    double accNew = idm_acc(e, a.idm, &leader, 1, follower, a.x, a.v, a.l, vdCap);
    SideGaps g = side_gaps(e, a, leader, follower);
    strictR = rss_safe(g.hasF, g.fGap, g.fV, g.hasB, g.bGap, g.bV, a.v, a.rss);
    relaxR = rss_safe(g.hasF, g.fGap, g.fV, g.hasB, g.bGap, g.bV, a.v, a.rssR);
    if (act.ax == a.rss.aMin) axLeft = act.ax;                        // so:0xf4b2e-0xf4b45
    axRight = accNew;
    vyRight = MAXSD(-1.0, -0.17 * a.v);
    double egoGain = accNew - act.ax;
    double dNew = 0.0;
    if (follower >= 0) {
      const Agent& nf = e.agents[follower];
      double without = idm_acc(e, nf.idm, &leader, 1, -1, nf.x, nf.v, nf.l, nf.vd);
      int f2[2] = {idx, leader};
      double withEgo = idm_acc(e, nf.idm, f2, 2, -1, nf.x, nf.v, nf.l, nf.vd);
      dNew = withEgo - without;
    }
    incR = (((dNew + deltaOld) * pf + egoGain) - dA) + bA;
  }
  // keep executing a committed lane change  so:0xf4f12-0xf4f68, 0xf5998, 0xf5ad8
  if (a.commit == MCS_COMMIT_LEFT && a.targetLaneId != c.id && relaxL) return out(axLeft, vyLeft);
  if (a.commit == MCS_COMMIT_RIGHT && a.targetLaneId != c.id && relaxR) return out(axRight, vyRight);
  a.keepStep = 0;                                                     // so:0xf4f76
  double eL = EXP(incL), eR = EXP(incR);                              // so:0xf4f6e-0xf501f
  double denom = (1.0 + eL) + eR;
  double p[3];
  p[0] = eL / denom;
  p[2] = eR / denom;
  p[1] = (1.0 - p[0]) - p[2];
  const double zero[3] = {0.0, 0.0, 0.0};
  if (inRollout) {                                                    // so:0xf58a0-0xf595a, 0xf5a60, 0xf5b71
    if (a.commit == MCS_COMMIT_NOT_DECIDED && a.states.size() <= 1) {
      if (a.aggressive) lane_change_probability(p, a.prior, relaxL, relaxR);
      else {
        bool r = strictR ? true : ((a.prior[2] >= 0.5) && relaxR);
        bool l = strictL ? true : ((a.prior[0] >= 0.5) && relaxL);
        lane_change_probability(p, a.prior, l, r);
      }
    } else if (a.aggressive) lane_change_probability(p, zero, relaxL, relaxR);
    else lane_change_probability(p, zero, strictL, strictR);
  } else if (requireRss) lane_change_probability(p, zero, strictL, strictR);   // so:0xf5858
  else lane_change_probability(p, zero, relaxL, relaxR);                       // so:0xf5053
  double pL = p[0];
  if (a.l > 6.0) {                                                    // so:0xf508f-0xf50ff, 0xf5ac8
    pL = pL * 0.5;
    if (c.hasLeft && c.leftId == e.leftmostLaneId) pL = 0.0;
    double k2 = p[1] + p[1], r2 = p[2] + p[2];
    double S = (k2 + pL) + r2;
    p[0] = pL; p[1] = k2 / S; p[2] = r2 / S;
  }
  double u = (double)e.rng.next() / 2147483647.0;                     // so:0xf510f-0xf5124
  if (pL > u && canLeft) {                                            // so:0xf512c-0xf5139, 0xf5b28
    a.commit = MCS_COMMIT_LEFT; a.targetLaneId = leftLaneId;
    return out(axLeft, vyLeft);
  }
  if (pL + p[2] > u && canRight) {                                    // so:0xf513f-0xf515e, 0xf59e8
    a.commit = MCS_COMMIT_RIGHT; a.targetLaneId = rightLaneId;
    return out(axRight, vyRight);
  }
  a.commit = MCS_COMMIT_KEEP_LANE; a.targetLaneId = -1;               // so:0xf5164-0xf5195
  return out(act.ax, act.vy);
}

// customizedLaneChangeBehavior(env, ego, command)  so:0xf5e30-0xf6a2f
Action customized_lane_change(Env& e, int idx, int cmd) {
  Agent& a = e.agents[idx];
  if (a.lanes.empty()) return {0.0, 0.0};
  Corr* cp = e.corrById(a.lane());
  if (!cp) return {0.0, 0.0};
  const Corr c = *cp;
  int front = front_agent(e, idx);
  push_thw_ratio(a, e, front);
  double vdCap = MINSD(a.vd, 1.1 * c.vLimit);                         // so:0xf5f33-0xf5f51
  bool haveAx = false;
  double ax = 0.0;
  if (cmd == MCS_ACT_KEEP_LANE) { ax = idm_acc(e, a.idm, &front, 1, -1, a.x, a.v, a.l, vdCap); haveAx = true; }
  if (cmd == MCS_ACT_DCC) {                                           // so:0xf6054-0xf6157
    Idm p = a.idm; p.thw = a.idm.thw + a.idm.thw;
    ax = idm_acc(e, p, &front, 1, -1, a.x, a.v, a.l, 0.8 * vdCap); haveAx = true;
  }
  if (cmd == MCS_ACT_ACC) {                                           // so:0xf6430-0xf652c
    Idm p = a.idm; p.thw = 0.9;
    ax = idm_acc(e, p, &front, 1, -1, a.x, a.v, a.l, 1.1 * vdCap); haveAx = true;
  }
  double vy = centering_vy(a, c);
  // RSS emergency: for left/right the binary reads an uninitialised stack slot here (so:0xf6263);
  // only the aMin outcome of the select is ever observable, tracked as `emergency`.
  bool emergency = false;
  if (front >= 0) {
    if (haveAx) ax = rss_override(a, e.agents[front], ax);
    else emergency = rss_emergency(a, e.agents[front]);
  }
  if (haveAx) return {ax, vy};
  int sideLane; double sign;
  if (cmd == MCS_ACT_LEFT) {                                          // so:0xf63be-0xf6421, 0xf65d6-0xf6631, 0xf6976
    if (!c.hasLeft) return {emergency ? a.rss.aMin : 0.0, vy};
    Corr* lc = e.corrById(c.leftId);
    if (!lc || lc->type != MCS_LANE_MAIN) return {emergency ? a.rss.aMin : 0.0, vy};
    sideLane = c.leftId; sign = 1.0;
  } else {                                                            // so:0xf656d-0xf65d1, 0xf6636
    if (!c.hasRight) return {emergency ? a.rss.aMin : 0.0, vy};
    Corr* rc = e.corrById(c.rightId);
    if (!rc || rc->type != MCS_LANE_MAIN) return {emergency ? a.rss.aMin : 0.0, vy};
    sideLane = c.rightId; sign = -1.0;
  }
  int nb[4];
  side_neighbors(e, idx, sideLane, nb);
  int leader = nb[2], follower = nb[1];
  int f2[2] = {leader, front};
  double axT = idm_acc(e, a.idm, f2, 2, follower, a.x, a.v, a.l, vdCap);   // so:0xf6798-0xf67c9
  if (emergency) axT = a.rss.aMin;                                          // so:0xf6802-0xf681d
  SideGaps g = side_gaps(e, a, leader, follower);
  bool safe = rss_safe(g.hasF, g.fGap, g.fV, g.hasB, g.bGap, g.bV, a.v, a.rss);
  if (safe) vy = MINSD(1.0, 0.17 * a.v) * sign;                             // so:0xf6909-0xf692e
  return {axT, vy};
}

// ---------------------------------------------------------------------------------------------
// merging / exit: gap selection and gap execution

// idmAccFrontBack(IdmParam, vector<IdmMatch> fronts, IdmMatch back, v, vd)  so:0xeb8d0-0xebbfd
// (differs from the Agent overload above: accMin where that one uses -accMax)
struct Match { int id; double dis, v; };
double idm_acc_match(const Idm& p, const Match* fronts, int nf, const Match& back, double v, double vd) {
  double freeAcc = (1.0 - POW4(v / vd)) * p.accMax;
  bool any = false;
  double d = 0.0;
  for (int k = 0; k < nf; ++k) {
    const Match& m = fronts[k];
    if (m.id < 0) continue;
    double term;
    if (m.dis > 0.0) {
      double dv = (v - m.v) * v;
      double dDes = (v * p.thw + p.minDis) * 0.7;
      double s = std::sqrt(p.accMin * p.accCom);
      double t = MAXSD(dv / (s + s) + dDes, 0.0);
      t = t / m.dis;
      term = (t * t) * p.accMin;
    } else {
      term = -10000000000.0;
    }
    if (!any) { d = term; any = true; }
    else d = MINSD(term, d);
  }
  double sum = any ? d + 0.0 : 0.0;
  if (back.id >= 0) {
    if (0.0 > back.dis) {
      double dDes = (back.v * p.thw + p.minDis) * 0.7;
      double dv = (back.v - v) * v;
      double s = std::sqrt(p.accMin * p.accCom);
      double t = MAXSD(dv / (s + s) + dDes, 0.0);
      t = t / back.dis;
      sum = sum + (t * t) * p.accMax;
    } else {
      sum = sum + 10000000000.0;
    }
  }
  double r = MINSD(sum + freeAcc, p.accMax);
  return MAXSD(p.accMin, r);
}

struct Opt { bool has; double v; };

// rssSafeRelax(optional gapFF, vFF, gapF, vF, gapB, vB, egoV, vd, RSSParam, IdmParam)  so:0xb9210-0xb95e4
// two look-ahead iterations of one second each: RSS-style checks with estimated (not worst-case) decelerations,
// then the ego is advanced with the IdmMatch IDM and the gaps are updated.
bool rss_safe_relax(Opt g1, Opt v1o, Opt g3, Opt v3o, Opt g5, Opt v5o, double egoV, double vd, const Rss& r, const Idm& idm) {
  Match back{-1, 0.0, 0.0};
  double v = egoV;
  if (g5.has) back = Match{1, g5.v, v5o.v};
  int id3 = -1, id1 = -1;
  double gap3 = 0.0, v3 = 0.0, gap1 = 0.0, v1 = 0.0;
  if (g3.has) { gap3 = g3.v; v3 = v3o.v; id3 = 1; }
  if (g1.has) { gap1 = g1.v; v1 = v1o.v; id1 = 1; }
  for (int it = 2;; it = 1) {
    const double aMaxHalf = r.aMax * 0.5;
    double a3 = aMaxHalf;                                  // so:0xb92fc-0xb934f: deceleration the front vehicle needs behind its own leader
    if (id1 == 1) {
      double num = (-0.5 * v3) * v3;
      double den = (v1 * v1) / (r.aMax * -2.0);
      den = den + (gap1 - gap3);
      den = den - v3 * r.rTOther;
      den = MAXSD(den, 0.01);
      a3 = num / den;
      a3 = MAXSD(r.aMax, a3);
      a3 = MINSD(aMaxHalf, a3);
    }
    double aEgo = aMaxHalf;
    if (id3 != -1) {                                       // so:0xb9358-0xb93ce
      double q = (v3 * v3) / (a3 + a3);
      double vr = v * r.rTEgo;
      double d = (v * v) / (-2.0 * r.aMin);
      d = d + vr;
      d = d + q;
      d = MAXSD(0.0, d);
      double num = (v * -0.5) * v;
      double den = (v3 * v3) / (r.aMax * -2.0);
      den = den + gap3;
      den = den - vr;
      den = MAXSD(den, 0.01);
      double ae = num / den;
      ae = MAXSD(r.aMax, ae);
      aEgo = MINSD(aMaxHalf, ae);
      if (d > gap3) return false;
    }
    if (back.id >= 0) {                                    // so:0xb93d8-0xb9425
      double q = (v * v) / (aEgo + aEgo);
      double d = (back.v * back.v) / (-2.0 * r.aMin);
      d = d + back.v * r.rTOther;
      d = d + q;
      d = MAXSD(0.0, d);
      if (d > -back.dis) return false;
    }
    Match fm{id3, gap3, v3};                               // so:0xb942b-0xb94a6
    double acc = idm_acc_match(idm, &fm, 1, back, v, vd);
    double s = acc * 0.5 + v;
    v = MAXSD(0.0, v + acc);
    s = MAXSD(s, 0.0);
    if (id1 != -1) gap1 = (v1 + gap1) - s;
    if (id3 != -1) gap3 = (v3 + gap3) - s;
    if (back.id >= 0) {                                    // so:0xb952c-0xb955b
      double nb = ((back.v + back.dis) - r.dSoft * 0.5) - s;
      back.dis = nb;
      back.v = MAXSD(back.v - r.dSoft, 0.0);
    }
    if (it == 1) return true;
  }
}

// timeToReachGap(xBack, vBack, xFront, vFront, xEgo, vEgo, vTarget)  so:0xe1630-0xe19dc
double time_to_reach_gap(double xB, double vB, double xF, double vF, double xE, double vE, double vT) {
  const double a = (vT - vE) / 5.0;
  if (xB > xE) {                                           // gap ahead: catch up with acceleration clamp(a, 0.5, 2)
    const double acc = (a > 2.0) ? 2.0 : MAXSD(0.5, a);
    const double dv = vB - vE;
    const double dx = xB - xE;
    const double inv = 1.0 / acc;
    const double disc = dx * (acc + acc) + dv * dv;
    const double sq = std::sqrt(disc);
    const double r1 = (dv + sq) * inv;
    const double r2 = (dv - sq) * inv;
    return MAXSD(r1, r2);
  }
  if (xE > xF) {                                           // gap behind: fall back with 2 m/s^2
    const double m = vE - vF;
    const double dv = vF - vE;
    const double disc = dv * dv - (xF - xE) * 4.0;
    const double sq = std::sqrt(disc);
    const double r1 = (m + sq) * 0.5;
    const double r2 = (m - sq) * 0.5;
    return MAXSD(r1, r2);
  }
  // alongside the gap: time to reach half the speed of the closer gap vehicle (sic)
  const double X = (std::fabs(vB - vE) < std::fabs(vF - vE)) ? vB : vF;
  double acc = a;
  if (a > 2.0) acc = 2.0;
  else if (0.5 > a) acc = 0.5;
  const double h = X * 0.5;
  if (vE > h) return 0.0;
  const double t = h - vE;
  return (t + t) / acc;
}

// common tail of customizedMergingBehavior (so:0xf7227-0xf72e8, 0xf7610) and mergingBehaviorClosestGap
// (so:0xff0df-0xff10e, 0xff5a8, 0xff5e0): drift towards the target lane while the vehicle's edge is more than 0.5 m
// inside its own lane; past that point only when the gap is safe and the target lane has started, otherwise back off.
Action merge_tail(const Agent& a, double sign, bool far, bool lateral, double targetStartX, double ax) {
  const double m = a.v * 0.17;
  const bool toward = far || (lateral && a.x >= targetStartX);
  double vy;
  if (toward) vy = (m > 1.0) ? sign : m * sign;
  else vy = ((m > 1.0) ? -1.0 : -m) * sign;
  return {ax, vy};
}

bool far_from_boundary(const Agent& a, const Corr& c, bool left) {
  if (left) return -0.5 > (a.w * 0.5 + a.y) - c.leftY;     // so:0xf727c-0xf7294 / 0xfe956-0xfe97a
  return (a.y - a.w * 0.5) - c.rightY > 0.5;               // so:0xf7640-0xf7654 / 0xff267-0xff283
}

// Neighbor::left() so:0xeae70 / right(): the side's [BB, B, F, FF] without the nulls
int side_vector(Env& e, int idx, int laneId, int out[4]) {
  int nb[4], n = 0;
  side_neighbors(e, idx, laneId, nb);
  for (int k = 0; k < 4; ++k) if (nb[k] >= 0) out[n++] = nb[k];
  return n;
}

Opt front_gap(const Env& e, const Agent& a, int j) {   // (o.x - a.x) - (a.l + o.l)/2
  if (j < 0) return Opt{false, 0.0};
  const Agent& o = e.agents[j];
  return Opt{true, (o.x - a.x) - (o.l + a.l) * 0.5};
}
Opt back_gap(const Env& e, const Agent& a, int j) {    // (a.l + o.l)/2 + (o.x - a.x)
  if (j < 0) return Opt{false, 0.0};
  const Agent& o = e.agents[j];
  return Opt{true, (o.l + a.l) * 0.5 + (o.x - a.x)};
}
Opt speed_of(const Env& e, int j) { return j < 0 ? Opt{false, 0.0} : Opt{true, e.agents[j].v}; }

// shared by both: RSS emergency towards the own leader, room left on the own lane, choice of ax  so:0xf7155-0xf7224 / 0xff02e-0xff0df
Action merge_finish(Env& e, Agent& a, const Corr& c, const Corr& tc, int front, bool left, bool safe, double axGap) {
  double brake = (a.v * a.v) / (a.rss.aMin * -2.0) + a.v * a.rss.rTEgo;
  bool restrict_ = false;
  if (front >= 0) {
    const Agent& f = e.agents[front];
    double gap = (f.x - a.x) - (f.l + a.l) * 0.5;
    double d = (f.v * f.v) / (a.rss.aMax + a.rss.aMax) + brake;
    d = MAXSD(0.0, d);
    if (d > gap) restrict_ = true;
  }
  if (!restrict_) {
    bool room = (c.centerX2 - a.x) - 15.0 >= brake;
    if (!room && !safe) restrict_ = true;
  }
  const double ax = restrict_ ? a.rss.aMin : axGap;
  const bool lateral = restrict_ ? false : safe;
  return merge_tail(a, left ? 1.0 : -1.0, far_from_boundary(a, c, left), lateral, tc.centerX1, ax);
}

// customizedMergingBehavior(env, ego, direction, gapId)  so:0xf6ba0-0xf7788
Action customized_merging(Env& e, int idx, int cmd, int gapId) {
  Agent& a = e.agents[idx];
  if (a.lanes.empty()) return {0.0, 0.0};                             // so:0xf73e0 "is not matched!"
  Corr* cp = e.corrById(a.lane());
  if (!cp) return {0.0, 0.0};
  const Corr c = *cp;
  int front = front_agent(e, idx);
  push_thw_ratio(a, e, front);
  const bool left = cmd == MCS_ACT_LEFT;                              // anything but "left" is "right" (so:0xf6cbe)
  if (!(left ? c.hasLeft : c.hasRight)) { e.error = true; return {0.0, 0.0}; }   // binary prints, then reads a missing corridor
  const int side = left ? c.leftId : c.rightId;
  int g = -1, gf = -1, gff = -1;
  auto git = (gapId == -1) ? e.agentIdx.end() : e.agentIdx.find(gapId);
  if (git != e.agentIdx.end()) {                                      // so:0xf6d46-0xf6e59
    g = git->second;
    gf = front_agent(e, g);
    if (gf >= 0) gff = front_agent(e, gf);
  } else {                                                            // -1 or unknown id: "go to last gap" so:0xf7488-0xf754f, 0xf7660
    int v[4];
    int n = side_vector(e, idx, side, v);
    if (n >= 1) gf = v[0];
    if (n >= 2) gff = v[1];
  }
  if (e.error) return {0.0, 0.0};
  const double vdCap = MINSD(a.vd, 1.1 * c.vLimit);
  bool safe = rss_safe_relax(front_gap(e, a, gff), speed_of(e, gff), front_gap(e, a, gf), speed_of(e, gf),
                             back_gap(e, a, g), speed_of(e, g), a.v, vdCap, a.rss, a.idm);
  Corr* tcp = e.corrById(side);
  if (!tcp) return {0.0, 0.0};
  const Corr tc = *tcp;
  int f3[3] = {gf, gff, front};                                       // so:0xf707a-0xf70d9
  double axGap = idm_acc(e, a.idm, f3, 3, g, a.x, a.v, a.l, vdCap);
  return merge_finish(e, a, c, tc, front, left, safe, axGap);
}

// mergingBehaviorClosestGap(env, agent, direction)  so:0xfe660-0xff6cf
Action merging_closest_gap(Env& e, int idx, int direction) {
  Agent& a = e.agents[idx];
  if (a.lanes.empty()) return {0.0, 0.0};
  Corr* cp = e.corrById(a.lane());
  if (!cp) return {0.0, 0.0};
  const Corr c = *cp;
  int front = front_agent(e, idx);
  push_thw_ratio(a, e, front);
  double vd = a.vd;
  if (a.isEgo) vd = MINSD(vd, 1.1 * c.vLimit);                        // so:0xfe757-0xfe77e
  double axFront = idm_acc(e, a.idm, &front, 1, -1, a.x, a.v, a.l, vd);
  const bool left = direction == MCS_ACT_LEFT;
  if (!(left ? c.hasLeft : c.hasRight)) { e.error = true; return {0.0, 0.0}; }
  const int side = left ? c.leftId : c.rightId;
  Corr* tcp = e.corrById(side);
  if (!tcp) return {0.0, 0.0};
  const Corr tc = *tcp;
  const bool far = far_from_boundary(a, c, left);
  int v[4];
  const int n = side_vector(e, idx, side, v);
  if (e.error) return {0.0, 0.0};
  if (n == 0) {                                                       // so:0xff2bf-0xff2f6, 0xff540
    if (a.x >= tc.centerX1 || far) return merge_tail(a, left ? 1.0 : -1.0, true, false, 0.0, axFront);
    return {axFront, 0.0};
  }
  double cost[5];                                                     // std::map<int,double>: gap k lies behind v[k]
  {
    const Agent& first = e.agents[v[0]];
    const Agent& last = e.agents[v[n - 1]];
    cost[0] = time_to_reach_gap(-10000000000.0, 0.0, first.x - (first.l + a.l) * 0.5, first.v, a.x, a.v, vd);
    cost[n] = time_to_reach_gap((last.l + a.l) * 0.5 + last.x, last.v, 10000000000.0, 100.0, a.x, a.v, vd);
    for (int i = 0; i + 1 < n; ++i) {
      const Agent& B = e.agents[v[i]];
      const Agent& F = e.agents[v[i + 1]];
      cost[i + 1] = time_to_reach_gap((a.l + B.l) * 0.5 + B.x, B.v, F.x - (a.l + F.l) * 0.5, F.v, a.x, a.v, vd);
    }
  }
  int k = 0;
  for (int j = 1; j <= n; ++j) if (cost[k] > cost[j]) k = j;          // so:0xfed88-0xfedaf
  int gf = -1, gff = -1, g = -1;
  double axGap;
  if (k == 0) {                                                       // so:0xff370-0xff4c8
    gf = v[0]; if (n >= 2) gff = v[1];
    axGap = idm_acc(e, a.idm, &gf, 1, -1, a.x, a.v, a.l, vd);
  } else if (k == n) {                                                // so:0xff610-0xff69c
    g = v[n - 1];
    axGap = idm_acc(e, a.idm, nullptr, 0, g, a.x, a.v, a.l, vd);
  } else {                                                            // so:0xfee31-0xfefd4
    gf = v[k]; g = v[k - 1]; if (k + 1 < n) gff = v[k + 1];
    axGap = idm_acc(e, a.idm, &gf, 1, g, a.x, a.v, a.l, vd);
  }
  bool safe = rss_safe_relax(front_gap(e, a, gff), speed_of(e, gff), front_gap(e, a, gf), speed_of(e, gf),
                             back_gap(e, a, g), speed_of(e, g), a.v, vd, a.rss, a.idm);
  return merge_finish(e, a, c, tc, front, left, safe, axGap);
}


// Simulation::run  so:0x10e790-0x10f8ed
bool run_rollout(Env& e, const mcs_sim_param& sp, mcs_status& st) {
  std::memset(&st, 0, sizeof st);
  int finishStep = sp.steps;
  bool finish = false;
  int step = 0;
  const size_t n = e.visit.size();
  std::vector<Action> acts(n);
  while (sp.steps > step) {
    for (size_t k = 0; k < n; ++k) {                                  // plan  so:0x10e808-0x10ea36
      int idx = e.visit[k];
      Agent& a = e.agents[idx];
      Action act{0.0, 0.0};
      if (a.isEgo) {
        if (a.beh == MCS_BEH_IDM_LC || a.beh == MCS_BEH_IDM) {
          bool lr = a.egoCmd == MCS_ACT_LEFT || a.egoCmd == MCS_ACT_RIGHT;
          if (lr && a.reached) act = idm_behavior(e, idx, false);
          else act = customized_lane_change(e, idx, a.egoCmd);
        } else if (a.beh == MCS_BEH_MERGING || a.beh == MCS_BEH_EXIT) {  // so:0x10f500-0x10f54a
          if (a.reached) act = idm_behavior(e, idx, false);
          else act = customized_merging(e, idx, a.egoCmd, a.targetGapId);
        } else { e.error = true; }
      } else if (a.beh == MCS_BEH_IDM_LC) {
        Action i = idm_behavior(e, idx, true);
        MobilOut m = mobil(e, idx, i, true, false);
        act = {m.ax, m.vy};
      } else if (a.beh == MCS_BEH_IDM) {
        act = idm_behavior(e, idx, true);
      } else if (a.beh == MCS_BEH_MERGING) {
        act = merging_closest_gap(e, idx, MCS_ACT_LEFT);
      } else {
        e.error = true;  // the binary pushes no action for such an agent and then indexes past the vector
      }
      if (e.error) return false;
      acts[k] = act;
    }
    for (size_t k = 0; k < n; ++k) {                                  // integrate  so:0x10ea36-0x10eb80, 0x10f4e8
      Agent& a = e.agents[e.visit[k]];
      double ax = acts[k].ax, vy = acts[k].vy;
      double v1 = ax * sp.dt + a.v;
      double acc;
      if (0.0 > v1) {
        a.v = 0.0;
        a.x = 0.0 * sp.dt + a.x;
        a.vy = vy;
        a.y = vy * sp.dt + a.y;
        acc = 0.0; v1 = 0.0;
      } else {
        a.v = v1;
        a.x = v1 * sp.dt + a.x;
        a.vy = vy;
        a.y = vy * sp.dt + a.y;
        acc = (v1 == 0.0) ? 0.0 : ax;
      }
      a.a = acc;
      a.states.push_back({a.x, a.y, v1, acc});
    }
    match_agents(e);                                                  // so:0x10edb3
    if (e.error) return false;
    e.stepCount++;
    if (e.egoId >= 0 && e.ego >= 0 && e.agents[e.ego].reached) {      // so:0x10edd0-0x10ee03
      finish = true;
      finishStep = std::min(finishStep, step);
      if (sp.min_steps <= step) break;
    }
    step++;
  }
  st.executedSteps = (int)e.agents[e.visit[0]].states.size() - 1;
  finishStep = std::max(finishStep, (int)(1.0 / sp.dt));              // so:0x10ee20-0x10ee6e
  // objects  so:0x10ee71-0x10f04b
  double utilObj = 1.0, comfObj = 1.0;
  bool anyObj = false;
  for (int idx : e.visit) {
    const Agent& a = e.agents[idx];
    if (a.isEgo) continue;
    double sumAbs = 0.0, sumV = 0.0;
    for (const State& s : a.states) sumAbs = sumAbs + std::fabs(s.a);
    for (const State& s : a.states) sumV = sumV + s.v;
    double cnt = (double)a.states.size();
    double meanV = sumV / cnt;
    double comf = (sumAbs / cnt) / a.rss.aMin + 1.0;
    double util = (meanV >= a.vd) ? 1.0 : meanV / a.vd;
    if (!anyObj) { utilObj = util; comfObj = comf; anyObj = true; }
    else { utilObj = MINSD(util, utilObj); comfObj = MINSD(comf, comfObj); }
  }
  st.utilityObjects = utilObj; st.comfortObjects = comfObj;
  st.finish = finish; st.finishStep = finishStep;
  if (e.egoId < 0 || e.ego < 0) { st.fallBack = 0; st.utility = 0.0; st.comfort = 0.0; return true; }   // so:0x10f620
  const Agent& ego = e.agents[e.ego];
  if (ego.beh0 == MCS_BEH_MERGING || ego.beh0 == MCS_BEH_EXIT) {      // so:0x10f075-0x10f0af, 0x10f780
    st.fallBack = 2.0 > ego.v;
  } else {                                                            // so:0x10f797-0x10f853
    st.fallBack = 0;
    for (size_t t = 0; t + 1 < ego.states.size(); ++t)
      if (ego.rss.aMin == ego.states[t].a && ego.rss.aMin == ego.states[t + 1].a) { st.fallBack = 1; break; }
  }
  double sumU = 0.0;                                                  // so:0x10f0c4-0x10f147
  for (const State& s : ego.states) {
    double r = s.v / ego.vd;
    double u = r;
    if (!(ego.vd >= s.v)) u = MAXSD(0.9, 1.0 - (r - 1.0));
    sumU = sumU + u;
  }
  double util = sumU / (double)ego.states.size();
  double thwTerm = 0.0;                                               // so:0x10f14b-0x10f19f
  if (!ego.thw.empty()) {
    double acc = 0.0;
    for (double t : ego.thw)
      if (0.9 > t && t > 0.0) acc = (t + acc) - 0.9;
    thwTerm = acc / (double)ego.thw.size();
  }
  st.utility = util + thwTerm;
  std::vector<double> c;                                              // so:0x10f1c3-0x10f49a
  for (const State& s : ego.states) c.push_back(s.a > 0.0 ? std::fabs(s.a) * 0.5 : std::fabs(s.a));
  std::sort(c.begin(), c.end(), [](double p, double q) { return p > q; });
  int k = (int)((double)c.size() * 0.2);
  double sumC = 0.0;
  for (int i = 0; i < k; ++i) sumC = sumC + (c[i] / ego.rss.aMin + 1.0);
  st.comfort = sumC / (double)k;
  return true;
}


int one_rollout(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, const mcs_candidate& cand,
                const mcs_sim_param& sp, uint64_t key, mcs_status& st, double* states, int32_t* nStates) {
  Env e;
  build_env(e, cs, nc, as, na, key);
  if (e.error) return MCS_ERR_SCENE;
  bool ok = set_ego(e, cand.ego_id, cand.action, cand.gap_id);
  bool ran = run_rollout(e, sp, st);
  if (!ran) return MCS_ERR_SCENE;
  st.ok = ok;
  st.draws = (uint32_t)e.rng.ctr;
  if (states) {
    const int stride = sp.steps + 1;
    for (int i = 0; i < na; ++i) {
      const auto& h = e.agents[i].states;
      for (size_t t = 0; t < h.size() && (int)t < stride; ++t) {
        double* o = states + ((size_t)i * stride + t) * 4;
        o[0] = h[t].x; o[1] = h[t].y; o[2] = h[t].v; o[3] = h[t].a;
      }
    }
  }
  if (nStates) *nStates = (int32_t)e.agents[0].states.size();
  return MCS_OK;
}

}  // namespace

extern "C" {

void orc_set_math_mode(int mode) { g_math_mode = mode; }

int orc_rollouts(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, const mcs_candidate* cand,
                 const mcs_sim_param* sp, uint64_t seed, int64_t first, int64_t count, mcs_status* status_out,
                 double* states_out, int32_t* n_states_out) {
  if (!cs || !as || !cand || !sp || !status_out || nc <= 0 || na <= 0) return MCS_ERR_ARG;
  const size_t per = (size_t)na * (sp->steps + 1) * 4;
  for (int64_t i = 0; i < count; ++i) {
    int rc = one_rollout(cs, nc, as, na, *cand, *sp, mcs_rollout_key(seed, (uint64_t)(first + i)), status_out[i],
                         states_out ? states_out + per * i : nullptr, n_states_out ? n_states_out + i : nullptr);
    if (rc != MCS_OK) return rc;
  }
  return MCS_OK;
}

int orc_rollouts_mt(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, const mcs_candidate* cand,
                    const mcs_sim_param* sp, uint64_t seed, int64_t first, int64_t count, int n_threads,
                    mcs_status* status_out) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  std::vector<int> rcs(n_threads, MCS_OK);
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      int64_t lo = count * t / n_threads, hi = count * (t + 1) / n_threads;
      rcs[t] = orc_rollouts(cs, nc, as, na, cand, sp, seed, first + lo, hi - lo, status_out + lo, nullptr, nullptr);
    });
  for (auto& x : th) x.join();
  for (int rc : rcs) if (rc != MCS_OK) return rc;
  return MCS_OK;
}

// runMTimes so:0xc2a6e-0xc2b1f per group, runMultiThreads so:0xbca08-0xbca92 across groups
int orc_reduce_features(const mcs_status* s, int groups, int m, int steps, mcs_feature* out) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int nOk = 0;
  for (int g = 0; g < groups; ++g) {
    const mcs_status* p = s + (size_t)g * m;
    if (m <= 0 || !p[0].ok) continue;   // setEgo failed → fromNSims = 0 for the group
    int nFall = 0, nFin = 0, sumStep = 0;
    double u = 0.0, c = 0.0, uo = 0.0, co = 0.0;
    for (int i = 0; i < m; ++i) {
      nFall += p[i].fallBack != 0;
      if (p[i].finish) { nFin++; sumStep += p[i].finishStep; }
      u = u + p[i].utility; c = c + p[i].comfort; uo = uo + p[i].utilityObjects; co = co + p[i].comfortObjects;
    }
    double M = (double)m;
    double f[7];
    f[0] = (double)nFall / M; f[1] = (double)nFin / M; f[2] = u / M; f[3] = c / M;
    f[4] = (double)(sumStep <= 0 ? 1 : sumStep) / (double)(nFin * steps <= 0 ? 1 : nFin * steps);
    f[5] = uo / M; f[6] = co / M;
    for (int k = 0; k < 7; ++k) acc[k] = acc[k] + f[k];
    nOk++;
  }
  double d = (double)nOk;
  out->fallBackRate = acc[0] / d; out->successRate = acc[1] / d; out->utility = acc[2] / d; out->comfort = acc[3] / d;
  out->expectedStepRatio = acc[4] / d; out->utilityObjects = acc[5] / d; out->comfortObjects = acc[6] / d;
  out->fromNSims = nOk * m; out->reserved = 0;
  return MCS_OK;
}

int orc_scene_orders(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, int32_t* agent_order,
                     int32_t* corridor_order, int32_t* left_ids, int32_t* right_ids) {
  Env e;
  build_env(e, cs, nc, as, na, 0);
  for (size_t i = 0; i < e.visit.size(); ++i) agent_order[i] = e.visit[i];
  for (size_t i = 0; i < e.corrOrder.size(); ++i) corridor_order[i] = e.corrOrder[i];
  for (int i = 0; i < nc; ++i) {
    left_ids[i] = e.corr[i].hasLeft ? e.corr[i].leftId : INT32_MIN;
    right_ids[i] = e.corr[i].hasRight ? e.corr[i].rightId : INT32_MIN;
  }
  return e.error ? MCS_ERR_SCENE : MCS_OK;
}

// python_converter::laneChangeBehaviorMOBILWrapper so:0xb9960: MOBIL with (inRollout=false, requireRss)
int orc_decide_mobil(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, int32_t agent_id,
                     const mcs_action* act, int require_rss, uint64_t key, mcs_action_with_type* out) {
  Env e;
  build_env(e, cs, nc, as, na, key);
  auto it = e.agentIdx.find(agent_id);
  if (e.error || it == e.agentIdx.end()) return MCS_ERR_SCENE;
  MobilOut m = mobil(e, it->second, {act->ax, act->vy}, false, require_rss != 0);
  if (e.error) return MCS_ERR_SCENE;
  out->ax = m.ax; out->vy = m.vy; out->commit = m.commit; out->commit_keep_lane_step = m.keepStep;
  out->target_lane_id = m.targetLane; out->reserved = 0;
  return MCS_OK;
}

int orc_decide_fixed_direction(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, int32_t agent_id,
                               int32_t action, uint64_t key, mcs_action* out) {
  Env e;
  build_env(e, cs, nc, as, na, key);
  auto it = e.agentIdx.find(agent_id);
  if (e.error || it == e.agentIdx.end()) return MCS_ERR_SCENE;
  Action a = customized_lane_change(e, it->second, action);
  if (e.error) return MCS_ERR_SCENE;
  out->ax = a.ax; out->vy = a.vy;
  return MCS_OK;
}

int orc_decide_fixed_gap(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, int32_t agent_id,
                         int32_t direction, int32_t gap_id, uint64_t key, mcs_action* out) {
  Env e;
  build_env(e, cs, nc, as, na, key);
  auto it = e.agentIdx.find(agent_id);
  if (e.error || it == e.agentIdx.end()) return MCS_ERR_SCENE;
  Action a = customized_merging(e, it->second, direction, gap_id);
  if (e.error) return MCS_ERR_SCENE;
  out->ax = a.ax; out->vy = a.vy;
  return MCS_OK;
}

int orc_decide_closest_gap(const mcs_corridor* cs, int nc, const mcs_agent* as, int na, int32_t agent_id,
                           int32_t direction, uint64_t key, mcs_action* out) {
  Env e;
  build_env(e, cs, nc, as, na, key);
  auto it = e.agentIdx.find(agent_id);
  if (e.error || it == e.agentIdx.end()) return MCS_ERR_SCENE;
  Action a = merging_closest_gap(e, it->second, direction);
  if (e.error) return MCS_ERR_SCENE;
  out->ax = a.ax; out->vy = a.vy;
  return MCS_OK;
}

double orc_time_to_reach_gap(const double* p) { return time_to_reach_gap(p[0], p[1], p[2], p[3], p[4], p[5], p[6]); }

}  // extern "C"

// TEST INFRASTRUCTURE ONLY — never linked into, imported by or shipped with the product.
//
// ref_harness: executes the reference's own machine code
// (/root/reference/mc_sim_highway/mc_sim_highway.so, copied to oracle/_ref/ by build_ref.sh)
// without CPython/Boost by dlopen-ing it against stub libraries
// and calling its exported C++ entry points directly.
//
// libc rand() is interposed by this executable (-rdynamic), so every random
// draw the reference makes comes from a stream we control:
//   rng counter <key>  : draw k of the stream = mcs_rand31(key, k)  (same counter-based
//                        generator as the product, include/mcsim_rng.h); the counter is
//                        reset by the commands that start a rollout
//   rng libc <seed>    : glibc srand(seed)/rand()  (to reproduce SURVEY's known-answer vectors)
//
// Scene / command protocol on stdin (whitespace separated, one record per line):
//   corridor <id> <type> <l.p1.x> <l.p1.y> <l.p2.x> <l.p2.y> <r...4> <c...4> <vLimit>
//   agent <id> <x> <y> <vx> <vy> <vd> <a> <l> <w> <idm:accMax minDis accMin accCom aE thw>
//         <rss:rTOther rTEgo aMin aMax aSoft dSoft> <yp:bias a b c pf dA bA> <behavior>
//         [<commit> <keepStep> <targetLane>]
//   sim <dt> <steps> <minSteps>
//   rng counter <key> | rng libc <seed>
//   math mcs | math libm                 which exp/pow the binary gets (default mcs = include/mcsim_math.h)
//   run <egoId> <action> <gapId> [hist] [dump]   one fresh Simulation, one rollout
//   rollouts <egoId> <action> <gapId> <first> <count> <seed> [hist]
//                                       fresh Simulation per rollout r, key = mcs_rollout_key(seed, r)
//   runm <M> <egoId> <action> <gapId>     reference runMTimes (one Simulation, M rollouts)
//   runmt <M> <egoId> <action> <gapId>    reference runMultiThreads (8 threads x M)
//   mobil <id> <ax> <vy> <requireRss>     python-facing laneChangeMobilBehavior
//   fixed_dir <id> <decision> | fixed_gap <id> <dir> <gapId> | closest_gap <id> <dir>
//   ttrg <7 doubles>                      timeToReachGap
//   order                                  agent visit order + corridor order of a fresh Environment
//   bench <mode:m|mt> <M> <egoId> <action> <gapId> <reps>   wall-clock timing
//   clear                                  drop the scene
// Every answer is one line starting with the command name; doubles are printed with %.17g.
#include "ref_abi.hpp"
#include "../../include/mcsim_rng.h"
#include "../../include/mcsim_math.h"

#include <chrono>
#include <iostream>
#include <sstream>
#include <thread>
#include <unistd.h>

using namespace refabi;

// ---------------------------------------------------------------- RNG interposition
static thread_local uint64_t g_key = 0;
static thread_local uint64_t g_ctr = 0;
static int g_mode = 0;  // 0 counter, 1 libc
static long g_reseed_after_build = -1;  // libc mode: srand(seed) again once the scene objects exist
static thread_local uint64_t g_draws = 0;
static int (*real_rand)(void) = nullptr;
static FILE* g_trace = nullptr;

// libm interposition: the binary's pow@plt / exp@plt bind here (-rdynamic).  math mode 0 routes them
// to include/mcsim_math.h (the pinned definitions shared with the product), 1 to the platform libm.
static int g_math = 0;
static double (*real_pow)(double, double) = nullptr;
static double (*real_exp)(double) = nullptr;
extern "C" double pow(double x, double y) {
  if (g_math == 0 && y == 4.0) return mcs_pow4(x);
  if (!real_pow) real_pow = reinterpret_cast<double (*)(double, double)>(dlsym(RTLD_NEXT, "pow"));
  return real_pow(x, y);
}
extern "C" double exp(double x) {
  if (g_math == 0) return mcs_exp(x);
  if (!real_exp) real_exp = reinterpret_cast<double (*)(double)>(dlsym(RTLD_NEXT, "exp"));
  return real_exp(x);
}

extern "C" int rand(void) {
  g_draws++;
  if (g_mode == 1) {
    if (!real_rand) real_rand = reinterpret_cast<int (*)(void)>(dlsym(RTLD_NEXT, "rand"));
    return real_rand();
  }
  int r = mcs_rand31(g_key, g_ctr);
  if (g_trace) std::fprintf(g_trace, "draw %llu %d\n", (unsigned long long)g_ctr, r);
  g_ctr++;
  return r;
}

// ---------------------------------------------------------------- API loading
void Api::load(const char* path) {
  // the stub libraries sit next to the binary; loading them first (by path) satisfies its
  // DT_NEEDED entries by soname without LD_LIBRARY_PATH
  std::string dir(path);
  size_t slash = dir.find_last_of('/');
  dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
  for (const char* lib : {"libpython3.6m.so.1.0", "libmc_sim_highway.so.0", "libboost_python3-py36.so.1.65.1"}) {
    std::string p = dir + "/" + lib;
    if (!dlopen(p.c_str(), RTLD_NOW | RTLD_GLOBAL)) {
      std::fprintf(stderr, "ref_harness: cannot load stub %s: %s\n", p.c_str(), dlerror());
      std::exit(2);
    }
  }
  h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { std::fprintf(stderr, "ref_harness: dlopen failed: %s\n", dlerror()); std::exit(2); }
  last_alloc = reinterpret_cast<void**>(dlsym(RTLD_DEFAULT, "oracle_last_alloc"));
  if (!last_alloc) { std::fprintf(stderr, "ref_harness: oracle_last_alloc missing\n"); std::exit(2); }
#define S(field, name) field = sym<decltype(field)>(name)
  S(make_corridor, "_ZN5boost6python7objects11make_holderILi6EE5applyINS1_12value_holderIN14mc_sim_highway8CorridorEEENS_3mpl7vector6IiNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEENS6_4LineESH_SH_dEEE7executeEP7_objectiSG_SH_SH_SH_d");
  S(make_agent, "_ZN5boost6python7objects11make_holderILi13EE5applyINS1_12value_holderIN14mc_sim_highway5AgentEEENS_3mpl8vector13IiddddddddNS6_8IdmParamENS6_8RSSParamENS6_13YieldingParamENSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEEEE7executeEP7_objectiddddddddSB_SC_SD_SJ_");
  S(corridor_copy, "_ZN14mc_sim_highway8CorridorC1ERKS0_");
  S(agent_copy, "_ZN14mc_sim_highway5AgentC1ERKS0_");
  S(agent_dtor, "_ZN14mc_sim_highway5AgentD1Ev");
  S(map_ctor, "_ZN14mc_sim_highway12MapMultiLaneC1ERKSt6vectorINS_8CorridorESaIS2_EE");
  S(map_dtor, "_ZN14mc_sim_highway12MapMultiLaneD1Ev");
  S(env_ctor, "_ZN14mc_sim_highway11EnvironmentC1ERKNS_12MapMultiLaneERKSt6vectorINS_5AgentESaIS5_EE");
  S(env_dtor, "_ZN14mc_sim_highway11EnvironmentD1Ev");
  S(env_set_ego, "_ZN14mc_sim_highway11Environment31setEgoAgentIdAndDrivingBehaviorERKiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES2_");
  S(env_match, "_ZN14mc_sim_highway11Environment21matchAgentsOnCorridorEv");
  S(sim_ctor, "_ZN14mc_sim_highway10SimulationC1ERKNS_12SimParameterERKNS_12MapMultiLaneERKSt6vectorINS_5AgentESaIS8_EE");
  S(sim_set_ego, "_ZN14mc_sim_highway10Simulation31setEgoAgentIdAndDrivingBehaviorERKiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES2_");
  S(sim_run, "_ZN14mc_sim_highway10Simulation3runEv");
  S(sim_reset, "_ZN14mc_sim_highway10Simulation5resetEv");
  S(sim_hist, "_ZN14mc_sim_highway10Simulation19agentsStatesHistoryEv");
  S(run_m_times, "_ZN14mc_sim_highway9runMTimesERKiRNS_15BehaviorFeatureERKNS_12SimParameterERKNS_12MapMultiLaneERKSt6vectorINS_5AgentESaISB_EES1_RKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES1_");
  S(run_multi_threads, "_ZN14mc_sim_highway15runMultiThreadsERKiRKNS_12SimParameterERKNS_12MapMultiLaneERKSt6vectorINS_5AgentESaIS9_EES1_RKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES1_");
  S(w_mobil, "_ZN16python_converter30laneChangeBehaviorMOBILWrapperERKN14mc_sim_highway11EnvironmentERKiRKNS0_6ActionEb");
  S(w_fixed_dir, "_ZN16python_converter35customizedLaneChangeBehaviorWrapperERKN14mc_sim_highway11EnvironmentERKiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE");
  S(w_fixed_gap, "_ZN16python_converter32customizedMergingBehaviorWrapperERKN14mc_sim_highway11EnvironmentERKiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES5_");
  S(w_closest_gap, "_ZN16python_converter32closestGapMergingBehaviorWrapperERKN14mc_sim_highway11EnvironmentERKiRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE");
  S(time_to_reach_gap, "_ZN14mc_sim_highway14timeToReachGapERKdS1_S1_S1_S1_S1_S1_");
#undef S
}

static Api api;

// ---------------------------------------------------------------- scene description
struct CorridorIn { int id; std::string type; Line l, r, c; double vLimit; };
struct AgentIn {
  int id; double x, y, vx, vy, vd, a, l, w; IdmParam idm; RSSParam rss; YieldingParam yp; std::string behavior;
  bool hasCommit = false; std::string commit; int keepStep = 0; int targetLane = -1;
};
static std::vector<CorridorIn> g_corr;
static std::vector<AgentIn> g_agents;
static SimParameter g_sim{0.3, 40, 30};

template <class T> static T& at(void* base, int off) { return *reinterpret_cast<T*>(static_cast<char*>(base) + off); }

struct Scene {  // reference objects built from g_corr / g_agents
  std::vector<CorridorBlob> corr;
  std::vector<AgentBlob> agents;
  RawVec cv{}, av{};
  MapBlob* map = nullptr;
  void build() {
    corr.resize(g_corr.size());
    for (size_t i = 0; i < g_corr.size(); ++i) {
      const auto& c = g_corr[i];
      api.make_corridor(nullptr, c.id, c.type, c.l, c.r, c.c, c.vLimit);
      api.corridor_copy(corr[i].b, static_cast<char*>(*api.last_alloc) + 0x10);
    }
    agents.resize(g_agents.size());
    for (size_t i = 0; i < g_agents.size(); ++i) {
      const auto& a = g_agents[i];
      api.make_agent(nullptr, a.id, a.x, a.y, a.vx, a.vy, a.vd, a.a, a.l, a.w, a.idm, a.rss, a.yp, a.behavior);
      void* obj = static_cast<char*>(*api.last_alloc) + 0x10;
      if (a.hasCommit) {
        at<std::string>(obj, ag::commit) = a.commit;
        at<int>(obj, ag::commitKeepLaneStep) = a.keepStep;
        at<int>(obj, ag::targetLaneId) = a.targetLane;
      }
      api.agent_copy(agents[i].b, obj);
    }
    cv = {corr.data(), corr.data() + corr.size(), corr.data() + corr.size()};
    av = {agents.data(), agents.data() + agents.size(), agents.data() + agents.size()};
    map = new MapBlob();
    std::memset(map->b, 0, sizeof(map->b));
    api.map_ctor(map->b, &cv);
    if (g_mode == 1 && g_reseed_after_build >= 0) srand((unsigned)g_reseed_after_build);
  }
  ~Scene() {
    for (auto& a : agents) api.agent_dtor(a.b);
    if (map) { api.map_dtor(map->b); delete map; }
  }
};

static void print_status(const char* tag, const StatusOneSim& st, uint64_t draws) {
  std::printf("%s finish=%d fallBack=%d finishStep=%d utility=%.17g comfort=%.17g utilityObjects=%.17g comfortObjects=%.17g draws=%llu\n",
              tag, st.finish, st.fallBack, st.finishStep, st.utility, st.comfort, st.utilityObjects, st.comfortObjects,
              (unsigned long long)draws);
}
static void print_feature(const char* tag, const BehaviorFeature& f) {
  std::printf("%s fallBackRate=%.17g successRate=%.17g utility=%.17g comfort=%.17g expectedStepRatio=%.17g utilityObjects=%.17g comfortObjects=%.17g fromNSims=%d\n",
              tag, f.fallBackRate, f.successRate, f.utility, f.comfort, f.expectedStepRatio, f.utilityObjects, f.comfortObjects, f.fromNSims);
}

using AgentMap = std::unordered_map<int, std::shared_ptr<char>>;   // ABI-compatible view of unordered_map<int, shared_ptr<Agent>>

static void dump_env(void* env) {
  auto& agents = at<AgentMap>(env, 0xd0);
  for (auto& kv : agents) {
    void* a = kv.second.get();
    auto& ch = at<std::vector<int>>(a, ag::corridorHistory);
    auto& thw = at<std::vector<double>>(a, ag::thwRatio);
    std::printf("agent id=%d x=%.17g y=%.17g v=%.17g vy=%.17g a=%.17g vd=%.17g prior=%.17g,%.17g,%.17g isEgo=%d reached=%d gap=%d tlane=%d cmd=%s commit=%s keep=%d aggr=%d beh=%s lane=%d nlanes=%zu nthw=%zu thwlast=%.17g nint=%zu\n",
                at<int>(a, ag::id), at<double>(a, ag::x), at<double>(a, ag::y), at<double>(a, ag::v), at<double>(a, ag::vy),
                at<double>(a, ag::a), at<double>(a, ag::vd), at<double>(a, ag::lcPrior), at<double>(a, ag::lcPrior + 8),
                at<double>(a, ag::lcPrior + 16), at<uint8_t>(a, ag::isEgo), at<uint8_t>(a, ag::reachedTarget),
                at<int>(a, ag::targetGapId), at<int>(a, ag::targetLaneId), at<std::string>(a, ag::egoCommand).c_str(),
                at<std::string>(a, ag::commit).c_str(), at<int>(a, ag::commitKeepLaneStep), at<uint8_t>(a, ag::aggressive),
                at<std::string>(a, ag::behavior).c_str(), ch.empty() ? -999 : ch.back(), ch.size(), thw.size(),
                thw.empty() ? 0.0 : thw.back(),
                at<std::unordered_map<int, std::array<char, 32>>>(a, ag::intentions).size());
  }
}

// one fresh Simulation + one rollout, RNG counter reset first; agents are re-constructed so that
// draws 0..N-1 are the Agent-ctor aggressive flags (input order) exactly like a first rollout.
static StatusOneSim one_rollout(int ego, const std::string& action, int gap, bool hist, bool dump, bool* ok_out) {
  g_ctr = 0; g_draws = 0;
  Scene sc; sc.build();
  SimBlob* sim = new SimBlob();
  std::memset(sim->b, 0, sizeof(sim->b));
  api.sim_ctor(sim->b, &g_sim, sc.map->b, &sc.av);
  bool ok = api.sim_set_ego(sim->b, &ego, &action, &gap);
  if (ok_out) *ok_out = ok;
  StatusOneSim st = api.sim_run(sim->b);
  if (hist) {
    auto h = api.sim_hist(sim->b);
    for (auto& kv : h) {
      std::printf("hist id=%d n=%zu", kv.first, kv.second.size());
      for (auto& s : kv.second) std::printf(" %.17g %.17g %.17g %.17g", s.x, s.y, s.v, s.a);
      std::printf("\n");
    }
  }
  if (dump) dump_env(sim->b + 8);
  // Simulation has no exported dtor; Environment at +8 has one.
  api.env_dtor(sim->b + 8);
  delete sim;
  return st;
}

int main(int argc, char** argv) {
  const char* so = argc > 1 ? argv[1] : "mc_sim_highway.so";
  api.load(so);
  if (const char* t = std::getenv("REF_TRACE_DRAWS")) g_trace = std::fopen(t, "w");
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream in(line);
    std::string cmd;
    if (!(in >> cmd) || cmd[0] == '#') continue;
    if (cmd == "clear") { g_corr.clear(); g_agents.clear(); }
    else if (cmd == "corridor") {
      CorridorIn c;
      in >> c.id >> c.type >> c.l.p1.x >> c.l.p1.y >> c.l.p2.x >> c.l.p2.y >> c.r.p1.x >> c.r.p1.y >> c.r.p2.x >> c.r.p2.y
         >> c.c.p1.x >> c.c.p1.y >> c.c.p2.x >> c.c.p2.y >> c.vLimit;
      g_corr.push_back(c);
    } else if (cmd == "agent") {
      AgentIn a;
      in >> a.id >> a.x >> a.y >> a.vx >> a.vy >> a.vd >> a.a >> a.l >> a.w
         >> a.idm.accMax >> a.idm.minDis >> a.idm.accMin >> a.idm.accCom >> a.idm.aE >> a.idm.thw
         >> a.rss.rTOther >> a.rss.rTEgo >> a.rss.aMin >> a.rss.aMax >> a.rss.aSoft >> a.rss.dSoft
         >> a.yp.bias >> a.yp.a >> a.yp.b >> a.yp.c >> a.yp.pf >> a.yp.dA >> a.yp.bA >> a.behavior;
      if (in >> a.commit >> a.keepStep >> a.targetLane) a.hasCommit = true;
      g_agents.push_back(a);
    } else if (cmd == "sim") { in >> g_sim.dt >> g_sim.steps >> g_sim.minSteps; }
    else if (cmd == "math") { std::string m; in >> m; g_math = (m == "libm") ? 1 : 0; }
    else if (cmd == "rng") {
      std::string m; unsigned long long s; in >> m >> s;
      g_reseed_after_build = -1;
      if (m == "libc") { g_mode = 1; srand((unsigned)s); }
      else if (m == "libc_after_build") { g_mode = 1; srand((unsigned)s); g_reseed_after_build = (long)s; } else { g_mode = 0; g_key = s; g_ctr = 0; }
    } else if (cmd == "run") {
      int ego, gap; std::string action, opt; in >> ego >> action >> gap;
      bool hist = false, dump = false;
      while (in >> opt) { if (opt == "hist") hist = true; if (opt == "dump") dump = true; }
      bool ok; StatusOneSim st = one_rollout(ego, action, gap, hist, dump, &ok);
      std::printf("run ok=%d ", ok); print_status("", st, g_draws);
    } else if (cmd == "rollouts") {
      int ego, gap; std::string action, opt; long first, count; unsigned long long seed;
      in >> ego >> action >> gap >> first >> count >> seed;
      bool hist = false; while (in >> opt) if (opt == "hist") hist = true;
      for (long r = first; r < first + count; ++r) {
        g_key = mcs_rollout_key(seed, (uint64_t)r);
        bool ok; StatusOneSim st = one_rollout(ego, action, gap, hist, false, &ok);
        std::printf("rollout r=%ld ok=%d ", r, ok); print_status("", st, g_draws);
      }
    } else if (cmd == "runm" || cmd == "runmt") {
      int M, ego, gap; std::string action; in >> M >> ego >> action >> gap;
      g_ctr = 0; g_draws = 0;
      Scene sc; sc.build();
      BehaviorFeature f{};
      if (cmd == "runm") api.run_m_times(&M, &f, &g_sim, sc.map->b, &sc.av, &ego, &action, &gap);
      else f = api.run_multi_threads(&M, &g_sim, sc.map->b, &sc.av, &ego, &action, &gap);
      print_feature(cmd.c_str(), f);
    } else if (cmd == "mobil" || cmd == "fixed_dir" || cmd == "fixed_gap" || cmd == "closest_gap" || cmd == "order" || cmd == "envdump") {
      g_ctr = 0; g_draws = 0;
      Scene sc; sc.build();
      EnvBlob* env = new EnvBlob(); std::memset(env->b, 0, sizeof(env->b));
      api.env_ctor(env->b, sc.map->b, &sc.av);
      if (cmd == "mobil") {
        int id, req; Action act; in >> id >> act.ax >> act.vy >> req;
        ActionWithType r = api.w_mobil(env->b, &id, &act, req != 0);
        std::printf("mobil ax=%.17g vy=%.17g commit=%s keep=%d target=%d draws=%llu\n", r.ax, r.vy,
                    r.commitToDecision.c_str(), r.commitKeepLaneStep, r.targetLaneId, (unsigned long long)g_draws);
      } else if (cmd == "fixed_dir") {
        int id; std::string d; in >> id >> d;
        Action r = api.w_fixed_dir(env->b, &id, &d);
        std::printf("fixed_dir ax=%.17g vy=%.17g draws=%llu\n", r.ax, r.vy, (unsigned long long)g_draws);
      } else if (cmd == "fixed_gap") {
        int id, gap; std::string d; in >> id >> d >> gap;
        Action r = api.w_fixed_gap(env->b, &id, &d, &gap);
        std::printf("fixed_gap ax=%.17g vy=%.17g draws=%llu\n", r.ax, r.vy, (unsigned long long)g_draws);
      } else if (cmd == "closest_gap") {
        int id; std::string d; in >> id >> d;
        Action r = api.w_closest_gap(env->b, &id, &d);
        std::printf("closest_gap ax=%.17g vy=%.17g draws=%llu\n", r.ax, r.vy, (unsigned long long)g_draws);
      } else if (cmd == "order") {
        auto& agents = at<AgentMap>(env->b, 0xd0);
        std::printf("order agents");
        for (auto& kv : agents) std::printf(" %d", kv.first);
        auto& corr = at<std::unordered_map<int, CorridorBlob>>(env->b, 0x50);
        std::printf(" corridors");
        for (auto& kv : corr) std::printf(" %d", kv.first);
        std::printf("\n");
      } else {
        dump_env(env->b);
        std::printf("envdump done\n");
      }
      api.env_dtor(env->b);
      delete env;
    } else if (cmd == "ttrg") {
      double d[7]; for (double& x : d) in >> x;
      std::printf("ttrg %.17g\n", api.time_to_reach_gap(&d[0], &d[1], &d[2], &d[3], &d[4], &d[5], &d[6]));
    } else if (cmd == "bench") {
      std::string mode, action; int M, ego, gap, reps; in >> mode >> M >> ego >> action >> gap >> reps;
      Scene sc; sc.build();
      double best = 1e300; std::vector<double> ts;
      BehaviorFeature f{};
      for (int i = 0; i < reps; ++i) {
        g_ctr = 0;
        auto t0 = std::chrono::steady_clock::now();
        if (mode == "m") api.run_m_times(&M, &f, &g_sim, sc.map->b, &sc.av, &ego, &action, &gap);
        else f = api.run_multi_threads(&M, &g_sim, sc.map->b, &sc.av, &ego, &action, &gap);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ts.push_back(dt); best = std::min(best, dt);
      }
      std::printf("bench mode=%s M=%d rollouts=%d best_s=%.9g expectedStepRatio=%.17g successRate=%.17g times=", mode.c_str(), M,
                  mode == "m" ? M : 8 * M, best, f.expectedStepRatio, f.successRate);
      for (double t : ts) std::printf("%.9g,", t);
      std::printf("\n");
    } else {
      std::printf("error unknown command %s\n", cmd.c_str());
    }
    std::fflush(stdout);
  }
  std::fflush(stdout);
  _exit(0);  // skip the reference binary's static destructors (they would call into the stubs)
}

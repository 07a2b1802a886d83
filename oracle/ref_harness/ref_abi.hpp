// TEST INFRASTRUCTURE ONLY — never linked into the product.
//
// Binary interface of the reference's precompiled inner simulator
// (/root/reference/mc_sim_highway/mc_sim_highway.so), as recovered from its
// exported symbols and disassembly.  Everything here
// is a *description* of that binary (struct layouts, mangled entry-point
// names); no reference source exists and none is copied.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>

namespace refabi {

struct Point { double x, y; };
struct Line { Point p1, p2; };
// memory order (ctor order of the Python class differs: accMax, accMin, minDis, accCom, aE, thw)
struct IdmParam { double accMax, minDis, accMin, accCom, aE, thw; };
struct RSSParam { double rTOther, rTEgo, aMin, aMax, aSoft, dSoft; };
struct YieldingParam { double bias, a, b, c, pf, dA, bA; };
struct SimParameter { double dt; int steps; int minSteps; };
struct Action { double ax, vy; };
struct State { double x, y, v, a; };
struct StatusOneSim {
  uint8_t fallBack, finish;
  double utility, comfort, utilityObjects, comfortObjects;
  int finishStep;
};
struct BehaviorFeature {
  double fallBackRate, successRate, utility, comfort, expectedStepRatio, utilityObjects, comfortObjects;
  int fromNSims;
};
struct ActionWithType {
  double ax, vy;
  std::string commitToDecision;
  int commitKeepLaneStep;
  int targetLaneId;
};

constexpr size_t kSizeofCorridor = 0xa0;
constexpr size_t kSizeofAgent = 0x270;

// Opaque storage for reference objects.
struct alignas(16) CorridorBlob { unsigned char b[kSizeofCorridor]; };
struct alignas(16) AgentBlob { unsigned char b[kSizeofAgent]; };
struct alignas(64) MapBlob { unsigned char b[0x400]; };
struct alignas(64) EnvBlob { unsigned char b[0x1000]; };
struct alignas(64) SimBlob { unsigned char b[0x1000]; };
// a raw std::vector<T> {begin,end,cap}
struct RawVec { void* begin; void* end; void* cap; };

// Agent field offsets
namespace ag {
constexpr int id = 0, x = 8, y = 0x10, v = 0x18, vy = 0x20, a = 0x28, l = 0x30, w = 0x38, vd = 0x40,
              lcPrior = 0x48, isEgo = 0x60, reachedTarget = 0x61, targetGapId = 0x64, targetLaneId = 0x68,
              egoCommand = 0x70, commit = 0x90, commitKeepLaneStep = 0xb0, aggressive = 0xb4, idm = 0xb8,
              idmYield = 0xe8, rss = 0x118, rssRelaxed = 0x148, behavior = 0x178, behaviorHistory = 0x198,
              states = 0x1b0, thwRatio = 0x1c8, corridorHistory = 0x1e0, yieldParam = 0x200, intentions = 0x238;
}

struct Api {
  void* h = nullptr;
  void** last_alloc = nullptr;
  template <class F> F sym(const char* name) {
    void* p = dlsym(h, name);
    if (!p) { std::fprintf(stderr, "ref_harness: missing symbol %s\n", name); std::exit(3); }
    return reinterpret_cast<F>(p);
  }
  void (*make_corridor)(void*, int, std::string, Line, Line, Line, double);
  void (*make_agent)(void*, int, double, double, double, double, double, double, double, double,
                     IdmParam, RSSParam, YieldingParam, std::string);
  void (*corridor_copy)(void*, const void*);
  void (*agent_copy)(void*, const void*);
  void (*agent_dtor)(void*);
  void (*map_ctor)(void*, const RawVec*);
  void (*map_dtor)(void*);
  void (*env_ctor)(void*, const void*, const RawVec*);
  void (*env_dtor)(void*);
  bool (*env_set_ego)(void*, const int*, const std::string*, const int*);
  void (*env_match)(void*);
  void (*sim_ctor)(void*, const SimParameter*, const void*, const RawVec*);
  bool (*sim_set_ego)(void*, const int*, const std::string*, const int*);
  StatusOneSim (*sim_run)(void*);
  void (*sim_reset)(void*);
  std::map<int, std::vector<State>> (*sim_hist)(void*);
  void (*run_m_times)(const int*, BehaviorFeature*, const SimParameter*, const void*, const RawVec*,
                      const int*, const std::string*, const int*);
  BehaviorFeature (*run_multi_threads)(const int*, const SimParameter*, const void*, const RawVec*,
                                       const int*, const std::string*, const int*);
  // python_converter wrappers (Environment const&, ...)
  ActionWithType (*w_mobil)(const void*, const int*, const Action*, bool);
  Action (*w_fixed_dir)(const void*, const int*, const std::string*);
  Action (*w_fixed_gap)(const void*, const int*, const std::string*, const int*);
  Action (*w_closest_gap)(const void*, const int*, const std::string*);
  double (*time_to_reach_gap)(const double*, const double*, const double*, const double*, const double*,
                              const double*, const double*);

  void load(const char* path);
};

}  // namespace refabi

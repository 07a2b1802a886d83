/* mcsim_rng.h — the counter-based random stream that replaces libc rand() on the
 * rollout path (reference call sites: mc_sim_highway.so 0xf89e6 Agent ctor,
 * 0xc2a32 runMTimes, 0xf510f MOBIL decision, 0x10469c yielding intention,
 * 0x104da0 acceleration noise; SURVEY.md §3.4 "RNG").
 *
 * The reference draws from the process-global, lock-protected libc stream, which
 * makes its 8-thread path nondeterministic.  Here every rollout owns an
 * independent stream: draw k of rollout r is a pure function of (seed, r, k), so
 * rollouts can run in any order on any GPU and stay bit-reproducible.  The value
 * range is exactly rand()'s, [0, 2^31-1], because the reference uses the draws as
 * r/2147483647.0, r/2147483647.0-0.5 and r%100.
 *
 * Generator: SplitMix64 output function applied to key + k*golden (a counter
 * hash; passes BigCrush as a sequential generator), top 31 bits.
 */
#ifndef MCSIM_RNG_H
#define MCSIM_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define MCS_HD __host__ __device__ __forceinline__
#else
#define MCS_HD static inline
#endif

MCS_HD uint64_t mcs_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* key of rollout `rollout` under user seed `seed` */
MCS_HD uint64_t mcs_rollout_key(uint64_t seed, uint64_t rollout) {
  return mcs_mix64(mcs_mix64(seed + 0x9E3779B97F4A7C15ULL) ^ (rollout * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL));
}

/* draw `ctr` of the stream `key`, in [0, 2^31-1] like rand() */
MCS_HD int mcs_rand31(uint64_t key, uint64_t ctr) {
  return (int)(mcs_mix64(key + (ctr + 1) * 0x9E3779B97F4A7C15ULL) >> 33);
}

#endif /* MCSIM_RNG_H */

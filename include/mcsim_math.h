/* mcsim_math.h — the two transcendental functions of the rollout path, defined
 * bit-for-bit so that host and device agree.
 *
 * The reference binary calls libm through the PLT at exactly two kinds of site:
 *   pow(v/vd, 4.0)   IDM free-road term       (mc_sim_highway.so 0xebcb2, 0xeb950, 0xb9743)
 *   exp(x)           logistic / softmax / Gaussian prior
 *                    (0x104669 yielding, 0xf4f80.. MOBIL, 0x10bbef.. lane-change prior, 0xc89c0)
 * glibc's results for these are not correctly rounded (≈0.51 ulp) and depend on the
 * libm build, so "the reference's numbers" are only defined up to the platform's libm.
 * This header pins them: every operation below is an IEEE-754 binary64 add, multiply
 * or fused multiply-add, which x86-64 (with or without FMA hardware: fma() is exact
 * either way) and sm_100a evaluate identically.  The oracle harness interposes these
 * two symbols into the reference binary (oracle/ref_harness), the C restatement and the
 * CUDA kernels include this header, so all three produce identical bits.  Against the
 * reference AS SHIPPED (platform libm, nothing interposed) the difference is one or two
 * units in the last place of a state — measured 3.9e-16 on the records of
 * tests/golden/libm_*.json.gz (every BASELINE.json configuration shape), with zero
 * discrete mismatches; stated tolerance 1e-12 per vehicle and step
 * (tests/test_libm_parity.py on the CPU, tests/test_gpu_parity.py::
 * test_against_the_stock_reference on the GPU).
 *
 * Compile hosts with -ffp-contract=off and devices with -fmad=false: only the fma()
 * calls written here may fuse.
 */
#ifndef MCSIM_MATH_H
#define MCSIM_MATH_H
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define MCS_MHD __host__ __device__ __forceinline__
#else
#define MCS_MHD static inline
#endif

MCS_MHD double mcs_bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, sizeof d);
  return d;
#endif
}

/* x^4 to within 2^-104 relative before the final rounding (double-double square of a
 * double-double square): correctly rounded except for astronomically rare ties. */
MCS_MHD double mcs_pow4(double x) {
  if (!(fabs(x) < 1.0e75)) return (x * x) * (x * x); /* inf / nan / overflow range */
  double hi = x * x;
  double lo = fma(x, x, -hi);
  double p = hi * hi;
  double e = fma(hi, hi, -p);
  e = fma(hi + hi, lo, e);
  return p + e;
}

/* exp(x), < 1 ulp.  k = round(x/ln2), r = x - k ln2 (two-word ln2), degree-13 Taylor
 * polynomial by Horner, scaled by 2^k in two exact steps so subnormal results round once. */
MCS_MHD double mcs_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return mcs_bits_to_double(0x7ff0000000000000ULL);
  if (x < -745.2) return 0.0;
  const double shifter = 6755399441055744.0; /* 1.5 * 2^52 */
  double t = fma(x, 1.4426950408889634, shifter);
  double kd = t - shifter;
  double r = fma(-kd, 0.693147180369123816490, x);      /* ln2 high part (32 significant bits) */
  r = fma(-kd, 1.90821492927058770002e-10, r);          /* ln2 low part */
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int k = (int)kd;
  /* 2^k as two exact factors, each a normal double */
  int k1 = k / 2, k2 = k - k1;
  double s1 = mcs_bits_to_double((uint64_t)(k1 + 1023) << 52);
  double s2 = mcs_bits_to_double((uint64_t)(k2 + 1023) << 52);
  return (p * s1) * s2;
}

#endif /* MCSIM_MATH_H */

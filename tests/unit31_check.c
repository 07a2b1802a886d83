/* Exhaustive check behind mcs_unit31 (csrc/rollout_kernel.cuh): for every 31-bit draw r, r / 2147483647.0 computed as the
 * product with the rounded reciprocal, corrected once by its exact remainder, equals the IEEE division bit for bit.
 * Prints the number of mismatches (0).  Test infrastructure: compiled and run by tests/test_host_logic.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void) {
  const double c = 2147483647.0, y = 1.0 / 2147483647.0;
  uint64_t bad = 0;
  for (int64_t r = 0; r < 2147483648LL; ++r) {
    const double a = (double)r, q = a / c, q0 = a * y;
    const double q1 = fma(fma(-q0, c, a), y, q0);
    if (memcmp(&q, &q1, sizeof q) != 0) bad++;
  }
  printf("%llu\n", (unsigned long long)bad);
  return 0;
}

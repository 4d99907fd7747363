/* Plain-C restatement of fast_hist (utils/utils_metrics.py:34-43 of the reference) -- TEST INFRASTRUCTURE ONLY.
 * k = (a >= 0) & (a < n); hist[n*a[k] + b[k]] += 1.  a, b are uint8 (np.array of 'L'/'P' PNGs), so a >= 0 always.
 * Returns the number of pairs whose bin would fall outside n*n (numpy's reshape raises in that case).
 * Built by oracle/Makefile into oracle/_ref/libfasthist.so; pinned by tests/test_oracle.py against tests/golden/. */
#include <stdint.h>
#include <stddef.h>

long long oracle_fast_hist_u8(const uint8_t* a, const uint8_t* b, size_t len, int n, long long* hist) {
  long long overflow = 0;
  for (size_t i = 0; i < len; ++i) {
    if (a[i] < n) {
      const long long bin = (long long)n * a[i] + b[i];
      if (bin < (long long)n * n) hist[bin] += 1; else overflow += 1;
    }
  }
  return overflow;
}

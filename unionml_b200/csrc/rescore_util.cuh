// Warp-level helpers of the float64 re-score kernels: they are latency-bound (a handful of rows per warp), so the
// reductions are arranged to need few dependent shuffle rounds.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace uml {

// Sum 16 per-lane values across the warp with 16 shuffles instead of 16 x 5: every butterfly step halves the number
// of values a lane is still responsible for (lanes with the offset bit set keep the upper half).  On return lane l
// holds the complete sum of value (l >> 1); the two lanes of a pair hold the same one.
__device__ __forceinline__ double warp_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) {
    const int o = 2 * m;  // 16, 8, 4, 2
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const double send = upper ? v[i] : v[i + m];
      const double keep = upper ? v[i + m] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Running (best, second, argmax) with numpy's first-maximum rule; `second` is the largest of the OTHER scores, so an
// exact tie gives second == best (margin 0 -> "ambiguous").
struct Top2 {
  double best, second;
  int idx;
};

__device__ __forceinline__ void top2_merge(Top2& t, double ob, double os, int oi) {
  const bool take = ob > t.best || (ob == t.best && oi < t.idx);
  const double loser = take ? t.best : ob;
  t.second = fmax(fmax(t.second, os), loser);
  if (take) {
    t.best = ob;
    t.idx = oi;
  }
}

// all-lanes top-2 over per-lane candidates; first_offset = 2 when lane pairs hold the same class (after warp_reduce16)
__device__ __forceinline__ void top2_butterfly(Top2& t, int first_offset) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (o < first_offset) break;
    const double ob = __shfl_xor_sync(0xffffffffu, t.best, o);
    const double os = __shfl_xor_sync(0xffffffffu, t.second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, t.idx, o);
    top2_merge(t, ob, os, oi);
  }
}

__device__ __forceinline__ double warp_max(double v, int first_offset) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (o < first_offset) break;
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}

}  // namespace uml

// float64 re-score of ONE row of the 2-layer MLP by one warp, with every operand in shared memory.  Shared by the
// stand-alone re-score kernel (mlp_kernels.cu) and the re-score warps inside the tensor-core kernel (mlp_tc_kernels.cu).
#pragma once

#include "rescore_util.cuh"
#include "tma_ring.cuh"

namespace uml {

// fp64 operands of one block, staged once: W1 [F][H], W2 [C][H + 1] (padded rows: lane-per-class reads are conflict
// free), biases, and the two bound vectors w1m[f] = max_n |w1_nf|, w2m[n] = max_c |w2_cn|
struct MlpRsView {
  const double* w1s;
  const double* w2s;
  const double* b1s;
  const double* b2s;
  const double* w1m;
  const double* w2m;
  double b1max, b2max, w2sum;
  int F, H, C;
};

__host__ __device__ inline size_t mlp_rs_weight_doubles(int F, int H, int C) {
  return static_cast<size_t>(F) * H + static_cast<size_t>(C) * (H + 1) + H + C + F + H;
}
__host__ __device__ inline size_t mlp_rs_strip_doubles(int F, int H) { return static_cast<size_t>(F) + H; }

// all threads of the block: copy the fp64 weights from global memory; returns the view.  After a block-wide barrier
// every thread calls mlp_rs_finish_stage (bound vectors from the staged copies), then another barrier before rows.
__device__ inline MlpRsView mlp_rs_stage(double* smem, const double* w1, const double* b1, const double* w2,
                                         const double* b2, int F, int H, int C) {
  const int HP = H + 1;
  double* w1s = smem;
  double* w2s = w1s + F * H;
  double* b1s = w2s + C * HP;
  double* b2s = b1s + H;
  for (int i = threadIdx.x; i < F * H; i += blockDim.x) w1s[i] = w1[i];
  for (int i = threadIdx.x; i < C * H; i += blockDim.x) w2s[(i / H) * HP + (i % H)] = w2[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) b1s[i] = b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) b2s[i] = b2[i];
  MlpRsView v;
  v.w1s = w1s;
  v.w2s = w2s;
  v.b1s = b1s;
  v.b2s = b2s;
  v.w1m = b2s + C;
  v.w2m = v.w1m + F;
  v.F = F;
  v.H = H;
  v.C = C;
  v.b1max = v.b2max = v.w2sum = 0.0;
  return v;
}

__device__ inline void mlp_rs_finish_stage(MlpRsView& v) {
  const int F = v.F, H = v.H, C = v.C, HP = v.H + 1;
  double* w1m = const_cast<double*>(v.w1m);
  double* w2m = const_cast<double*>(v.w2m);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double m = 0.0;
    for (int hn = 0; hn < H; ++hn) m = fmax(m, fabs(v.w1s[f * H + hn]));
    w1m[f] = m;
  }
  for (int hn = threadIdx.x; hn < H; hn += blockDim.x) {
    double m = 0.0;
    for (int c = 0; c < C; ++c) m = fmax(m, fabs(v.w2s[c * HP + hn]));
    w2m[hn] = m;
  }
  // scalars: every thread derives its own copy from shared memory (broadcast reads)
  double b1max = 0.0, b2max = 0.0, w2sum = 0.0;
  for (int hn = 0; hn < H; ++hn) {
    b1max = fmax(b1max, fabs(v.b1s[hn]));
    double m = 0.0;
    for (int c = 0; c < C; ++c) m = fmax(m, fabs(v.w2s[c * HP + hn]));
    w2sum += m;  // sum_n max_c |w2_cn|: how far a hidden-layer error can move any logit
  }
  for (int c = 0; c < C; ++c) b2max = fmax(b2max, fabs(v.b2s[c]));
  v.b1max = b1max;
  v.b2max = b2max;
  v.w2sum = w2sum;
}

struct MlpRowResult {
  int idx;
  bool bad;        // NaN/Inf in the row
  bool ambiguous;  // fp64 logit margin inside the fp64 rounding bound (a true tie; first index wins)
};

// one warp, one row: xr = the row's fp32 features (global memory), xs / hv = the warp's strip (F and H doubles)
__device__ __forceinline__ MlpRowResult mlp_rs_row(const MlpRsView& v, const float* __restrict__ xr, double* xs, double* hv,
                                                  int lane) {
  const double u = 1.1102230246251565e-16;  // 2^-53
  const int F = v.F, H = v.H, C = v.C, HP = v.H + 1;
  bool bad = false;
  double a1 = 0.0;  // sum_f |x_f| max_n |w1_nf|: bounds every hidden unit's absolute sum (one chain instead of H)
  for (int f = lane; f < F; f += 32) {
    const float xf = xr[f];
    bad |= !isfinite(xf);
    const double xd = static_cast<double>(xf);
    xs[f] = xd;
    a1 = fma(fabs(xd), v.w1m[f], a1);
  }
  bad = __any_sync(0xffffffffu, bad);
  a1 = warp_sum(a1) + v.b1max;
  const double herr = (F + 6.0) * u * a1;  // any hidden unit's own fp64 rounding error (four partial chains + their sum)
  __syncwarp();  // the strip writes above are read by other lanes below
  // ---- hidden layer: lane per unit, four chains over the features ----
  double a2 = 0.0;  // sum_n h_n max_c |w2_cn|: bounds every logit's absolute sum
  for (int hn = lane; hn < H; hn += 32) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int f = 0;
    for (; f + 4 <= F; f += 4) {
      s0 = fma(xs[f], v.w1s[f * H + hn], s0);
      s1 = fma(xs[f + 1], v.w1s[(f + 1) * H + hn], s1);
      s2 = fma(xs[f + 2], v.w1s[(f + 2) * H + hn], s2);
      s3 = fma(xs[f + 3], v.w1s[(f + 3) * H + hn], s3);
    }
    for (; f < F; ++f) s0 = fma(xs[f], v.w1s[f * H + hn], s0);
    const double h = fmax(((s0 + s1) + (s2 + s3)) + v.b1s[hn], 0.0);
    hv[hn] = h;
    a2 = fma(h, v.w2m[hn], a2);
  }
  const double amax = warp_sum(a2) + herr * v.w2sum + v.b2max;
  __syncwarp();
  // ---- output layer: lane per class ----
  double best = 0.0, second = -INFINITY;
  int idx = 0;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    double s0 = 0.0, s1 = 0.0;
    if (c < C) {
      const double* w2c = v.w2s + c * HP;
      int nn = 0;
      for (; nn + 2 <= H; nn += 2) {
        s0 = fma(hv[nn], w2c[nn], s0);
        s1 = fma(hv[nn + 1], w2c[nn + 1], s1);
      }
      for (; nn < H; ++nn) s0 = fma(hv[nn], w2c[nn], s0);
    }
    Top2 t;
    t.best = c < C ? (s0 + s1) + v.b2s[c] : -INFINITY;
    t.second = -INFINITY;
    t.idx = c;
    top2_butterfly(t, 1);
    if (c0 == 0) {
      best = t.best;
      second = t.second;
      idx = t.idx;
    } else if (t.best > best) {
      second = fmax(best, t.second);
      best = t.best;
      idx = t.idx;
    } else {
      second = fmax(second, t.best);
    }
  }
  __syncwarp();  // the strip may be reused by the caller's next row
  MlpRowResult r;
  r.idx = idx >= C ? 0 : idx;  // idx >= C only with NaN scores, which are reported through `bad`
  r.bad = bad;
  // fp64 error of a logit: the hidden units' own errors carried through W2, plus the output layer's chain
  const double err = herr * v.w2sum + (static_cast<double>(H) + 16.0) * u * amax;
  r.ambiguous = !((best - second) > 2.0 * err);
  return r;
}

}  // namespace uml

// fm_rowgroup.cuh -- the sub-warp "row group" used by the fp32 kernels.
//
// One training example (one CSR row) is handled by E = G*S lanes of a warp:
//   G lanes span one factor row V[id,:] as G float4 chunks (kp = 4*gp floats,
//     gp <= G; lanes with c >= gp idle when kp/4 is not a power of two),
//   S "slots" walk the row's non-zeros S at a time.
// A warp therefore processes 32/E examples at once (k=8, 2 nnz/row: G=2, S=2,
// 8 examples per warp) -- the "one warp per example" of the north star realised
// as sub-warp tiles so that 16 different V rows are gathered per LDG.128.
//
// The work is split in two phases so that a caller can put several rows'
// gathers in flight before consuming any of them (memory-level parallelism):
//   gather()  issues the loads of the first R entries per lane into registers,
//   reduce()  restates fm_model::predict (reference src/fm_core/fm_model.h:105-127)
//             in fp32 with the O(k*nnz) trick: per-lane partial sums, segmented
//             __shfl_xor reductions over the slot bits (per-factor sums) and then
//             over the whole group (scalar score).  Entries beyond the R cached
//             ones are gathered inside reduce() and re-gathered by the update.
#pragma once
#include "fm_device.cuh"

namespace fmb {

template <int G, int S, int R>
struct RowGroup {
  static constexpr int E = G * S;
  static_assert(E <= 32 && (E & (E - 1)) == 0, "group must be a power-of-two slice of a warp");

  float4 acc;       // per-factor sums s_f for this lane's 4 factors (complete after reduce())
  float4 vc[R];     // cached V chunks
  float xc[R];      // cached x values (0 for inactive entries)
  float wc[R];      // cached w values (lane c == 0 only)
  uint32_t idc[R];  // cached feature ids
  int beg, end;     // this row's entries: [beg, end) in the id / value arrays
  int maxit;        // warp-uniform iteration count of the entry loop
  float hrow;       // (WANT_H) curvature of the score w.r.t. all of this row's w/V blocks

  template <typename IdPtr, typename ValPtr>
  __device__ __forceinline__ void gather(const float4* __restrict__ V4,
                                         const float* __restrict__ w, int gp, int ws, bool use_w,
                                         IdPtr ids, ValPtr xs, int beg_, int end_, int c, int s) {
    beg = beg_;
    end = end_;
    const bool chunk_on = c < gp;
#pragma unroll
    for (int it = 0; it < R; ++it) {
      const int j = beg + s + it * S;
      const bool on = j < end;
      uint32_t id = 0;
      float x = 0.f;
      if (on) {
        id = ids[j];
        x = xs[j];
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float wv = 0.f;
      if (on && chunk_on) v = ld_cg_f4(V4 + (size_t)id * gp + c);
      if (on && use_w && c == 0) wv = ld_cg_f(w + (size_t)id * ws);
      idc[it] = id;
      xc[it] = x;
      vc[it] = v;
      wc[it] = wv;
    }
  }

  // Returns the score WITHOUT the bias term, replicated in all E lanes.
  // WANT_H additionally fills hrow = sum_i |d score / d (w_i, V_i)|^2, evaluated with
  // the one-hot identity  A*k1 + (A-2)*sum_f s_f^2 + sum_f sum_i (v_if x_i)^2,
  // A = sum_i x_i^2  (exact for x in {0,1}, a damping heuristic otherwise).
  template <bool WANT_H, typename IdPtr, typename ValPtr>
  __device__ __forceinline__ float reduce(const float4* __restrict__ V4,
                                          const float* __restrict__ w, int gp, int ws, bool use_w,
                                          IdPtr ids, ValPtr xs, int c, int s) {
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f, lin = 0.f, xx = 0.f;
    const int iters = (end - beg + S - 1) / S;
    maxit = __reduce_max_sync(0xffffffffu, iters);
#pragma unroll
    for (int it = 0; it < R; ++it) {
      accumulate(vc[it], xc[it], wc[it], sq, lin);
      if (WANT_H && c == 0) xx += xc[it] * xc[it];
    }
    const bool chunk_on = c < gp;
    for (int it = R; it < maxit; ++it) {
      const int j = beg + s + it * S;
      if (j < end) {
        const uint32_t id = ids[j];
        const float x = xs[j];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float wv = 0.f;
        if (chunk_on) v = ld_cg_f4(V4 + (size_t)id * gp + c);
        if (use_w && c == 0) wv = ld_cg_f(w + (size_t)id * ws);
        accumulate(v, x, wv, sq, lin);
        if (WANT_H && c == 0) xx += x * x;
      }
    }
    // per-factor sums: reduce over the slot bits (lane strides G, 2G, ... < E)
#pragma unroll
    for (int o = G; o < E; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    // 0.5*(sum_f^2 - sumsq_f): the square term once per chunk (slot 0), the
    // rest from every lane; then reduce the scalar over the whole group
    float part = lin - 0.5f * sq;
    float s2 = 0.f;
    if (s == 0) {
      s2 = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
      part += 0.5f * s2;
    }
#pragma unroll
    for (int o = 1; o < E; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (WANT_H) {
#pragma unroll
      for (int o = 1; o < E; o <<= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
      }
      hrow = (use_w ? xx : 0.f) + fmaxf((xx - 2.f) * s2 + sq, 0.f);
    }
    return part;
  }

  // gather + reduce in one go (scoring kernels)
  template <typename IdPtr, typename ValPtr>
  __device__ __forceinline__ float score(const float4* __restrict__ V4,
                                         const float* __restrict__ w, int gp, int ws, bool use_w,
                                         IdPtr ids, ValPtr xs, int beg_, int end_, int c, int s) {
    gather(V4, w, gp, ws, use_w, ids, xs, beg_, end_, c, s);
    return reduce<false>(V4, w, gp, ws, use_w, ids, xs, c, s);
  }

  __device__ __forceinline__ void accumulate(const float4& v, float x, float wv, float& sq,
                                             float& lin) {
    const float dx = v.x * x, dy = v.y * x, dz = v.z * x, dw = v.w * x;
    acc.x += dx;
    acc.y += dy;
    acc.z += dz;
    acc.w += dw;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
    lin += wv * x;
  }
};

}  // namespace fmb

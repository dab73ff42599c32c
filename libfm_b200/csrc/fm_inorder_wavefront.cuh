// fm_inorder_wavefront.cuh -- the wavefront schedule of the sequential-equivalent epoch.
// Included by fm_inorder.cu (compiled with --fmad=false).  Kept in a header of its own so
// that tests/simt/ can compile this exact source for the host and run it as 32 threads
// (one per lane, barriers where the kernel synchronises) against the oracle and under
// ThreadSanitizer.
#pragma once
#include "fmb200_internal.h"

namespace fmb {

// ---------------------------------------------------------------------------------------
// Wavefront schedule of the in-order epoch (short rows, small k).
//
// The row-at-a-time kernel above pays three dependent L2 round trips per example.  The
// sequential semantics only force ONE chain through the epoch: the bias.  Every example
// reads w0 as the FIRST addend of its score (fm_model.h:107-109) and writes it back
// (fm_sgd.h:34-37); because floating-point addition is not associative the whole
// left-to-right accumulation  ((w0 + a_1) + a_2 ...) + b_1 ... + b_k  has to run after
// w0 is known, but the addends themselves -- a_i = w[id_i]*x_i and
// b_f = 0.5*(sum_f^2 - sumsq_f) -- depend only on the rows of w and V the example touches.
// Consecutive examples that share no feature can therefore gather and form their addends
// in parallel, run the scalar chain one after the other, and scatter their updates in
// parallel, with every rounding identical to the sequential loop.
//
// One warp, lane = example.  Per step:
//   1. the next 32 examples load their entries; a shared-memory hash table finds the
//      longest prefix P in which no example touches a feature of an earlier one (hash
//      collisions only shorten P -- conservative);
//   2. lanes < P gather w / V rows (fp64) and form their addends into shared memory;
//   3. all lanes walk the chain t = 0..P-1 redundantly from shared memory (broadcast
//      reads, software-prefetched one example ahead): about nnz + k + 6 dependent fp64
//      operations per example, the only serial part left;
//   4. lanes < P apply fm_SGD to their own rows (cached values; a row that names the
//      same feature twice re-reads memory so the second update sees the first).
// The next step starts at example base + P.
constexpr int WF_Z = 4;        // entries per row
constexpr int WF_K = 8;        // factors
constexpr int WF_HASH = 2048;  // hash slots (power of two)

__device__ __forceinline__ uint32_t wf_hash(uint32_t id) {
  return (id * 2654435761u) >> (32 - 11);
}
static_assert(WF_HASH == (1 << 11), "wf_hash produces 11 bits");

template <bool K0, int TASK>
__global__ void __launch_bounds__(32, 1)
    fm_sgd_inorder_wavefront_kernel(Params64 p, int n_factor, int use_w0, int use_w, HParams hp,
                                    uint64_t n_rows, const uint64_t* __restrict__ row_ptr,
                                    const uint32_t* __restrict__ col,
                                    const float* __restrict__ val,
                                    const float* __restrict__ target) {
  constexpr int NA = WF_Z + WF_K;  // addends per example
  __shared__ __align__(16) double s_add[32][NA];
  __shared__ float s_y[32];
  __shared__ unsigned int s_hash[WF_HASH];  // (step << 5) | (31 - lane): max = current step, lowest lane

  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int k = n_factor;
  constexpr bool k0 = K0;
  const bool k1 = use_w != 0;
  double* w = p.w();
  double* v = p.v();
  double w0 = k0 ? *p.w0() : 0.0;
  const double lr = hp.lr, reg0 = hp.reg0, regw = hp.regw, regv = hp.regv;

  const bool clamp_inverted = hp.max_target < hp.min_target;  // degenerate bounds: still fmin-then-fmax

  for (int i = lane; i < WF_HASH; i += 32) s_hash[i] = 0;
  __syncwarp();

  unsigned int seq = 0;
  uint64_t base = 0;
  // Row bounds, targets and entries of the coming step are fetched one step ahead (the
  // loads are in flight while the chain of the current step runs).
  uint64_t beg = 0, end = 0;
  float yf = 0.f;
  uint32_t id[WF_Z];
  float xf[WF_Z];
  if (lane < n_rows) {
    beg = row_ptr[lane];
    end = row_ptr[lane + 1];
    yf = target[lane];
  }
#pragma unroll
  for (int j = 0; j < WF_Z; j++) {
    id[j] = 0;
    xf[j] = 0.f;
    if (j < (int)(end - beg)) {
      id[j] = col[beg + j];
      xf[j] = val[beg + j];
    }
  }

  while (base < n_rows) {
    const bool valid = base + lane < n_rows;
    const int size = valid ? (int)(end - beg) : 0;
    double x[WF_Z];
#pragma unroll
    for (int j = 0; j < WF_Z; j++) x[j] = (double)xf[j];
    // (1) conflict-free prefix
    bool dup = false;
#pragma unroll
    for (int j = 1; j < WF_Z; j++)
#pragma unroll
      for (int j2 = 0; j2 < j; j2++)
        if (j < size && id[j] == id[j2]) dup = true;
    if (++seq == (1u << 27)) {  // step counter about to leave its 27 bits: start over
      for (int i = lane; i < WF_HASH; i += 32) s_hash[i] = 0;
      seq = 1;
      __syncwarp();
    }
    const unsigned int tag = (seq << 5) | (unsigned int)(31 - lane);
#pragma unroll
    for (int j = 0; j < WF_Z; j++)
      if (j < size) atomicMax(&s_hash[wf_hash(id[j])], tag);
    __syncwarp();
    bool conflict = false;
#pragma unroll
    for (int j = 0; j < WF_Z; j++)
      if (j < size) {
        const unsigned int h = s_hash[wf_hash(id[j])];  // written in this step: (h >> 5) == seq
        if (31 - (int)(h & 31u) < lane) conflict = true;
      }
    const unsigned stop = __ballot_sync(full, conflict || !valid);
    const int P = stop ? __ffs(stop) - 1 : 32;  // >= 1: lane 0 is valid and never in conflict
    const bool active = lane < P;

    // (2) gather + addends
    double wv[WF_Z], vv[WF_Z][WF_K], sum[WF_K];
    if (active) {
#pragma unroll
      for (int j = 0; j < WF_Z; j++) {
        wv[j] = 0;
        if (j < size && k1) wv[j] = w[id[j]];
#pragma unroll
        for (int f = 0; f < WF_K; f++) {
          vv[j][f] = 0;
          if (j < size && f < k) vv[j][f] = v[(size_t)id[j] * k + f];
        }
      }
    }
    // bounds of the next step's examples: in flight during the chain
    const uint64_t nr = base + P + lane;
    uint64_t nbeg = 0, nend = 0;
    float nyf = 0.f;
    if (nr < n_rows) {
      nbeg = row_ptr[nr];
      nend = row_ptr[nr + 1];
      nyf = target[nr];
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < WF_Z; j++) s_add[lane][j] = (j < size && k1) ? wv[j] * x[j] : -0.0;
#pragma unroll
      for (int f = 0; f < WF_K; f++) {
        double sf = 0, ss = 0;
#pragma unroll
        for (int j = 0; j < WF_Z; j++)
          if (j < size) {  // fm_model.h:113-121
            double d = vv[j][f] * x[j];
            sf += d;
            ss += d * d;
          }
        sum[f] = sf;
        s_add[lane][WF_Z + f] = (f < k) ? 0.5 * (sf * sf - ss) : -0.0;
      }
      s_y[lane] = yf;
    }
    __syncwarp();
    // entries of the next step's examples: in flight during the chain
    uint32_t nid[WF_Z];
    float nxf[WF_Z];
#pragma unroll
    for (int j = 0; j < WF_Z; j++) {
      nid[j] = 0;
      nxf[j] = 0.f;
      if (j < (int)(nend - nbeg)) {
        nid[j] = col[nbeg + j];
        nxf[j] = val[nbeg + j];
      }
    }

    // (3) the bias chain, every lane redundantly
    double my_mult = 0;
    double cur[NA], nxt[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) cur[a] = s_add[0][a];
    for (int t = 0; t < P; t++) {
      const int tn = (t + 1 < P) ? t + 1 : t;
#pragma unroll
      for (int a = 0; a < NA; a++) nxt[a] = s_add[tn][a];
      const double y = (double)s_y[t];
      const double m_lo = -(y - hp.min_target), m_hi = -(y - hp.max_target);  // off the chain
      // Unused addend slots hold -0.0, the exact identity of IEEE addition (x + -0.0 == x
      // for every x, signed zeros included): no select sits in the dependent chain.  A
      // model without bias keeps the local w0 at +0.0, so 0.0 + w0 is the reference's
      // `result = 0`.
      double pr = 0.0 + w0;
#pragma unroll
      for (int a = 0; a < NA; a++) pr += cur[a];
      double mult = 0;  // fm_learn_sgd_element.h:58-65
      if (TASK == FMB200_TASK_REGRESSION) {
        // mult = -(y - clamp(pr)).  The two comparisons and the unclamped difference all
        // start from pr at once and one select picks among three candidates (two of them
        // known before the chain), instead of compare -> select -> compare -> select ->
        // subtract in series.  Same values as fmin(max, .) then fmax(min, .): a NaN or
        // too-large score takes max_target, then anything below min_target takes it.
        const bool hi = !(pr <= hp.max_target);
        const bool lo = hi ? clamp_inverted : (pr < hp.min_target);
        const double m_mid = -(y - pr);
        mult = lo ? m_lo : (hi ? m_hi : m_mid);
      } else {
        mult = -y * (1.0 - 1.0 / (1.0 + exp(-y * pr)));
      }
      if (k0) w0 -= lr * (mult + reg0 * w0);  // fm_sgd.h:34-37
      if (lane == t) my_mult = mult;
#pragma unroll
      for (int a = 0; a < NA; a++) cur[a] = nxt[a];
    }

    // (4) scatter: fm_sgd.h:38-50 for the lane's own example
    if (active) {
      if (k1) {
#pragma unroll
        for (int j = 0; j < WF_Z; j++)
          if (j < size) {
            double* wi = &w[id[j]];
            double c = dup ? *wi : wv[j];
            c -= lr * (my_mult * x[j] + regw * c);
            *wi = c;
          }
      }
#pragma unroll
      for (int f = 0; f < WF_K; f++)
        if (f < k) {
#pragma unroll
          for (int j = 0; j < WF_Z; j++)
            if (j < size) {
              double* vp = &v[(size_t)id[j] * k + f];
              double c = dup ? *vp : vv[j][f];
              double grad = sum[f] * x[j] - c * x[j] * x[j];
              c -= lr * (my_mult * grad + regv * c);
              *vp = c;
            }
        }
    }
    __syncwarp();  // the next step's gathers (other lanes) must see these stores
    base += P;
    beg = nbeg;
    end = nend;
    yf = nyf;
#pragma unroll
    for (int j = 0; j < WF_Z; j++) {
      id[j] = nid[j];
      xf[j] = nxf[j];
    }
  }
  if (lane == 0 && k0) *p.w0() = w0;
}

}  // namespace fmb

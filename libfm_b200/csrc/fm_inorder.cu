// fm_inorder.cu -- sequential-equivalent fp64 path (FMB200_MODE_INORDER).
//
// This translation unit is compiled with --fmad=false: the reference is built
// by g++ -O3 for baseline x86-64 (no FMA contraction), and the parity gate for
// this mode is bit-level agreement of w0/w/V with fm_learn_sgd_element::learn
// (reference src/libfm/src/fm_learn_sgd_element.h:56-67) for regression.
//
// Mapping: ONE warp walks the rows strictly in file order (example t must see
// every update of examples < t, including the bias w0 that every example
// touches -- fm_sgd.h:34-37 -- so the epoch is one serial dependency chain and
// no parallel schedule is sequentially equivalent).  Within a row the warp
// parallelises across factors: lane l owns factors l, l+32, ...  All
// floating-point sums are formed in the reference's order:
//   result = w0 ; += w[id_i]*x_i (i ascending) ; += 0.5*(sum_f^2 - sumsq_f) (f ascending)
//   sum_f  = ((0 + d_0) + d_1) + ...                       (fm_model.h:105-127)
// V is attribute-major fp64 [n][k] here (lane-contiguous, coalesced), versus
// the reference's factor-major layout; only the addressing differs.
#include "fm_device.cuh"
#include "fmb200_internal.h"

namespace fmb {

constexpr int KF_MAX = 8;  // factors per lane -> num_factor <= 256 in this mode

struct RowCtx {
  int k;
  bool k0, k1;
  const double* w;
  const double* v;
};

// Exact fm_model::predict for one row, executed by a full warp.  Returns the
// score in every lane; sum[j] holds sum_f for f = lane + 32*j.
// KF = factors per lane (1, 2, 4 or 8): the single warp of the in-order epoch has no
// other warp to hide behind, so dead predicated code for unused factor slots (and the
// instruction-cache misses it causes) is paid in full -- the kernels are instantiated per KF.
template <int KF>
__device__ __forceinline__ double predict_row_exact(const RowCtx& m, double w0,
                                                    const uint32_t* __restrict__ col,
                                                    const float* __restrict__ val, uint32_t size,
                                                    double (&sum)[KF], int lane) {
  // Issue this lane's factor gathers BEFORE lane 0 walks the linear weights: both are
  // L2 round trips, and the warp would otherwise serialise them (lane 0's branch runs
  // first).  Short rows of models with k <= 32 keep the values in registers.
  constexpr int VC = 8;
  const bool cached = size <= (uint32_t)VC && m.k <= 32;
  double vcache[VC];
  if (cached && lane < m.k) {
#pragma unroll
    for (int i = 0; i < VC; i++)
      if ((uint32_t)i < size) vcache[i] = m.v[(size_t)col[i] * m.k + lane];
  }
  double result = 0;
  if (lane == 0) {
    if (m.k0) result += w0;
    if (m.k1) {
      for (uint32_t i = 0; i < size; i++) {
        result += m.w[col[i]] * (double)val[i];
      }
    }
  }
  result = __shfl_sync(0xffffffffu, result, 0);
  double term[KF];
#pragma unroll
  for (int j = 0; j < KF; j++) {
    int f = lane + 32 * j;
    double s = 0, ss = 0;
    if (f < m.k) {
      if (cached) {  // j == 0 only (k <= 32)
#pragma unroll
        for (int i = 0; i < VC; i++) {
          if ((uint32_t)i < size) {
            double d = vcache[i] * (double)val[i];
            s += d;
            ss += d * d;
          }
        }
      } else {
        for (uint32_t i = 0; i < size; i++) {
          double d = m.v[(size_t)col[i] * m.k + f] * (double)val[i];
          s += d;
          ss += d * d;
        }
      }
    }
    sum[j] = s;
    term[j] = 0.5 * (s * s - ss);
  }
  // ordered accumulation over f = 0..k-1 (all lanes redundantly, same ops)
#pragma unroll
  for (int j = 0; j < KF; j++) {
    int fbase = 32 * j;
    if (fbase < m.k) {
      int cnt = min(32, m.k - fbase);
      for (int l = 0; l < cnt; l++) {
        double t = __shfl_sync(0xffffffffu, term[j], l);
        result += t;
      }
    }
  }
  return result;
}

template <int KF>
__global__ void __launch_bounds__(32, 1)
    fm_sgd_inorder_kernel(Params64 p, int n_factor, int use_w0, int use_w, HParams hp,
                          uint64_t n_rows, const uint64_t* __restrict__ row_ptr,
                          const uint32_t* __restrict__ col, const float* __restrict__ val,
                          const float* __restrict__ target) {
  const int lane = threadIdx.x;
  RowCtx m;
  m.k = n_factor;
  m.k0 = use_w0 != 0;
  m.k1 = use_w != 0;
  m.w = p.w();
  m.v = p.v();
  double* w = p.w();
  double* v = p.v();
  double w0 = *p.w0();
  const double lr = hp.lr, reg0 = hp.reg0, regw = hp.regw, regv = hp.regv;
  double sum[KF];

  for (uint64_t r0 = 0; r0 < n_rows; r0 += 32) {
    // the CSR is immutable: fetch 32 rows' bounds and targets at once
    uint64_t rr = r0 + lane;
    uint64_t my_beg = 0, my_end = 0;
    float my_y = 0.f;
    if (rr < n_rows) {
      my_beg = row_ptr[rr];
      my_end = row_ptr[rr + 1];
      my_y = target[rr];
    }
    int cnt = (int)min((uint64_t)32, n_rows - r0);
    for (int q = 0; q < cnt; q++) {
      uint64_t beg = __shfl_sync(0xffffffffu, my_beg, q);
      uint64_t end = __shfl_sync(0xffffffffu, my_end, q);
      double y = (double)__shfl_sync(0xffffffffu, my_y, q);
      uint32_t size = (uint32_t)(end - beg);
      const uint32_t* c = col + beg;
      const float* x = val + beg;

      double pr = predict_row_exact<KF>(m, w0, c, x, size, sum, lane);
      // fm_learn_sgd_element.h:58-65
      double mult = 0;
      if (hp.task == FMB200_TASK_REGRESSION) {
        pr = fmin(hp.max_target, pr);
        pr = fmax(hp.min_target, pr);
        mult = -(y - pr);
      } else {
        mult = -y * (1.0 - 1.0 / (1.0 + exp(-y * pr)));
      }
      // fm_sgd.h:34-37 (replicated in every lane, identical arithmetic)
      if (m.k0) w0 -= lr * (mult + reg0 * w0);
      // fm_sgd.h:38-43 (lane 0 is the only reader/writer of w)
      if (m.k1 && lane == 0) {
        for (uint32_t i = 0; i < size; i++) {
          double* wi = &w[c[i]];
          double cur = *wi;
          cur -= lr * (mult * (double)x[i] + regw * cur);
          *wi = cur;
        }
      }
      // fm_sgd.h:44-50 (lane f%32 is the only reader/writer of V[:, f])
#pragma unroll
      for (int j = 0; j < KF; j++) {
        int f = lane + 32 * j;
        if (f < m.k) {
          for (uint32_t i = 0; i < size; i++) {
            double* vp = &v[(size_t)c[i] * m.k + f];
            double xv = (double)x[i];
            double cur = *vp;
            double grad = sum[j] * xv - cur * xv * xv;
            cur -= lr * (mult * grad + regv * cur);
            *vp = cur;
          }
        }
      }
    }
  }
  if (lane == 0 && m.k0) *p.w0() = w0;
}

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

// Exact fp64 scores: one warp per row; optional metric partials per block in
// fixed (deterministic) order: each block reduces its warps in warp order.
template <int KF>
__global__ void __launch_bounds__(256)
    fm_predict64_kernel(Params64 p, int n_factor, int use_w0, int use_w, HParams hp, int transform,
                        uint64_t n_rows, const uint64_t* __restrict__ row_ptr,
                        const uint32_t* __restrict__ col, const float* __restrict__ val,
                        const float* __restrict__ target, double* __restrict__ out_pred,
                        double* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  RowCtx m;
  m.k = n_factor;
  m.k0 = use_w0 != 0;
  m.k1 = use_w != 0;
  m.w = p.w();
  m.v = p.v();
  const double w0 = *p.w0();
  double sum[KF];
  double sq = 0, ab = 0, ok = 0;
  // contiguous row ranges per warp keep the reduction order a pure function of
  // (n_rows, grid, block)
  uint64_t total_warps = (uint64_t)gridDim.x * nwarp;
  uint64_t gw = (uint64_t)blockIdx.x * nwarp + warp;
  uint64_t per = (n_rows + total_warps - 1) / total_warps;
  uint64_t rbeg = gw * per, rend = min(n_rows, rbeg + per);
  for (uint64_t r = rbeg; r < rend; r++) {
    uint64_t beg = row_ptr[r], end = row_ptr[r + 1];
    double pr = predict_row_exact<KF>(m, w0, col + beg, val + beg, (uint32_t)(end - beg), sum, lane);
    double y = (double)target[r];
    if (hp.task == FMB200_TASK_REGRESSION) {
      // fm_learn.h:138-142
      double pc = fmin(hp.max_target, pr);
      pc = fmax(hp.min_target, pc);
      double err = pc - y;
      sq += err * err;
      ab += fabs(err);
      if (transform) pr = pc;
    } else {
      // fm_learn.h:118-120
      if (((pr >= 0) && (y >= 0)) || ((pr < 0) && (y < 0))) ok += 1;
      if (transform) pr = 1.0 / (1.0 + exp(-pr));  // fm_learn_sgd.h:84
    }
    if (out_pred != nullptr && lane == 0) out_pred[r] = pr;
  }
  if (partials != nullptr) {
    __shared__ double s_part[8][3];
    if (lane == 0) {
      s_part[warp][0] = sq;
      s_part[warp][1] = ab;
      s_part[warp][2] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0, c = 0;
      for (int i = 0; i < nwarp; i++) {
        a += s_part[i][0];
        b += s_part[i][1];
        c += s_part[i][2];
      }
      partials[3 * blockIdx.x + 0] = a;
      partials[3 * blockIdx.x + 1] = b;
      partials[3 * blockIdx.x + 2] = c;
    }
  }
}

cudaError_t launch_sgd_inorder(fmb200_ctx* c, const DataSlot& d) {
  if (c->k > 32 * KF_MAX) return cudaErrorInvalidValue;
  // Opt-in (fmb200_set_tuning variant 4) until it has been measured on the device.
  if (c->tune_variant == 4 && c->k <= WF_K && d.max_row_nnz <= (uint32_t)WF_Z && d.n_rows > 0) {
#define FMB_WAVEFRONT(K0, TASK)                                                                       \
  fm_sgd_inorder_wavefront_kernel<K0, TASK><<<1, 32, 0, c->stream>>>(c->p64, c->k, c->k0, c->k1, c->hp, \
                                                                     d.n_rows, d.row_ptr, d.col, d.val, \
                                                                     d.target)
    const bool reg = c->hp.task == FMB200_TASK_REGRESSION;
    if (c->k0 && reg) FMB_WAVEFRONT(true, FMB200_TASK_REGRESSION);
    else if (c->k0) FMB_WAVEFRONT(true, FMB200_TASK_CLASSIFICATION);
    else if (reg) FMB_WAVEFRONT(false, FMB200_TASK_REGRESSION);
    else FMB_WAVEFRONT(false, FMB200_TASK_CLASSIFICATION);
#undef FMB_WAVEFRONT
    c->launches++;
    c->last_cfg = EpochConfig{32, WF_Z, 32, 1, 32, 0};
    return cudaGetLastError();
  }
  const int kf = (c->k + 31) / 32;
#define FMB_INORDER(KF)                                                                         \
  fm_sgd_inorder_kernel<KF><<<1, 32, 0, c->stream>>>(c->p64, c->k, c->k0, c->k1, c->hp, d.n_rows, \
                                                     d.row_ptr, d.col, d.val, d.target)
  if (kf <= 1) FMB_INORDER(1);
  else if (kf <= 2) FMB_INORDER(2);
  else if (kf <= 4) FMB_INORDER(4);
  else FMB_INORDER(8);
#undef FMB_INORDER
  c->launches++;
  c->last_cfg = EpochConfig{32, 1, 1, 1, 32, 0};
  return cudaGetLastError();
}

cudaError_t launch_predict64(fmb200_ctx* c, const DataSlot& d, int transform, double* out_pred,
                             double* partials, int n_blocks) {
  if (c->k > 32 * KF_MAX) return cudaErrorInvalidValue;
  const int kf = (c->k + 31) / 32;
#define FMB_PREDICT64(KF)                                                                          \
  fm_predict64_kernel<KF><<<n_blocks, 256, 0, c->stream>>>(c->p64, c->k, c->k0, c->k1, c->hp,      \
                                                           transform, d.n_rows, d.row_ptr, d.col, \
                                                           d.val, d.target, out_pred, partials)
  if (kf <= 1) FMB_PREDICT64(1);
  else if (kf <= 2) FMB_PREDICT64(2);
  else if (kf <= 4) FMB_PREDICT64(4);
  else FMB_PREDICT64(8);
#undef FMB_PREDICT64
  c->launches++;
  return cudaGetLastError();
}

}  // namespace fmb

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
#include "fm_inorder_wavefront.cuh"

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
      if (transform == 1) pr = pc;
      if (transform == 2) pr = err;  // evaluate in row order on the host (fm_learn.h:139)
    } else {
      // fm_learn.h:118-120
      if (((pr >= 0) && (y >= 0)) || ((pr < 0) && (y < 0))) ok += 1;
      if (transform == 1) pr = 1.0 / (1.0 + exp(-pr));  // fm_learn_sgd.h:84
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

// MCMC / ALS e-term pass (reference libfm/src/fm_learn_mcmc.h:148-378, no relations): the
// learner re-predicts every case once per iteration through the TRANSPOSED copy of the data, so
// each case accumulates its terms in ascending feature id (ties in row order) and in a different
// association than fm_model::predict:
//   e = sum_f 0.5 q_f^2 ;  q = sum_f sum_i -0.5 v_if^2 x_i^2  (+ sum_i w_i x_i) ;  e = (e + q) + w0
// One thread per case, every operation in that order (this TU is compiled with --fmad=false):
// bit-identical e-terms.  Rows whose ids are not ascending are visited through a per-thread
// order array (rows of <= ET_LOCAL entries) or by repeated selection (longer rows).
constexpr int ET_LOCAL = 64;

struct RowOrder {
  const uint32_t* c;
  uint32_t size;
  bool sorted;
  unsigned short ord[ET_LOCAL];
  __device__ __forceinline__ void init(const uint32_t* col, uint32_t n) {
    c = col;
    size = n;
    sorted = true;
    for (uint32_t i = 1; i < n; i++)
      if (col[i] < col[i - 1]) sorted = false;
    if (!sorted && n <= (uint32_t)ET_LOCAL) {  // stable insertion sort by id
      for (uint32_t i = 0; i < n; i++) ord[i] = (unsigned short)i;
      for (uint32_t i = 1; i < n; i++) {
        const unsigned short o = ord[i];
        uint32_t j = i;
        while (j > 0 && col[ord[j - 1]] > col[o]) {
          ord[j] = ord[j - 1];
          j--;
        }
        ord[j] = o;
      }
    }
  }
  // position of the i-th entry in (id, position) order; `prev` = position of the (i-1)-th
  __device__ __forceinline__ uint32_t at(uint32_t i, uint32_t prev) const {
    if (sorted) return i;
    if (size <= (uint32_t)ET_LOCAL) return ord[i];
    // selection: the smallest (id, position) greater than (c[prev], prev)
    uint32_t best = 0xffffffffu;
    for (uint32_t j = 0; j < size; j++) {
      const bool after = (i == 0) || c[j] > c[prev] || (c[j] == c[prev] && j > prev);
      if (!after) continue;
      if (best == 0xffffffffu || c[j] < c[best]) best = j;
    }
    return best;
  }
};

__global__ void __launch_bounds__(128)
    fm_eterm64_kernel(Params64 p, int k, int use_w0, int use_w, uint64_t n_rows,
                      const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                      const float* __restrict__ val, double* __restrict__ e_out) {
  const double* w = p.w();
  const double* v = p.v();
  const double w0 = *p.w0();
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t beg = row_ptr[r];
    const uint32_t size = (uint32_t)(row_ptr[r + 1] - beg);
    const uint32_t* c = col + beg;
    const float* x = val + beg;
    RowOrder o;
    o.init(c, size);
    double e = 0.0, q = 0.0;
    for (int f = 0; f < k; f++) {  // fm_learn_mcmc.h:172-252
      q = 0.0;
      uint32_t pos = 0;
      for (uint32_t i = 0; i < size; i++) {
        pos = o.at(i, pos);
        q += v[(size_t)c[pos] * k + f] * (double)x[pos];
      }
      e += 0.5 * q * q;
    }
    q = 0.0;
    for (int f = 0; f < k; f++) {  // :255-306
      uint32_t pos = 0;
      for (uint32_t i = 0; i < size; i++) {
        pos = o.at(i, pos);
        const double vif = v[(size_t)c[pos] * k + f];
        const float xi = x[pos];
        q -= 0.5 * vif * vif * xi * xi;  // (((0.5*v)*v)*x)*x, x promoted to double per factor
      }
    }
    if (use_w) {  // :309-346
      uint32_t pos = 0;
      for (uint32_t i = 0; i < size; i++) {
        pos = o.at(i, pos);
        q += w[c[pos]] * (double)x[pos];
      }
    }
    e = e + q;  // :350-362
    if (use_w0) e += w0;
    e_out[r] = e;
  }
}

cudaError_t launch_mcmc_eterms(fmb200_ctx* c, const DataSlot& d, double* e_out) {
  if (d.n_rows == 0) return cudaSuccess;
  const uint64_t blocks = (d.n_rows + 127) / 128;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  fm_eterm64_kernel<<<grid, 128, 0, c->stream>>>(c->p64, c->k, c->k0, c->k1, d.n_rows, d.row_ptr, d.col, d.val,
                                                 e_out);
  c->launches++;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// SGDA: SGD with self-adaptive regularisation (reference
// libfm/src/fm_learn_sgd_element_adapt_reg.h).  Every training row's theta-step (:136-169) is
// followed by a lambda-step on the next validation row (:201-248) that moves the per-group
// regularisation values; both touch state every later step reads (w0, reg_w, reg_v), so the
// epoch is one chain like the plain in-order epoch and runs the same way: one warp, lane =
// factor, every operation in the reference's order (--fmad=false): bit-identical parameters and
// regularisation values.  Per-group accumulators of a lambda-step live in shared memory
// ([group][factor], one column per lane: no two lanes share a word).
struct SgdaArgs {
  Params64 p;
  double* grad_w;        // [n]
  double* grad_v;        // [n][k] attribute-major
  double* reg_w;         // [G]
  double* reg_v;         // [G][k]
  const uint32_t* group; // [n]
  uint32_t n_groups;
  int k, use_w0, use_w, lambda_steps;
  HParams hp;
  uint64_t n_rows, v_rows;
  const uint64_t *row_ptr, *v_row_ptr;
  const uint32_t *col, *v_col;
  const float *val, *v_val, *target, *v_target;
};

template <int KF>
__global__ void __launch_bounds__(32, 1) fm_sgda_epoch_kernel(const SgdaArgs a) {
  extern __shared__ double sg_smem[];  // reg_w[G] | reg_v[G][k] | sum_f[G][k] | sum_f_dash_f[G][k] | lwg[G]
  const int lane = threadIdx.x;
  const int k = a.k;
  const uint32_t G = a.n_groups;
  double* s_reg_w = sg_smem;
  double* s_reg_v = s_reg_w + G;
  double* s_sum_f = s_reg_v + (size_t)G * k;
  double* s_sdf = s_sum_f + (size_t)G * k;
  double* s_lwg = s_sdf + (size_t)G * k;
  for (uint32_t i = lane; i < G; i += 32) s_reg_w[i] = a.reg_w[i];
  for (uint32_t i = lane; i < G * (uint32_t)k; i += 32) s_reg_v[i] = a.reg_v[i];
  __syncwarp();
  RowCtx m;
  m.k = k;
  m.k0 = a.use_w0 != 0;
  m.k1 = a.use_w != 0;
  m.w = a.p.w();
  m.v = a.p.v();
  double* w = a.p.w();
  double* v = a.p.v();
  double w0 = *a.p.w0();
  const double lr = a.hp.lr;
  const unsigned full = 0xffffffffu;
  double sum[KF];
  uint64_t vc = 0;
  for (uint64_t r = 0; r < a.n_rows; r++) {
    {  // ---- sgd_theta_step, :136-169 ----
      const uint64_t beg = a.row_ptr[r];
      const uint32_t size = (uint32_t)(a.row_ptr[r + 1] - beg);
      const uint32_t* c = a.col + beg;
      const float* x = a.val + beg;
      const float target = a.target[r];
      double p = predict_row_exact<KF>(m, w0, c, x, size, sum, lane);
      double mult = 0;
      if (a.hp.task == FMB200_TASK_REGRESSION) {
        p = fmin(a.hp.max_target, p);
        p = fmax(a.hp.min_target, p);
        mult = 2 * (p - target);
      } else {
        mult = target * ((1.0 / (1.0 + exp(-target * p))) - 1.0);
      }
      if (m.k0) w0 -= lr * (mult + 2 * 0.0 * w0);  // reg_0 stays 0 (:60,79)
      if (m.k1 && lane == 0) {
        for (uint32_t i = 0; i < size; i++) {
          const uint32_t id = c[i];
          const uint32_t g = a.group[id];
          double cur = w[id];
          const double gw = mult * x[i];
          a.grad_w[id] = gw;
          cur -= lr * (gw + 2 * s_reg_w[g] * cur);
          w[id] = cur;
        }
      }
#pragma unroll
      for (int j = 0; j < KF; j++) {
        const int f = lane + 32 * j;
        if (f < k) {
          for (uint32_t i = 0; i < size; i++) {
            const uint32_t id = c[i];
            const uint32_t g = a.group[id];
            double* vp = &v[(size_t)id * k + f];
            double cur = *vp;
            const double gv = mult * (x[i] * (sum[j] - cur * x[i]));
            a.grad_v[(size_t)id * k + f] = gv;
            cur -= lr * (gv + 2 * s_reg_v[(size_t)g * k + f] * cur);
            *vp = cur;
          }
        }
      }
      __syncwarp();
    }
    if (a.lambda_steps && a.v_rows > 0) {  // ---- sgd_lambda_step, :201-248 ----
      if (vc == a.v_rows) vc = 0;  // :302-305
      const uint64_t beg = a.v_row_ptr[vc];
      const uint32_t size = (uint32_t)(a.v_row_ptr[vc + 1] - beg);
      const uint32_t* c = a.v_col + beg;
      const float* x = a.v_val + beg;
      const float target = a.v_target[vc];
      vc++;
      // predict_scaled, :171-199
      double p = 0.0;
      if (lane == 0) {
        if (m.k0) p += w0;
        if (m.k1)
          for (uint32_t i = 0; i < size; i++) {
            const uint32_t id = c[i];
            const double wv = w[id];
            const double w_dash = wv - lr * (a.grad_w[id] + 2 * s_reg_w[a.group[id]] * wv);
            p += w_dash * x[i];
          }
      }
      p = __shfl_sync(full, p, 0);
      double term[KF];
#pragma unroll
      for (int j = 0; j < KF; j++) {
        const int f = lane + 32 * j;
        double s = 0.0, ss = 0.0;
        if (f < k)
          for (uint32_t i = 0; i < size; i++) {
            const uint32_t id = c[i];
            const double vv = v[(size_t)id * k + f];
            const double v_dash =
                vv - lr * (a.grad_v[(size_t)id * k + f] + 2 * s_reg_v[(size_t)a.group[id] * k + f] * vv);
            const double d = v_dash * x[i];
            s += d;
            ss += d * d;
          }
        term[j] = 0.5 * (s * s - ss);
      }
#pragma unroll
      for (int j = 0; j < KF; j++) {
        const int fbase = 32 * j;
        if (fbase < k) {
          const int cnt = min(32, k - fbase);
          for (int l = 0; l < cnt; l++) p += __shfl_sync(full, term[j], l);
        }
      }
      double grad_loss = 0;
      if (a.hp.task == FMB200_TASK_REGRESSION) {
        p = fmin(a.hp.max_target, p);
        p = fmax(a.hp.min_target, p);
        grad_loss = 2 * (p - target);
      } else {
        grad_loss = target * ((1.0 / (1.0 + exp(-target * p))) - 1.0);
      }
      if (m.k1) {  // :213-223: lane g owns group g, g+32, ...
        for (uint32_t g = lane; g < G; g += 32) {
          double acc = 0.0;
          for (uint32_t i = 0; i < size; i++)
            if (a.group[c[i]] == g) acc += x[i] * w[c[i]];
          acc = -2 * lr * acc;
          double rw = s_reg_w[g] - lr * grad_loss * acc;
          s_reg_w[g] = (0.0 < rw) ? rw : 0.0;  // std::max(0.0, .)
        }
      }
#pragma unroll
      for (int j = 0; j < KF; j++) {  // :224-247
        const int f = lane + 32 * j;
        if (f < k) {
          double sum_f_dash = 0.0;
          for (uint32_t g = 0; g < G; g++) {
            s_sum_f[(size_t)g * k + f] = 0.0;
            s_sdf[(size_t)g * k + f] = 0.0;
          }
          for (uint32_t i = 0; i < size; i++) {
            const uint32_t id = c[i];
            const uint32_t g = a.group[id];
            const double vv = v[(size_t)id * k + f];
            const double v_dash = vv - lr * (a.grad_v[(size_t)id * k + f] + 2 * s_reg_v[(size_t)g * k + f] * vv);
            sum_f_dash += v_dash * x[i];
            s_sum_f[(size_t)g * k + f] += vv * x[i];
            s_sdf[(size_t)g * k + f] += v_dash * x[i] * vv * x[i];
          }
          for (uint32_t g = 0; g < G; g++) {
            const double lvg = -2 * lr * (sum_f_dash * s_sum_f[(size_t)g * k + f] - s_sdf[(size_t)g * k + f]);
            const double rv = s_reg_v[(size_t)g * k + f] - lr * grad_loss * lvg;
            s_reg_v[(size_t)g * k + f] = (0.0 < rv) ? rv : 0.0;
          }
        }
      }
      __syncwarp();
    }
  }
  (void)s_lwg;
  if (lane == 0 && m.k0) *a.p.w0() = w0;
  for (uint32_t i = lane; i < G; i += 32) a.reg_w[i] = s_reg_w[i];
  for (uint32_t i = lane; i < G * (uint32_t)k; i += 32) a.reg_v[i] = s_reg_v[i];
}

cudaError_t launch_sgda_epoch(fmb200_ctx* c, const DataSlot& tr, const DataSlot& va, int lambda_steps) {
  if (c->k > 32 * KF_MAX) return cudaErrorInvalidValue;
  SgdaArgs a;
  a.p = c->p64;
  a.grad_w = c->sgda_grad_w;
  a.grad_v = c->sgda_grad_v;
  a.reg_w = c->sgda_reg_w;
  a.reg_v = c->sgda_reg_v;
  a.group = c->sgda_group;
  a.n_groups = c->sgda_groups;
  a.k = c->k;
  a.use_w0 = c->k0;
  a.use_w = c->k1;
  a.lambda_steps = lambda_steps;
  a.hp = c->hp;
  a.n_rows = tr.n_rows;
  a.row_ptr = tr.row_ptr;
  a.col = tr.col;
  a.val = tr.val;
  a.target = tr.target;
  a.v_rows = va.n_rows;
  a.v_row_ptr = va.row_ptr;
  a.v_col = va.col;
  a.v_val = va.val;
  a.v_target = va.target;
  const size_t smem = sizeof(double) * ((size_t)c->sgda_groups * (2 + 3 * (size_t)c->k));
  if (smem > (size_t)c->max_smem_optin) return cudaErrorInvalidConfiguration;
  const int kf = (c->k + 31) / 32;
#define FMB_SGDA(KF)                                                                                  \
  do {                                                                                                \
    cudaError_t e_ = cudaFuncSetAttribute(fm_sgda_epoch_kernel<KF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem);                                                 \
    if (e_ != cudaSuccess) return e_;                                                                 \
    fm_sgda_epoch_kernel<KF><<<1, 32, smem, c->stream>>>(a);                                          \
  } while (0)
  if (kf <= 1) FMB_SGDA(1);
  else if (kf <= 2) FMB_SGDA(2);
  else if (kf <= 4) FMB_SGDA(4);
  else FMB_SGDA(8);
#undef FMB_SGDA
  c->launches++;
  c->last_cfg = EpochConfig{32, 1, 1, 1, 32, (int)smem, 0};
  return cudaGetLastError();
}

cudaError_t launch_sgd_inorder(fmb200_ctx* c, const DataSlot& d) {
  if (c->k > 32 * KF_MAX) return cudaErrorInvalidValue;
  // The wavefront schedule (k <= 8, rows of <= 4 entries) is the default for eligible shapes:
  // bit-identical to the row-at-a-time kernel on the device (tests/test_wavefront_gpu.py, r02) and
  // 6.7x faster on C2 (0.210 s vs 1.416 s per epoch).  variant 1 forces the row-at-a-time kernel.
  if (c->tune_variant != 1 && c->k <= WF_K && d.max_row_nnz <= (uint32_t)WF_Z && d.n_rows > 0) {
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

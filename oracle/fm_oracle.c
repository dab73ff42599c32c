/* oracle/fm_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, "port" oracle).
 *
 * A plain-C restatement of the reference's (srendle/libfm) SGD hot path.  It is
 * the checker the parity tests, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg compare the CUDA path against.  Nothing under libfm_b200/
 * may include, link, dlopen or execute it.
 *
 * PINNING: this file is validated against the reference itself (compiled in
 * place as oracle/_ref/libfm_ref.so by oracle/Makefile) by tests/test_oracle.py
 * -- bit-exact on parameters after several epochs -- and against the golden
 * vectors in tests/golden/ that scripts/make_golden.py generated from the
 * reference.  The reference ships no tests/golden vectors of its own
 * (SURVEY.md section 4), so running it is the only pin available.
 *
 * Arithmetic contract (matches g++ -O3 on x86-64 SSE2: no FMA contraction, IEEE
 * double, float inputs promoted to double): compile with -ffp-contract=off.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/src).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- RNG: util/random.h:148-174 ---------------------------------------- */

void fmo_srand(long seed) { srand((unsigned)seed); } /* libfm.cpp:115-116 */

/* util/random.h:172-174 : rand()/(RAND_MAX+1.0) */
double fmo_ran_uniform(void) { return rand() / ((double)RAND_MAX + 1); }

/* util/random.h:148-162 : Leva's ratio-of-uniforms normal generator */
double fmo_ran_gaussian(void) {
  double u, v, x, y, Q;
  do {
    do {
      u = fmo_ran_uniform();
    } while (u == 0.0);
    v = 1.7156 * (fmo_ran_uniform() - 0.5);
    x = u - 0.449871;
    y = fabs(v) + 0.386595;
    Q = x * x + y * (0.19600 * y - 0.25472 * x);
    if (Q < 0.27597) break;
  } while ((Q > 0.27846) || ((v * v) > (-4.0 * u * u * log(u))));
  return v / u;
}

/* util/random.h:164-170 */
static double ran_gaussian_ms(double mean, double stdev) {
  if ((stdev == 0.0) || isnan(stdev)) return mean;
  return mean + stdev * fmo_ran_gaussian();
}

/* fm_model::init, fm_core/fm_model.h:91-99 with util/matrix.h:398-404:
 * w0 = 0, w = 0, v(f,i) ~ N(mean, stdev) drawn f-outer / i-inner.
 * v is FACTOR-MAJOR [k][n] (util/matrix.h:152-175). */
void fmo_init(uint32_t n, int k, double mean, double stdev, double* w0, double* w, double* v) {
  *w0 = 0;
  for (uint32_t i = 0; i < n; i++) w[i] = 0;
  for (int f = 0; f < k; f++)
    for (uint32_t i = 0; i < n; i++) v[(size_t)f * n + i] = ran_gaussian_ms(mean, stdev);
}

/* ---- predict: fm_core/fm_model.h:105-127 ------------------------------- */
double fmo_predict_row(uint32_t n, int k, int k0, int k1, double w0, const double* w,
                       const double* v, uint32_t size, const uint32_t* col, const float* val,
                       double* sum, double* sum_sqr) {
  double result = 0;
  if (k0) result += w0;
  if (k1)
    for (uint32_t i = 0; i < size; i++) result += w[col[i]] * val[i];
  for (int f = 0; f < k; f++) {
    sum[f] = 0;
    sum_sqr[f] = 0;
    for (uint32_t i = 0; i < size; i++) {
      double d = v[(size_t)f * n + col[i]] * val[i];
      sum[f] += d;
      sum_sqr[f] += d * d;
    }
    result += 0.5 * (sum[f] * sum[f] - sum_sqr[f]);
  }
  return result;
}

/* ---- fm_SGD: fm_core/fm_sgd.h:33-51 ------------------------------------ */
void fmo_sgd_row(uint32_t n, int k, int k0, int k1, double* w0, double* w, double* v, double lr,
                 double reg0, double regw, double regv, uint32_t size, const uint32_t* col,
                 const float* val, double mult, const double* sum) {
  if (k0) *w0 -= lr * (mult + reg0 * *w0);
  if (k1)
    for (uint32_t i = 0; i < size; i++) {
      double* wi = &w[col[i]];
      *wi -= lr * (mult * val[i] + regw * *wi);
    }
  for (int f = 0; f < k; f++)
    for (uint32_t i = 0; i < size; i++) {
      double* vp = &v[(size_t)f * n + col[i]];
      double grad = sum[f] * val[i] - *vp * val[i] * val[i];
      *vp -= lr * (mult * grad + regv * *vp);
    }
}

/* ---- one epoch: libfm/src/fm_learn_sgd_element.h:56-67 ------------------
 * task 0: p = clamp(p); mult = -(y - p)
 * task 1: mult = -y * (1 - 1/(1+exp(-y p)))                                 */
void fmo_sgd_epoch(uint32_t n, int k, int k0, int k1, double* w0, double* w, double* v, double lr,
                   double reg0, double regw, double regv, int task, double min_target,
                   double max_target, uint64_t n_rows, const uint64_t* row_ptr,
                   const uint32_t* col, const float* val, const float* target) {
  double* sum = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double* sum_sqr = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  for (uint64_t r = 0; r < n_rows; r++) {
    uint32_t size = (uint32_t)(row_ptr[r + 1] - row_ptr[r]);
    const uint32_t* c = col + row_ptr[r];
    const float* x = val + row_ptr[r];
    double p = fmo_predict_row(n, k, k0, k1, *w0, w, v, size, c, x, sum, sum_sqr);
    double mult = 0;
    if (task == 0) {
      p = fmin(max_target, p);
      p = fmax(min_target, p);
      mult = -(target[r] - p);
    } else if (task == 1) {
      mult = -target[r] * (1.0 - 1.0 / (1.0 + exp(-target[r] * p)));
    }
    fmo_sgd_row(n, k, k0, k1, w0, w, v, lr, reg0, regw, regv, size, c, x, mult, sum);
  }
  free(sum);
  free(sum_sqr);
}

/* ---- evaluate: libfm/src/fm_learn.h:113-153 -----------------------------
 * task 0 -> *sum_sq_err, *sum_abs_err of clamp(p) - y ; task 1 -> *n_correct */
void fmo_evaluate(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                  int task, double min_target, double max_target, uint64_t n_rows,
                  const uint64_t* row_ptr, const uint32_t* col, const float* val,
                  const float* target, double* sum_sq_err, double* sum_abs_err,
                  uint64_t* n_correct) {
  double* sum = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double* sum_sqr = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double sq = 0, ab = 0;
  uint64_t ok = 0;
  for (uint64_t r = 0; r < n_rows; r++) {
    uint32_t size = (uint32_t)(row_ptr[r + 1] - row_ptr[r]);
    double p = fmo_predict_row(n, k, k0, k1, w0, w, v, size, col + row_ptr[r], val + row_ptr[r],
                               sum, sum_sqr);
    if (task == 0) {
      p = fmin(max_target, p);
      p = fmax(min_target, p);
      double err = p - target[r];
      sq += err * err;
      ab += fabs(err);
    } else {
      if (((p >= 0) && (target[r] >= 0)) || ((p < 0) && (target[r] < 0))) ok++;
    }
  }
  *sum_sq_err = sq;
  *sum_abs_err = ab;
  *n_correct = ok;
  free(sum);
  free(sum_sqr);
}

/* ---- predict output: libfm/src/fm_learn_sgd.h:76-90 ---------------------
 * mode 0: raw score; mode 1: task transform (clamp / sigmoid) as -out writes */
void fmo_predict(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                 int task, double min_target, double max_target, int transform, uint64_t n_rows,
                 const uint64_t* row_ptr, const uint32_t* col, const float* val, double* out) {
  double* sum = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double* sum_sqr = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  for (uint64_t r = 0; r < n_rows; r++) {
    uint32_t size = (uint32_t)(row_ptr[r + 1] - row_ptr[r]);
    double p = fmo_predict_row(n, k, k0, k1, w0, w, v, size, col + row_ptr[r], val + row_ptr[r],
                               sum, sum_sqr);
    if (transform) {
      if (task == 0) {
        p = fmin(max_target, p);
        p = fmax(min_target, p);
      } else {
        p = 1.0 / (1.0 + exp(-p));
      }
    }
    out[r] = p;
  }
  free(sum);
  free(sum_sqr);
}

/* ---- boundary layout work (bit-exact contract) --------------------------
 * AoS sparse_entry{uint id; float value} (util/fmatrix.h:34-37) -> SoA       */
void fmo_aos_to_soa(uint64_t nnz, const void* entries, uint32_t* col, float* val) {
  const unsigned char* p = (const unsigned char*)entries;
  for (uint64_t j = 0; j < nnz; j++) {
    memcpy(&col[j], p + 8 * j, 4);
    memcpy(&val[j], p + 8 * j + 4, 4);
  }
}

/* factor-major double [k][n] -> attribute-major float [n][kp] (zero padded) */
void fmo_v_to_device_layout(uint32_t n, int k, int kp, const double* v, float* out) {
  for (uint32_t i = 0; i < n; i++)
    for (int f = 0; f < kp; f++) out[(size_t)i * kp + f] = f < k ? (float)v[(size_t)f * n + i] : 0.f;
}

/* ---- MCMC / ALS e-term pass: libfm/src/fm_learn_mcmc.h:148-378 (no relations) -------------
 * The learner re-predicts every case once per iteration THROUGH THE TRANSPOSED COPY of the data
 * (Data::create_data_t, Data.h:292-337: feature i lists its cases in ascending case order, and a
 * case's own entries in their row order).  Seen from one case c, the terms therefore arrive in
 * ascending feature id, ties in row order -- not in the row's stored order:
 *   (1) :172-252  for f: q = sum_i v_if x_i ; e += 0.5 q q
 *   (2) :255-306  for f: for i: q -= 0.5 v_if v_if x_i x_i          (one running q over all f)
 *   (3) :309-346  for i: q += w_i x_i                               (if k1)
 *       :350-362  e = e + q ; if k0: e += w0
 * `order` (scratch, max_row_nnz entries) receives the row's entries sorted by (id, position). */
static void sort_row(uint32_t size, const uint32_t* col, uint32_t* order) {
  for (uint32_t i = 0; i < size; i++) order[i] = i;
  for (uint32_t i = 1; i < size; i++) { /* stable insertion sort by id */
    uint32_t o = order[i];
    uint32_t j = i;
    while (j > 0 && col[order[j - 1]] > col[o]) {
      order[j] = order[j - 1];
      j--;
    }
    order[j] = o;
  }
}

void fmo_mcmc_eterms(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                     uint64_t n_rows, const uint64_t* row_ptr, const uint32_t* col, const float* val,
                     double* e_out) {
  uint32_t max_size = 1;
  for (uint64_t r = 0; r < n_rows; r++) {
    uint64_t s = row_ptr[r + 1] - row_ptr[r];
    if (s > max_size) max_size = (uint32_t)s;
  }
  uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * max_size);
  for (uint64_t r = 0; r < n_rows; r++) {
    const uint32_t size = (uint32_t)(row_ptr[r + 1] - row_ptr[r]);
    const uint32_t* c = col + row_ptr[r];
    const float* x = val + row_ptr[r];
    sort_row(size, c, order);
    double e = 0.0, q = 0.0;
    for (int f = 0; f < k; f++) {
      q = 0.0;
      for (uint32_t i = 0; i < size; i++) q += v[(size_t)f * n + c[order[i]]] * x[order[i]];
      e += 0.5 * q * q;
    }
    q = 0.0;
    for (int f = 0; f < k; f++)
      for (uint32_t i = 0; i < size; i++) {
        const double vif = v[(size_t)f * n + c[order[i]]];
        const float xi = x[order[i]];
        q -= 0.5 * vif * vif * xi * xi;
      }
    if (k1)
      for (uint32_t i = 0; i < size; i++) q += w[c[order[i]]] * x[order[i]];
    e = e + q;
    if (k0) e += w0;
    e_out[r] = e;
  }
  free(order);
}

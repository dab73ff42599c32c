/* oracle/fm_oracle_sgda.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, "port" oracle).
 *
 * SGDA: SGD with self-adaptive regularisation, reference
 * libfm/src/fm_learn_sgd_element_adapt_reg.h (Rendle, WSDM 2012).  One epoch interleaves, row for
 * row, a theta-step on a training row (:136-169) with a lambda-step on a validation row (:201-248;
 * skipped in the first epoch, :301), wrapping the validation cursor (:302-305).
 * Pinned bit-exact to the reference's own learner by tests/test_oracle.py (via oracle/_ref).
 * Arithmetic contract as fm_oracle.c: compile with -ffp-contract=off.
 * v, grad_v are FACTOR-MAJOR [k][n]; reg_v is [num_groups][k] (util/matrix.h DMatrix row-major).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

double fmo_predict_row(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                       uint32_t size, const uint32_t* col, const float* val, double* sum, double* sum_sqr);

/* sgd_theta_step, :136-169 */
static void theta_step(uint32_t n, int k, int k0, int k1, double* w0, double* w, double* v, double* grad_w,
                       double* grad_v, const double* reg_w, const double* reg_v, const uint32_t* group, double lr,
                       int task, double min_target, double max_target, uint32_t size, const uint32_t* col,
                       const float* val, float target, double* sum, double* sum_sqr) {
  double p = fmo_predict_row(n, k, k0, k1, *w0, w, v, size, col, val, sum, sum_sqr);
  double mult = 0;
  if (task == 0) {
    p = fmin(max_target, p);
    p = fmax(min_target, p);
    mult = 2 * (p - target);
  } else if (task == 1) {
    mult = target * ((1.0 / (1.0 + exp(-target * p))) - 1.0);
  }
  if (k0) {
    double grad_0 = mult;
    *w0 -= lr * (grad_0 + 2 * 0.0 * *w0); /* reg_0 == 0 always (:60,79) */
  }
  if (k1)
    for (uint32_t i = 0; i < size; i++) {
      uint32_t g = group[col[i]];
      double* wi = &w[col[i]];
      grad_w[col[i]] = mult * val[i];
      *wi -= lr * (grad_w[col[i]] + 2 * reg_w[g] * *wi);
    }
  for (int f = 0; f < k; f++)
    for (uint32_t i = 0; i < size; i++) {
      uint32_t g = group[col[i]];
      double* vp = &v[(size_t)f * n + col[i]];
      grad_v[(size_t)f * n + col[i]] = mult * (val[i] * (sum[f] - *vp * val[i]));
      *vp -= lr * (grad_v[(size_t)f * n + col[i]] + 2 * reg_v[(size_t)g * k + f] * *vp);
    }
}

/* predict_scaled, :171-199 */
static double predict_scaled(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                             const double* grad_w, const double* grad_v, const double* reg_w, const double* reg_v,
                             const uint32_t* group, double lr, uint32_t size, const uint32_t* col,
                             const float* val, double* sum, double* sum_sqr) {
  double p = 0.0;
  if (k0) p += w0;
  if (k1)
    for (uint32_t i = 0; i < size; i++) {
      uint32_t g = group[col[i]];
      double wv = w[col[i]];
      double w_dash = wv - lr * (grad_w[col[i]] + 2 * reg_w[g] * wv);
      p += w_dash * val[i];
    }
  for (int f = 0; f < k; f++) {
    sum[f] = 0.0;
    sum_sqr[f] = 0.0;
    for (uint32_t i = 0; i < size; i++) {
      uint32_t g = group[col[i]];
      double vv = v[(size_t)f * n + col[i]];
      double v_dash = vv - lr * (grad_v[(size_t)f * n + col[i]] + 2 * reg_v[(size_t)g * k + f] * vv);
      double d = v_dash * val[i];
      sum[f] += d;
      sum_sqr[f] += d * d;
    }
    p += 0.5 * (sum[f] * sum[f] - sum_sqr[f]);
  }
  return p;
}

/* sgd_lambda_step, :201-248 */
static void lambda_step(uint32_t n, int k, int k0, int k1, double w0, const double* w, const double* v,
                        const double* grad_w, const double* grad_v, double* reg_w, double* reg_v,
                        const uint32_t* group, uint32_t n_groups, double lr, int task, double min_target,
                        double max_target, uint32_t size, const uint32_t* col, const float* val, float target,
                        double* sum, double* sum_sqr, double* lambda_w_grad, double* sum_f, double* sum_f_dash_f) {
  double p = predict_scaled(n, k, k0, k1, w0, w, v, grad_w, grad_v, reg_w, reg_v, group, lr, size, col, val, sum,
                            sum_sqr);
  double grad_loss = 0;
  if (task == 0) {
    p = fmin(max_target, p);
    p = fmax(min_target, p);
    grad_loss = 2 * (p - target);
  } else if (task == 1) {
    grad_loss = target * ((1.0 / (1.0 + exp(-target * p))) - 1.0);
  }
  if (k1) {
    for (uint32_t g = 0; g < n_groups; g++) lambda_w_grad[g] = 0.0;
    for (uint32_t i = 0; i < size; i++) lambda_w_grad[group[col[i]]] += val[i] * w[col[i]];
    for (uint32_t g = 0; g < n_groups; g++) {
      lambda_w_grad[g] = -2 * lr * lambda_w_grad[g];
      reg_w[g] -= lr * grad_loss * lambda_w_grad[g];
      reg_w[g] = fmax(0.0, reg_w[g]);
    }
  }
  for (int f = 0; f < k; f++) {
    double sum_f_dash = 0.0;
    for (uint32_t g = 0; g < n_groups; g++) {
      sum_f[g] = 0.0;
      sum_f_dash_f[g] = 0.0;
    }
    for (uint32_t i = 0; i < size; i++) {
      uint32_t g = group[col[i]];
      double vv = v[(size_t)f * n + col[i]];
      double v_dash = vv - lr * (grad_v[(size_t)f * n + col[i]] + 2 * reg_v[(size_t)g * k + f] * vv);
      sum_f_dash += v_dash * val[i];
      sum_f[g] += vv * val[i];
      sum_f_dash_f[g] += v_dash * val[i] * vv * val[i];
    }
    for (uint32_t g = 0; g < n_groups; g++) {
      double lambda_v_grad = -2 * lr * (sum_f_dash * sum_f[g] - sum_f_dash_f[g]);
      reg_v[(size_t)g * k + f] -= lr * grad_loss * lambda_v_grad;
      reg_v[(size_t)g * k + f] = fmax(0.0, reg_v[(size_t)g * k + f]);
    }
  }
}

/* one epoch of fm_learn_sgd_element_adapt_reg::learn, :295-311.  lambda_steps = 0 for the first
 * epoch (:301).  The validation cursor restarts at every epoch (:297) and wraps (:302-305). */
void fmo_sgda_epoch(uint32_t n, int k, int k0, int k1, double* w0, double* w, double* v, double* grad_w,
                    double* grad_v, double* reg_w, double* reg_v, const uint32_t* group, uint32_t n_groups,
                    double lr, int task, double min_target, double max_target, int lambda_steps, uint64_t n_rows,
                    const uint64_t* row_ptr, const uint32_t* col, const float* val, const float* target,
                    uint64_t v_rows, const uint64_t* v_row_ptr, const uint32_t* v_col, const float* v_val,
                    const float* v_target) {
  double* sum = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double* sum_sqr = (double*)malloc(sizeof(double) * (k > 0 ? k : 1));
  double* lwg = (double*)malloc(sizeof(double) * n_groups);
  double* sf = (double*)malloc(sizeof(double) * n_groups);
  double* sfd = (double*)malloc(sizeof(double) * n_groups);
  uint64_t vc = 0;
  for (uint64_t r = 0; r < n_rows; r++) {
    theta_step(n, k, k0, k1, w0, w, v, grad_w, grad_v, reg_w, reg_v, group, lr, task, min_target, max_target,
               (uint32_t)(row_ptr[r + 1] - row_ptr[r]), col + row_ptr[r], val + row_ptr[r], target[r], sum,
               sum_sqr);
    if (lambda_steps && v_rows > 0) {
      if (vc == v_rows) vc = 0;
      lambda_step(n, k, k0, k1, *w0, w, v, grad_w, grad_v, reg_w, reg_v, group, n_groups, lr, task, min_target,
                  max_target, (uint32_t)(v_row_ptr[vc + 1] - v_row_ptr[vc]), v_col + v_row_ptr[vc],
                  v_val + v_row_ptr[vc], v_target[vc], sum, sum_sqr, lwg, sf, sfd);
      vc++;
    }
  }
  free(sum);
  free(sum_sqr);
  free(lwg);
  free(sf);
  free(sfd);
}
